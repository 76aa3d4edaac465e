"""Host logic of the staged kernel's balanced tiling (csrc/uic_forward.cu: make_tile_tab / make_tile_tab_linear):
every row of every segment of a pair belongs to exactly one warp tile, for both kinds of pairs, and the CTA
counts fill the resident slots.  Runs without a GPU (the table is built on the host)."""
import ctypes

import pytest

from deep_prob_feature_track_b200 import _lib

MAXW = 40


def tile_table(H, W, B, linear):
    L = _lib.lib()
    tiles = (ctypes.c_int * (2 * MAXW * 2 * 3))()
    ctas = (ctypes.c_int * 2)()
    n_more, nseg, wpc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    on = L.dpft_debug_tile_table(H, W, B, int(linear), MAXW, tiles, ctas, ctypes.byref(n_more), ctypes.byref(nseg),
                                 ctypes.byref(wpc))
    if not on:
        return None
    out = []
    for kind in range(2):
        warps = []
        for w in range(ctas[kind] * wpc.value):
            base = ((kind * MAXW + w) * 2) * 3
            warps.append([tuple(tiles[base + 3 * k: base + 3 * k + 3]) for k in range(2)])
        out.append(warps)
    return out, list(ctas), n_more.value, nseg.value, wpc.value


@pytest.mark.parametrize("linear", [False, True])
@pytest.mark.parametrize("H,W", [(120, 160), (60, 80), (240, 320), (480, 640), (77, 100)])
@pytest.mark.parametrize("B", [32, 64, 100, 128])
def test_every_row_is_walked_exactly_once(H, W, B, linear):
    tab = tile_table(H, W, B, linear)
    if tab is None:
        pytest.skip("rectangular tiling for this shape")
    kinds, ctas, n_more, nseg, wpc = tab
    assert ctas[1] == ctas[0] + 1 and 0 <= n_more < B
    assert ctas[0] * wpc >= nseg                      # at least one warp per segment
    for warps in kinds:
        seen = {}
        for sub in warps:
            for seg, y0, y1 in sub:
                assert 0 <= seg < nseg and 0 <= y0 <= y1 <= H
                for y in range(y0, y1):
                    seen[(seg, y)] = seen.get((seg, y), 0) + 1
        assert len(seen) == nseg * H and set(seen.values()) == {1}
        rows = [sum(y1 - y0 for _, y0, y1 in sub) for sub in warps]
        if linear:
            assert max(rows) - min(rows) <= 5         # equal ranges up to the snapping at segment boundaries


def test_batches_the_table_does_not_cover_fall_back():
    assert tile_table(120, 160, 1, False) is None     # hundreds of CTAs per pair: rectangular tiling
    assert tile_table(120, 160, 4096, False) is None  # fewer slots than pairs
