"""Host-side policies of the package that need no GPU."""


def test_default_queue_levels_rule():
    """How many pyramid levels join the work queue (algorithms.default_queue_levels): pairs AND warp rows decide, narrow
    levels need the staged routine's narrow form (no object masks), the coarsest level of a 4-level pyramid never joins."""
    from deep_prob_feature_track_b200.algorithms import default_queue_levels as f
    tum = [(15, 20), (30, 40), (60, 80), (120, 160)]
    vga = [(60, 80), (120, 160), (240, 320), (480, 640)]
    assert [f(tum, 64 * g) for g in (1, 2, 4, 6, 20)] == [1, 1, 2, 3, 3]
    assert [f(vga, b) for b in (1, 16, 64)] == [1, 1, 2]
    assert f(tum, 1280, masks=True) == 2            # the 30x40 level would run the plain routine: stays launch-per-iteration
    assert f([(120, 160)], 1280) == 1 and f([(60, 82), (120, 164)], 1280) == 1   # W % 4 != 0: not staged


def test_queue_tile_rows_planner():
    """Rows per work-queue tile (csrc/uic_forward.cu: queue_tile_rows) at the sizes the choice was measured on
    (profiles/r2/r2e_fine_tile_rows_probe.txt, r2e_coarse_tile_rows_probe.txt, r2e_vga_tile_rows_probe.txt): whole columns
    once an iteration offers 2.5 waves of tiles, at most 48 rows below that.  Host code: runs without a GPU."""
    import ctypes
    from deep_prob_feature_track_b200 import _lib
    L = _lib.lib()
    L.dpft_debug_queue_tile_rows.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
    L.dpft_debug_queue_tile_rows.restype = ctypes.c_int
    tr = lambda H, W, B: L.dpft_debug_queue_tile_rows(H, W, B, 1776)
    assert tr(120, 160, 64 * 20) == 120 and tr(120, 160, 64 * 12) == 120      # 4.3 / 2.6 waves of whole columns
    assert tr(120, 160, 64 * 8) == 60                                          # whole columns: 1.7 waves; halves: 3.5
    assert tr(120, 160, 64 * 4) <= 48 and tr(120, 160, 64) <= 48
    assert tr(480, 640, 64) == 120 and tr(480, 640, 16) == 48                  # 64 / 16 live frames against one keyframe
    assert tr(60, 80, 64 * 20) == 30 and tr(30, 40, 64 * 20) == 15             # the coarse queue levels keep their heights
    for H, W, B in ((120, 160, 1280), (480, 640, 16), (15, 20, 64), (7, 8, 3)):
        t = tr(H, W, B)
        assert 1 <= t <= H
