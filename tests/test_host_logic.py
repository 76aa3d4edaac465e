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
