"""patch_tracker swaps the reference's tr_update0..3 children for the CUDA-backed classes and keeps every
learned child under the same attribute name, so checkpoints still load (no GPU needed: nothing is called)."""
import pytest
import torch.nn as nn

from deep_prob_feature_track_b200 import algorithms as A


def fake_reference_class(name, fields):
    """A stand-in with the reference's class NAME and attributes (patch_tracker dispatches on the name)."""
    def __init__(self, **kw):
        nn.Module.__init__(self)
        for k in fields:
            setattr(self, k, kw.get(k))
    return type(name, (nn.Module,), {"__init__": __init__})


UIC = fake_reference_class("TrustRegionInverseWUncertainty",
                           ["max_iterations", "mEstimator", "directSolver", "timers", "uncer_prop", "combine_icp",
                            "scale_func", "remove_tru_sigma"])
IC = fake_reference_class("TrustRegionBase", ["max_iterations", "mEstimator", "directSolver", "timers"])


def build(cls, **kw):
    net = nn.Module()
    for i in range(4):
        setattr(net, f"tr_update{i}", cls(max_iterations=3, **kw))
    return net


def test_uic_children_are_swapped_and_state_dict_keys_survive():
    scaler = nn.Linear(2, 1)
    net = build(UIC, scale_func=scaler, remove_tru_sigma=True, combine_icp=True, uncer_prop=False)
    before = set(net.state_dict().keys())
    A.patch_tracker(net)
    for i in range(4):
        m = getattr(net, f"tr_update{i}")
        assert isinstance(m, A.TrustRegionInverseWUncertainty)
        assert m.remove_tru_sigma and m.combine_icp and m.max_iterations == 3 and m.scale_func is scaler
    assert set(net.state_dict().keys()) == before and "tr_update0.scale_func.weight" in before


def test_ic_children_keep_their_networks():
    mest, solver = nn.Conv2d(4, 1, 3), A.DirectSolverNet("Direct-ResVol")
    net = build(IC, mEstimator=mest, directSolver=solver)
    before = set(net.state_dict().keys())
    A.patch_tracker(net)
    m = net.tr_update2
    assert isinstance(m, A.TrustRegionBase) and m.mEstimator is mest and m.directSolver is solver
    assert set(net.state_dict().keys()) == before
    assert any(k.startswith("tr_update0.directSolver.net.0.0.") for k in before)   # the reference's key layout


def test_other_trackers_are_refused():
    other = fake_reference_class("Inverse_ICP", ["max_iterations"])
    with pytest.raises(NotImplementedError):
        A.patch_tracker(build(other))


def test_direct_solver_net_mirrors_the_reference_layout():
    s = A.DirectSolverNet("Direct-ResVol", samples=10)
    assert s.type == A.DirectSolverNet.SOLVER_RESIDUAL_VOLUME and s.samples == 10
    assert s.net[0][0].in_features == 96 and s.net[2][0].out_features == 6
    assert A.DirectSolverNet("Direct-Nodamping").net is None
    with pytest.raises(NotImplementedError):
        A.DirectSolverNet("something-else")


def test_real_reference_tracker_is_patched():
    """The same swap on a LeastSquareTracking built by the reference's own constructor (baseline/_ref)."""
    from baseline import reference as REF
    if not REF.available():
        pytest.skip("baseline/_ref not installed (python baseline/install_reference.py)")
    net = REF.make_tracker()
    before = set(net.state_dict().keys())
    A.patch_tracker(net)
    for i in range(4):
        m = getattr(net, f"tr_update{i}")
        assert isinstance(m, A.TrustRegionInverseWUncertainty) and m.remove_tru_sigma and m.max_iterations == 3
    assert set(net.state_dict().keys()) == before
