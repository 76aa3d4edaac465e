"""B200-native trust-region inverse-compositional solver behind the reference's module surface.

Public API (resolved lazily, so that ``import deep_prob_feature_track_b200`` neither imports torch nor loads
``libdpft.so``; every entry point raises when the library has not been built -- there is no CPU fallback):

* drop-in modules of ``code/models/algorithms.py``: ``TrustRegionInverseWUncertainty``, ``TrustRegionBase``,
  ``DirectSolverNet``; ``patch_tracker(net)`` swaps them into an existing ``LeastSquareTracking``
* functional entry points: ``uic_solve`` / ``uic_track`` (whole coarse-to-fine solve in one C call),
  ``uic_residual_loss``, ``depth_pyramids``, ``KeyframeTracker``
* ``compute_RT_EPE_loss`` (``code/models/criterions.py``)
* throughput front ends: ``BatchedSolver`` (streams, CUDA-graph replay), ``HostStreamSolver`` (pinned host batches)
* data-parallel training: ``FlatBucketReducer``, ``broadcast_parameters``; ``shard_range`` for the tracking workloads
"""
from importlib import import_module

_EXPORTS = {
    "TrustRegionInverseWUncertainty": "algorithms", "TrustRegionBase": "algorithms", "DirectSolverNet": "algorithms",
    "patch_tracker": "algorithms", "uic_solve": "algorithms", "uic_track": "algorithms", "uic_residual_loss": "algorithms",
    "depth_pyramids": "algorithms", "KeyframeTracker": "algorithms", "SolveResult": "algorithms",
    "pack_pose": "algorithms", "unpack_pose": "algorithms", "default_queue_levels": "algorithms",
    "compute_RT_EPE_loss": "criterions",
    "BatchedSolver": "batched", "HostStreamSolver": "batched", "bind_to_gpu_numa_node": "batched",
    "FlatBucketReducer": "ddp", "broadcast_parameters": "ddp",
    "shard_range": "sharding", "max_over_ranks": "sharding", "gather_poses": "sharding",
    "make_frame_pairs": "synthetic",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        return getattr(import_module(f"{__name__}.{_EXPORTS[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + __all__)
