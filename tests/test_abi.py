"""The C-ABI library builds, loads without a GPU and exports what include/dpft.h declares."""
import ctypes
import os
import re

from deep_prob_feature_track_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = _lib.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "dpft.h")).read()
    declared = set(re.findall(r"\b(dpft_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.exported_symbols())
    L = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(L, name), name
    L.dpft_abi_version.restype = ctypes.c_int
    assert L.dpft_abi_version() == _lib.DPFT_ABI_VERSION


def test_level_struct_matches_header():
    # 12 pointers + 2 int32 = 104 bytes on LP64; offsets are what the kernels read
    assert ctypes.sizeof(_lib.DpftLevel) == 12 * 8 + 8
    assert _lib.DpftLevel.H.offset == 96 and _lib.DpftLevel.W.offset == 100


def test_bad_arguments_are_rejected_without_a_gpu():
    L = _lib.lib()
    arr = (_lib.DpftLevel * 1)()
    arr[0].H, arr[0].W = 8, 8
    # NULL maps -> workspace query reports 0 and an error message
    assert L.dpft_uic_workspace_bytes(arr, 1, 2, 4, 3, 0) == 0
    assert b"required" in L.dpft_last_error()
    assert L.dpft_uic_workspace_bytes(arr, 0, 2, 4, 3, 0) == 0
    assert b"n_levels" in L.dpft_last_error()


def test_package_exports_resolve():
    """Every name the package advertises resolves to an object of the module it is said to live in."""
    import deep_prob_feature_track_b200 as pkg
    for name in pkg.__all__:
        assert getattr(pkg, name) is not None, name
    assert pkg.patch_tracker.__module__.endswith("algorithms")
    assert pkg.BatchedSolver.__module__.endswith("batched")
