"""ctypes binding of the C-ABI library (include/dpft.h) and its in-tree build.

The library is built IN-TREE as ``deep_prob_feature_track_b200/libdpft.so`` by ``build()`` (called from
``__graft_entry__.build()``): a plain ``nvcc -shared`` for sm_100a, no torch headers -- the ABI is plain
pointers and sizes.  There is no fallback of any kind: if the library is missing or a CUDA device is not
there, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from typing import List, Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# DPFT_LIB_PATH points at another build of the SAME library (tuning sweeps build several); nothing else is ever loaded
LIB_PATH = os.environ.get("DPFT_LIB_PATH") or os.path.join(PKG_DIR, "libdpft.so")
SOURCES = ["dpft_abi.cu", "uic_forward.cu", "uic_queue.cu", "uic_backward.cu", "icp_term.cu", "uic_residual.cu", "context.cu",
           "ic_path.cu", "ic_backward.cu", "uic_persistent.cu", "preprocess.cu", "pose_loss.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(PKG_DIR, "build")

DPFT_ABI_VERSION = 1
DPFT_MAX_LEVELS = 8
DPFT_REMOVE_TRU_SIGMA = 0x01
DPFT_COMBINE_ICP = 0x02
DPFT_NO_PDL = 0x04
DPFT_FUSED_SOBEL = 0x08
DPFT_LAUNCH_PER_ITERATION = 0x10
DPFT_STAGED_FOOTPRINT = 0x20
DPFT_SHARED_KEYFRAME = 0x40
DPFT_PAIRWISE_EXTREMES = 0x80
DPFT_SIGMA_BROADCAST = 0x100
DPFT_QUEUE = 0x200
DPFT_ST_NONFINITE = 0x01
DPFT_ST_SINGULAR = 0x02

c_float_p = ctypes.c_void_p   # device pointers travel as integers


class DpftLevel(ctypes.Structure):
    """struct dpft_level (include/dpft.h)."""
    _fields_ = [
        ("x0", c_float_p), ("x1", c_float_p),
        ("sigma0", c_float_p), ("sigma1", c_float_p),
        ("invd0", c_float_p), ("invd1", c_float_p),
        ("depth0", c_float_p), ("depth1", c_float_p),
        ("K", c_float_p),
        ("obj_mask0", ctypes.c_void_p), ("obj_mask1", ctypes.c_void_p),
        ("occ_out", ctypes.c_void_p),
        ("H", ctypes.c_int32), ("W", ctypes.c_int32),
    ]


class DpftUicOptions(ctypes.Structure):
    """struct dpft_uic_options (include/dpft.h); zero fields mean "default"."""
    _fields_ = [
        ("struct_bytes", ctypes.c_uint32), ("group", ctypes.c_int32),
        ("tile_rows", ctypes.c_int32 * DPFT_MAX_LEVELS),
        ("queue_ctas", ctypes.c_int32), ("queue_levels", ctypes.c_int32), ("cta_slots", ctypes.c_int32),
        ("tiling", ctypes.c_int32),
        ("generic_geometry", ctypes.c_int32),
        ("launch_ms", ctypes.POINTER(ctypes.c_float)),
        ("icp_weight", ctypes.c_void_p * DPFT_MAX_LEVELS),
        ("queue_kernel_ms", ctypes.POINTER(ctypes.c_float)),
        ("small_levels", ctypes.c_int32), ("sigma_detect", ctypes.c_int32),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.struct_bytes = ctypes.sizeof(DpftUicOptions)
        tile_rows = kw.pop("tile_rows", None)
        for i, v in enumerate(tile_rows or ()):
            self.tile_rows[i] = int(v)
        for i, v in enumerate(kw.pop("icp_weight", None) or ()):
            self.icp_weight[i] = v
        for k, v in kw.items():
            setattr(self, k, v)


class DpftLevelGrad(ctypes.Structure):
    """struct dpft_level_grad (include/dpft.h)."""
    _fields_ = [("g_x0", c_float_p), ("g_x1", c_float_p), ("g_sigma0", c_float_p), ("g_sigma1", c_float_p)]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libdpft.so")


def _headers() -> List[str]:
    return ([os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")]
            + [os.path.join(REPO_DIR, "include", "dpft.h")])


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(REPO_DIR, "include", "dpft.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a (cross-compiles without a GPU) and link libdpft.so.  Every translation unit
    becomes an object file under build/ (compiled in parallel, recompiled only when it or a header changed);
    temporary names carry the pid so that several ranks building at once cannot install a truncated file."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("DPFT_NVCC_EXTRA", "").split()   # e.g. -DDPFT_STAGED_COLS=32 for tuning sweeps
    tag_path = os.path.join(OBJ_DIR, "flags.txt")
    tag = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(tag_path) or open(tag_path).read() != tag:
        force = True
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    pid = os.getpid()

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(hdr_t, os.path.getmtime(path)):
            return obj, ""
        tmp = f"{obj}.{pid}.tmp"
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", os.path.join(REPO_DIR, "include"), "-I", CSRC, "-c", "-o", tmp, path]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        os.replace(tmp, obj)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    tmp = f"{LIB_PATH}.{pid}.tmp"
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *[o for o, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    with open(tag_path, "w") as f:
        f.write(tag)
    if verbose:
        print("".join(err for _, err in results))
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA library is the only implementation of this path)")
    L = ctypes.CDLL(LIB_PATH)
    L.dpft_abi_version.restype = ctypes.c_int
    L.dpft_last_error.restype = ctypes.c_char_p
    L.dpft_uic_workspace_bytes.restype = ctypes.c_size_t
    L.dpft_uic_workspace_bytes.argtypes = [ctypes.POINTER(DpftLevel), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_uint32]
    L.dpft_uic_forward.restype = ctypes.c_int
    L.dpft_uic_forward.argtypes = [ctypes.POINTER(DpftLevel), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.dpft_uic_forward_timed.restype = ctypes.c_int
    L.dpft_uic_forward_timed.argtypes = L.dpft_uic_forward.argtypes + [ctypes.POINTER(ctypes.c_float)]
    L.dpft_uic_workspace_bytes_ex.restype = ctypes.c_size_t
    L.dpft_uic_workspace_bytes_ex.argtypes = L.dpft_uic_workspace_bytes.argtypes + [ctypes.POINTER(DpftUicOptions)]
    L.dpft_uic_forward_ex.restype = ctypes.c_int
    L.dpft_uic_forward_ex.argtypes = L.dpft_uic_forward.argtypes + [ctypes.POINTER(DpftUicOptions)]
    L.dpft_uic_backward_workspace_bytes.restype = ctypes.c_size_t
    L.dpft_uic_backward_workspace_bytes.argtypes = L.dpft_uic_workspace_bytes.argtypes
    L.dpft_uic_backward.restype = ctypes.c_int
    L.dpft_uic_backward.argtypes = [ctypes.POINTER(DpftLevel), ctypes.POINTER(DpftLevelGrad), ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_float,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                    ctypes.c_void_p]
    L.dpft_uic_residual_workspace_bytes.restype = ctypes.c_size_t
    L.dpft_uic_residual_workspace_bytes.argtypes = [ctypes.POINTER(DpftLevel), ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_uint32]
    L.dpft_uic_residual_loss.restype = ctypes.c_int
    L.dpft_uic_residual_loss.argtypes = [ctypes.POINTER(DpftLevel), ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                         ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_size_t, ctypes.c_void_p]
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lp = ctypes.POINTER(DpftLevel)
    for name, args in (("dpft_ic_gradients", [lp, ci, ci, vp, vp, vp]),
                       ("dpft_ic_residual", [lp, ci, ci, vp, vp, vp, vp]),
                       ("dpft_ic_normal_matrix", [lp, ci, ci, vp, vp, vp, vp, vp]),
                       ("dpft_ic_rhs", [lp, ci, ci, vp, vp, vp, vp, ci, vp, vp]),
                       ("dpft_ic_update", [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
                       ("dpft_ic_gradients_backward", [lp, ci, ci, vp, vp, vp, vp]),
                       ("dpft_ic_residual_backward", [lp, ci, ci, vp, vp, vp, vp, vp, vp]),
                       ("dpft_ic_normal_matrix_backward", [lp, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]),
                       ("dpft_ic_rhs_backward", [lp, ci, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp]),
                       ("dpft_ic_update_backward", [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])):
        fn = getattr(L, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args
    L.dpft_ic_context.restype = ctypes.c_int
    L.dpft_ic_context.argtypes = [lp, ci, vp, vp, ci, ci, vp, vp, vp]
    L.dpft_uic_icp_context.restype = ctypes.c_int
    L.dpft_uic_icp_context.argtypes = [lp, ci, ci, ctypes.c_uint32, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.dpft_pose_epe_loss.restype = ctypes.c_int
    L.dpft_pose_epe_loss.argtypes = [vp] * 7 + [ci] * 4 + [vp, vp]
    L.dpft_pose_epe_loss_backward.restype = ctypes.c_int
    L.dpft_pose_epe_loss_backward.argtypes = [vp] * 7 + [ci] * 4 + [vp, vp, vp, vp]
    L.dpft_preprocess_depth.restype = ctypes.c_int
    L.dpft_preprocess_depth.argtypes = [vp, ci, ci, ci, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), vp,
                                        ctypes.c_size_t, vp]
    if L.dpft_abi_version() != DPFT_ABI_VERSION:
        raise RuntimeError("libdpft.so ABI version mismatch; rebuild")
    _lib = L
    return L


def exported_symbols() -> List[str]:
    """Entry points include/dpft.h declares (kept in sync by tests/test_abi.py)."""
    return ["dpft_abi_version", "dpft_last_error", "dpft_uic_workspace_bytes", "dpft_uic_forward",
            "dpft_uic_workspace_bytes_ex", "dpft_uic_forward_ex", "dpft_uic_forward_timed", "dpft_uic_backward_workspace_bytes", "dpft_uic_backward",
            "dpft_uic_residual_workspace_bytes", "dpft_uic_residual_loss", "dpft_ic_context", "dpft_uic_icp_context",
            "dpft_ic_gradients", "dpft_ic_residual",
            "dpft_ic_normal_matrix", "dpft_ic_rhs", "dpft_ic_update", "dpft_ic_gradients_backward",
            "dpft_ic_residual_backward", "dpft_ic_normal_matrix_backward", "dpft_ic_rhs_backward",
            "dpft_ic_update_backward", "dpft_preprocess_depth", "dpft_pose_epe_loss", "dpft_pose_epe_loss_backward"]


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().dpft_last_error().decode()
        raise RuntimeError(f"{what} failed with code {code}: {msg}")
