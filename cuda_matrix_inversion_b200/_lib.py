"""ctypes loader for libinvgpu.so -- the only way this package computes anything.

There is deliberately no Python/NumPy/torch fallback: if the shared library is missing the
import fails loudly (build it with ``make lib`` or ``python -c 'import __graft_entry__ as g; g.build()'``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("INVGPU_LIB", os.path.join(_HERE, "lib", "libinvgpu.so"))   # override: kernel-variant experiments

_i64 = C.c_longlong
_vp = C.c_void_p
_int = C.c_int


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built and this package has no "
            "CPU fallback. Run `make lib` at the repository root.")
    lib = C.CDLL(LIB_PATH)
    sig = {
        "invgpu_version": (C.c_char_p, []),
        "invgpu_device_count": (_int, []),
        "invgpu_has_lab": (_int, []),
        "invgpu_gp_upper_h2d": (_int, [_int, _int]),
        "invgpu_set_device": (_int, [_int]),
        "invgpu_error_string": (C.c_char_p, [_int]),
        "invgpu_launch_count": (_i64, []),
        "invgpu_tier_name": (C.c_char_p, [_int, _int, _int]),
        "invgpu_host_alloc": (_vp, [C.c_ulonglong]),
        "invgpu_host_free": (None, [_vp]),
        "invgpu_release_workspace": (None, []),
        "invgpu_xfer_roundtrip_host": (_int, [_vp, _vp, C.c_ulonglong, _i64]),
        "invgpu_device_numa_node": (_int, [_int]),
    }
    for sfx in ("f32", "f64"):
        sig[f"invgpu_spd_inverse_{sfx}"] = (_int, [_vp, _vp, _int, _i64, _vp, _vp])
        sig[f"invgpu_spd_factor_{sfx}"] = (_int, [_vp, _vp, _int, _i64, _vp, _vp])
        sig[f"invgpu_general_inverse_{sfx}"] = (_int, [_vp, _vp, _int, _i64, _vp, _vp])
        sig[f"invgpu_spd_stages_ptrs_{sfx}"] = (_int, [_vp, _vp, _int, _int, _int, _vp, _vp])
        sig[f"invgpu_general_inverse_ptrs_{sfx}"] = (_int, [_vp, _vp, _int, _int, _vp, _vp])
        sig[f"invgpu_mixed_spd_inverse_{sfx}"] = (_int, [_vp, _vp, _vp, _i64, _vp, _vp])
        sig[f"invgpu_getrf_{sfx}"] = (_int, [_vp, _int, _vp, _vp, _i64, _vp])
        sig[f"invgpu_getrf_ptrs_{sfx}"] = (_int, [_vp, _int, _vp, _vp, _int, _vp])
        sig[f"invgpu_getri_{sfx}"] = (_int, [_vp, _vp, _vp, _int, _vp, _i64, _vp])
        sig[f"invgpu_getri_ptrs_{sfx}"] = (_int, [_vp, _vp, _vp, _int, _vp, _int, _vp])
        sig[f"invgpu_gesv_{sfx}"] = (_int, [_vp, _vp, _vp, _int, _int, _vp, _i64, _vp])
        sig[f"invgpu_gp_{sfx}"] = (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp])
        sig[f"invgpu_spd_inverse_host_{sfx}"] = (_int, [_vp, _vp, _int, _i64, _vp])
        sig[f"invgpu_general_inverse_host_{sfx}"] = (_int, [_vp, _vp, _int, _i64, _vp])
        sig[f"invgpu_gp_host_{sfx}"] = (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp])
    host_legacy = ["inverse_gauss_batched_gpu", "inverse_lu_cuda_batched_gpu", "inverse_cholesky_batched_gpu",
                   "inverse_cholesky_mm_batched_gpu", "inverse_cholesky_mm2_batched_gpu",
                   "inverse_cholesky_stride_batched_gpu"]
    dev_legacy = ["inverse_gauss_batched_device", "inverse_lu_cuda_batched_device",
                  "inverse_cholesky_stride_batched_device", "decompose_cholesky_stride_batched_device",
                  "inverse_upper_stride_batched_device", "multiply_upper_stride_batched_device",
                  "inverse_cholesky_batched_device", "decompose_cholesky_batched_device",
                  "inverse_cholesky_mm_batched_device", "decompose_cholesky_mm_batched_device",
                  "inverse_cholesky_mm2_batched_device"]
    for name in host_legacy + dev_legacy:
        sig[name] = (None, [_vp, _int, _vp, _vp, _int])
    for name in ("calcluateMeanGPU", "calcluateVarianceGPU", "calcluateMeanSolveGPU", "calcluateVarianceSolveGPU"):
        sig[name] = (None, [_int, _vp, _vp, _vp, _vp, _vp, _int])
    sig["readMatricesFile"] = (None, [C.c_char_p, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_vp)])
    sig["replicateMatrices"] = (None, [C.POINTER(_vp), _int, _int, _int, _int])
    sig["writeMatricesFile"] = (_int, [C.c_char_p, _int, _int, _int, _vp, _int])
    sig["printMatrix"] = (None, [_vp, _int, _int])
    sig["printMatrixList"] = (None, [_vp, _int, _int])
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    lib._declared = sorted(sig)
    return lib


lib = _load()
LEGACY_HOST = ("inverse_gauss_batched_gpu", "inverse_lu_cuda_batched_gpu", "inverse_cholesky_batched_gpu",
               "inverse_cholesky_mm_batched_gpu", "inverse_cholesky_mm2_batched_gpu",
               "inverse_cholesky_stride_batched_gpu")
