"""Python mirror of the reference-facing C API (include/inverse_gpu.h, include/gauss_gpu.h,
include/invgpu.h).  Thin ctypes calls into libinvgpu.so; every function here ends in a CUDA
kernel launch -- there is no CPU path.

Arrays follow the reference's ``Array`` convention: flat buffers, matrices back to back,
column-major with lda == n (reference src/helper.cu:45).  Host-flavour functions take NumPy
arrays; device-flavour functions take raw device addresses (``tensor.data_ptr()``) so that
PyTorch is only ever plumbing for memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib

SPD_POTRF, SPD_TRTRI, SPD_LAUUM, SPD_INVERSE = 1, 2, 4, 7


class InvGpuError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where}: {lib.invgpu_error_string(code).decode()} (code {code})")


def _check(code: int, where: str) -> None:
    if code != 0:
        raise InvGpuError(code, where)


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"unsupported dtype {dtype} (float32 / float64 only)")


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _flat(a, dtype=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.reshape(-1)


def device_count() -> int:
    return lib.invgpu_device_count()


def launch_count() -> int:
    return lib.invgpu_launch_count()


def tier_name(op: str, n: int, dtype=np.float32) -> str:
    opi = {"spd": 0, "general": 1, "gp": 2}[op]
    return lib.invgpu_tier_name(opi, n, np.dtype(dtype).itemsize).decode()


# ----------------------------------------------------------------------------- host flavour
def spd_inverse_host(As: np.ndarray, n: int, out: np.ndarray | None = None):
    """SPD inverse of a dense host batch -> (aInvs, info).  invgpu_spd_inverse_host_*
    (replaces inverse_cholesky_batched_gpu, reference src/inverse_cholesky_gpu.cu:397)."""
    As = _flat(As)
    batch = As.size // (n * n)
    out = np.empty_like(As) if out is None else out
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib, "invgpu_spd_inverse_host_" + _sfx(As.dtype))
    _check(fn(_np_ptr(As), _np_ptr(out), n, batch, _np_ptr(info)), "spd_inverse_host")
    return out, info


def general_inverse_host(As: np.ndarray, n: int, out: np.ndarray | None = None):
    """General inverse (Gauss-Jordan, partial pivoting) of a dense host batch -> (aInvs, info).
    Replaces inverse_gauss_batched_gpu / inverse_lu_cuda_batched_gpu
    (reference src/gauss/batched_invert.cu:99, src/gauss/inverse_gpu.cu:60)."""
    As = _flat(As)
    batch = As.size // (n * n)
    out = np.empty_like(As) if out is None else out
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib, "invgpu_general_inverse_host_" + _sfx(As.dtype))
    _check(fn(_np_ptr(As), _np_ptr(out), n, batch, _np_ptr(info)), "general_inverse_host")
    return out, info


def gp_host(n: int, As, Bs, Cs, Ds=None, Es=None):
    """Fused GP mean and/or variance on host arrays -> (means|None, variances|None, info).
    Replaces calcluateMean / calcluateVariance (reference src/gauss_bench.cu:127, 275)."""
    Bs = _flat(Bs)
    dt = Bs.dtype
    As, Cs = _flat(As, dt), _flat(Cs, dt)
    batch = Bs.size // (n * n)
    Ds = None if Ds is None else _flat(Ds, dt)
    Es = None if Es is None else _flat(Es, dt)
    means = np.empty(batch, dtype=dt) if Ds is not None else None
    variances = np.empty(batch, dtype=dt) if Es is not None else None
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib, "invgpu_gp_host_" + _sfx(dt))
    p = lambda a: None if a is None else _np_ptr(a)
    _check(fn(n, _np_ptr(As), _np_ptr(Bs), _np_ptr(Cs), p(Ds), p(Es), p(means), p(variances), batch,
              _np_ptr(info)), "gp_host")
    return means, variances, info


# ----------------------------------------------------------------------------- legacy names (fp32, abort on error)
def _legacy_host(name: str, n: int, As: np.ndarray) -> np.ndarray:
    As = _flat(As, np.float32)
    out = np.empty_like(As)
    getattr(lib, name)(None, n, _np_ptr(As), _np_ptr(out), As.size // (n * n))
    return out


def inverse_cholesky_batched_gpu(n, As): return _legacy_host("inverse_cholesky_batched_gpu", n, As)
def inverse_cholesky_mm_batched_gpu(n, As): return _legacy_host("inverse_cholesky_mm_batched_gpu", n, As)
def inverse_cholesky_mm2_batched_gpu(n, As): return _legacy_host("inverse_cholesky_mm2_batched_gpu", n, As)
def inverse_cholesky_stride_batched_gpu(n, As): return _legacy_host("inverse_cholesky_stride_batched_gpu", n, As)
def inverse_gauss_batched_gpu(n, As): return _legacy_host("inverse_gauss_batched_gpu", n, As)
def inverse_lu_cuda_batched_gpu(n, As): return _legacy_host("inverse_lu_cuda_batched_gpu", n, As)


def calcluateMeanGPU(n, As, Bs, Cs, Ds) -> np.ndarray:
    As, Bs, Cs, Ds = (_flat(v, np.float32) for v in (As, Bs, Cs, Ds))
    batch = Bs.size // (n * n)
    means = np.empty(batch, dtype=np.float32)
    lib.calcluateMeanGPU(n, _np_ptr(As), _np_ptr(Bs), _np_ptr(Cs), _np_ptr(Ds), _np_ptr(means), batch)
    return means


def calcluateVarianceGPU(n, As, Bs, Cs, Es) -> np.ndarray:
    As, Bs, Cs, Es = (_flat(v, np.float32) for v in (As, Bs, Cs, Es))
    batch = Bs.size // (n * n)
    var = np.empty(batch, dtype=np.float32)
    lib.calcluateVarianceGPU(n, _np_ptr(As), _np_ptr(Bs), _np_ptr(Cs), _np_ptr(Es), _np_ptr(var), batch)
    return var


# ----------------------------------------------------------------------------- device flavour (raw addresses)
def spd_inverse_device(dA: int, dAinv: int, n: int, batch: int, dtype, d_info: int = 0, stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_spd_inverse_" + _sfx(dtype))
    _check(fn(dA, dAinv, n, batch, d_info or None, stream or None), "spd_inverse_device")


def spd_factor_device(dA: int, dL: int, n: int, batch: int, dtype, d_info: int = 0, stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_spd_factor_" + _sfx(dtype))
    _check(fn(dA, dL, n, batch, d_info or None, stream or None), "spd_factor_device")


def general_inverse_device(dA: int, dAinv: int, n: int, batch: int, dtype, d_info: int = 0, stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_general_inverse_" + _sfx(dtype))
    _check(fn(dA, dAinv, n, batch, d_info or None, stream or None), "general_inverse_device")


def gp_device(n: int, dA: int, dB: int, dC: int, dD: int, dE: int, d_means: int, d_vars: int, batch: int, dtype,
              d_info: int = 0, stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_gp_" + _sfx(dtype))
    _check(fn(n, dA, dB, dC, dD or None, dE or None, d_means or None, d_vars or None, batch, d_info or None,
              stream or None), "gp_device")


def getrf_device(dA: int, n: int, batch: int, dtype, d_pivots: int = 0, d_info: int = 0, stream: int = 0) -> None:
    """P A = L U in place with 1-based pivots and sgetrf info (cublasSgetrfBatched semantics)."""
    fn = getattr(lib, "invgpu_getrf_" + _sfx(dtype))
    _check(fn(dA, n, d_pivots or None, d_info or None, batch, stream or None), "getrf_device")


def getri_device(dLU: int, d_pivots: int, dAinv: int, n: int, batch: int, dtype, d_info: int = 0, stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_getri_" + _sfx(dtype))
    _check(fn(dLU, d_pivots, dAinv, n, d_info or None, batch, stream or None), "getri_device")


def gesv_device(dA: int, dB: int, n: int, nrhs: int, batch: int, dtype, d_pivots: int = 0, d_info: int = 0,
                stream: int = 0) -> None:
    """Solve A X = B for nrhs right-hand sides per matrix: A := LU, B := X."""
    fn = getattr(lib, "invgpu_gesv_" + _sfx(dtype))
    _check(fn(dA, d_pivots or None, dB, n, nrhs, d_info or None, batch, stream or None), "gesv_device")


def spd_stages_ptrs_device(d_ptrs_in: int, d_ptrs_out: int, n: int, batch: int, stages: int, dtype,
                           d_info: int = 0, stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_spd_stages_ptrs_" + _sfx(dtype))
    _check(fn(d_ptrs_in, d_ptrs_out, n, batch, stages, d_info or None, stream or None), "spd_stages_ptrs_device")


def general_inverse_ptrs_device(d_ptrs_in: int, d_ptrs_out: int, n: int, batch: int, dtype, d_info: int = 0,
                                stream: int = 0) -> None:
    fn = getattr(lib, "invgpu_general_inverse_ptrs_" + _sfx(dtype))
    _check(fn(d_ptrs_in, d_ptrs_out, n, batch, d_info or None, stream or None), "general_inverse_ptrs_device")


def mixed_spd_inverse_device(ptrs_in: np.ndarray, ptrs_out: np.ndarray, ns: np.ndarray, dtype, d_info: int = 0,
                             stream: int = 0) -> None:
    """Persistent-CTA scheduler over mixed dimensions: host arrays of device addresses and orders."""
    ptrs_in = np.ascontiguousarray(ptrs_in, dtype=np.uint64)
    ptrs_out = np.ascontiguousarray(ptrs_out, dtype=np.uint64)
    ns = np.ascontiguousarray(ns, dtype=np.int32)
    fn = getattr(lib, "invgpu_mixed_spd_inverse_" + _sfx(dtype))
    _check(fn(_np_ptr(ptrs_in), _np_ptr(ptrs_out), _np_ptr(ns), ns.size, d_info or None, stream or None),
           "mixed_spd_inverse_device")


# ----------------------------------------------------------------------------- .mats I/O through the C library
def read_mats_file(path: str) -> np.ndarray:
    """readMatricesFile (include/helper_cpu.h) -> flat column-major float32 buffer plus shape."""
    k, m, n = C.c_int(), C.c_int(), C.c_int()
    ptr = C.c_void_p()
    lib.readMatricesFile(path.encode(), C.byref(k), C.byref(m), C.byref(n), C.byref(ptr))
    count = k.value * m.value * n.value
    buf = (C.c_float * count).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=np.float32).copy()
    C.CDLL(None).free(ptr)
    return out, (k.value, m.value, n.value)
