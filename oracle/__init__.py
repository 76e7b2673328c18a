"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY.

ctypes access to ``oracle/liboracle.so`` (the plain-C restatement in
``oracle.c``) and, when it was built, to ``oracle/_ref/libref_cpu.so`` (the
reference's own ``src/gauss_cpu.c`` + ``src/inverse.c`` compiled unmodified
against OpenBLAS 0.3.15).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; the product (``cuda_matrix_inversion_b200``) never does.

Array convention (same as the reference's ``Array``): flat buffers, matrices
back to back, each column-major with lda == n (reference src/helper.cu:45).
``A[k, r, c]`` "math-indexed" numpy arrays are converted with
:func:`to_colmajor` / :func:`from_colmajor`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OBLAS_DIR = os.environ.get(
    "ORACLE_OPENBLAS_DIR",
    "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs")
_LIB = None
_REF = None
_REF_IO = None


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


def ref_available() -> bool:
    return (os.path.exists(os.path.join(_HERE, "_ref", "libref_cpu.so"))
            and os.path.isdir(_OBLAS_DIR))


def ref() -> C.CDLL:
    """The reference's own CPU path (fp32 only), compiled unmodified."""
    global _REF
    if _REF is None:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        # OpenBLAS' own dependencies sit next to it in the opencv wheel and are
        # not on any search path: preload them by full name.
        import glob
        for pat in ("libquadmath-*.so*", "libgfortran-*.so*"):
            for p in sorted(glob.glob(os.path.join(_OBLAS_DIR, pat))):
                C.CDLL(p, mode=C.RTLD_GLOBAL)
        _REF = C.CDLL(os.path.join(_HERE, "_ref", "libref_cpu.so"))
    return _REF


def ref_io():
    global _REF_IO
    if _REF_IO is None:
        p = os.path.join(_HERE, "_ref", "libref_io.so")
        _REF_IO = C.CDLL(p) if os.path.exists(p) else False
    return _REF_IO or None


# --------------------------------------------------------------------------- helpers
def to_colmajor(a: np.ndarray) -> np.ndarray:
    """A[k, r, c] -> flat column-major buffer (element (r,c) of matrix k at k*m*n + c*m + r)."""
    a = np.asarray(a)
    if a.ndim == 2:
        a = a[None]
    return np.ascontiguousarray(a.transpose(0, 2, 1)).reshape(-1)


def from_colmajor(flat: np.ndarray, m: int, n: int | None = None) -> np.ndarray:
    n = m if n is None else n
    return np.asarray(flat).reshape(-1, n, m).transpose(0, 2, 1)


def read_mats(path: str, dtype=np.float64) -> np.ndarray:
    """Parse a `.mats` file into A[k, i, j] (format: SURVEY.md App. B; helper.cu:15-52)."""
    with open(path) as f:
        tok = f.read().split()
    k, m, n = int(tok[0]), int(tok[1]), int(tok[2])
    data = np.array(tok[3:3 + k * m * n], dtype=np.float64)
    if data.size != k * m * n:
        raise ValueError(f"{path}: truncated ({data.size} of {k * m * n} values)")
    return data.reshape(k, m, n).astype(dtype)


def read_mats_c(path: str) -> np.ndarray:
    """Same file through the C reader in oracle.c; returns the flat column-major fp64 buffer."""
    k, m, n = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_read_mats(path.encode(), C.byref(k), C.byref(m), C.byref(n), None)
    if rc:
        raise OSError(f"orc_read_mats({path}) -> {rc}")
    out = np.empty(k.value * m.value * n.value, dtype=np.float64)
    rc = lib().orc_read_mats(path.encode(), C.byref(k), C.byref(m), C.byref(n),
                             out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OSError(f"orc_read_mats({path}) -> {rc}")
    return out


def _suffix(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- oracle calls
def chol_inverse(flat: np.ndarray, n: int, full: bool = True):
    """SPD inverse of every matrix in the flat column-major batch -> (ainv_flat, info[batch])."""
    flat = np.ascontiguousarray(flat)
    batch = flat.size // (n * n)
    out = np.zeros_like(flat) if full else flat.copy()
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib(), "orc_chol_inverse_batch_" + _suffix(flat.dtype))
    fn(_p(flat), _p(out), C.c_int(n), C.c_long(batch), C.c_int(int(full)), _p(info))
    return out, info


def gauss_jordan_inverse(flat: np.ndarray, n: int):
    flat = np.ascontiguousarray(flat)
    batch = flat.size // (n * n)
    out = np.zeros_like(flat)
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib(), "orc_gauss_jordan_inverse_batch_" + _suffix(flat.dtype))
    fn(_p(flat), _p(out), C.c_int(n), C.c_long(batch), _p(info))
    return out, info


def lu_inverse(flat: np.ndarray, n: int):
    flat = np.ascontiguousarray(flat)
    batch = flat.size // (n * n)
    out = np.zeros_like(flat)
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib(), "orc_lu_inverse_batch_" + _suffix(flat.dtype))
    fn(_p(flat), _p(out), C.c_int(n), C.c_long(batch), _p(info))
    return out, info


def getrf(flat: np.ndarray, n: int):
    """LU factors of every matrix (LAPACK sgetf2 semantics) -> (lu_flat, ipiv[batch, n] 1-based, info[batch])."""
    lu = np.ascontiguousarray(flat).copy()
    batch = lu.size // (n * n)
    ipiv = np.zeros((batch, n), dtype=np.int32)
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib(), "orc_getrf_batch_" + _suffix(lu.dtype))
    fn(_p(lu), C.c_int(n), C.c_long(batch), _p(ipiv), _p(info))
    return lu, ipiv, info


def getrs(lu: np.ndarray, ipiv: np.ndarray, b: np.ndarray, n: int, nrhs: int) -> np.ndarray:
    """Solve with the factors (sgetrs 'N'); b is batch x (n x nrhs column-major)."""
    x = np.ascontiguousarray(b).copy()
    batch = lu.size // (n * n)
    fn = getattr(lib(), "orc_getrs_batch_" + _suffix(lu.dtype))
    fn(_p(np.ascontiguousarray(lu)), _p(np.ascontiguousarray(ipiv, dtype=np.int32)), _p(x), C.c_int(n), C.c_int(nrhs), C.c_long(batch))
    return x


def _gp(which, n, a, b, c, x):
    a, b, c, x = (np.ascontiguousarray(v) for v in (a, b, c, x))
    batch = b.size // (n * n)
    out = np.zeros(batch, dtype=b.dtype)
    info = np.zeros(batch, dtype=np.int32)
    fn = getattr(lib(), "orc_gp_batch_" + _suffix(b.dtype))
    fn(C.c_int(which), C.c_int(n), _p(a), _p(b), _p(c), _p(x), _p(out), C.c_long(batch), _p(info))
    return out, info


def gp_mean(n, a, b, c, d):
    """means[i] = A_i^T (B_i + diag C_i)^-1 D_i   (gauss_cpu.c:46-72)."""
    return _gp(0, n, a, b, c, d)


def gp_variance(n, a, b, c, e):
    """var[i] = E_i - A_i^T (B_i + diag C_i)^-1 A_i   (sign per gauss_cpu.h:34, not gauss_cpu.c:198)."""
    return _gp(1, n, a, b, c, e)


def gp_mean_solve(n, a, b, c, d):
    return _gp(2, n, a, b, c, d)


# --------------------------------------------------------------------------- reference (_ref) calls
def ref_chol_inverse_upper(flat32: np.ndarray, n: int) -> np.ndarray:
    """inverse_chol_blas_omp (reference src/inverse.c:100): in place, upper triangle valid."""
    buf = np.ascontiguousarray(flat32, dtype=np.float32).copy()
    ref().inverse_chol_blas_omp(_p(buf), C.c_int(n), C.c_int(buf.size // (n * n)))
    return buf


def ref_chol_inverse_inplace(buf32: np.ndarray, n: int) -> None:
    """inverse_chol_blas_omp on the caller's own float32 buffer, no copy: what bench.py times (the reference's own
    bench refreshes its input OUTSIDE the timed region too, src/inverse_bench.c:88-97)."""
    assert buf32.dtype == np.float32 and buf32.flags.c_contiguous
    ref().inverse_chol_blas_omp(_p(buf32), C.c_int(n), C.c_int(buf32.size // (n * n)))


def ref_gp_mean_inplace(n, a, b, c, d, out) -> None:
    """calcluateMeanCPU on the caller's own float32 buffers (Bs and Cs are destroyed), no copies: bench.py's timed call."""
    for v in (a, b, c, d, out):
        assert v.dtype == np.float32 and v.flags.c_contiguous
    ref().calcluateMeanCPU(C.c_int(n), _p(a), _p(b), _p(c), _p(d), _p(out), C.c_int(b.size // (n * n)))


def ref_lu_inverse(flat32: np.ndarray, n: int) -> np.ndarray:
    """inverse_lu_blas_omp (reference src/inverse.c:71): in place."""
    buf = np.ascontiguousarray(flat32, dtype=np.float32).copy()
    ref().inverse_lu_blas_omp(_p(buf), C.c_int(n), C.c_int(buf.size // (n * n)))
    return buf


def ref_gp_mean(n, a, b, c, d) -> np.ndarray:
    """calcluateMeanCPU (reference src/gauss_cpu.c:23); works on copies (it destroys Bs, Cs)."""
    a, b, c, d = (np.ascontiguousarray(v, dtype=np.float32).copy() for v in (a, b, c, d))
    batch = b.size // (n * n)
    out = np.zeros(batch, dtype=np.float32)
    ref().calcluateMeanCPU(C.c_int(n), _p(a), _p(b), _p(c), _p(d), _p(out), C.c_int(batch))
    return out


def ref_gp_variance_raw(n, a, b, c, e) -> np.ndarray:
    """calcluateVarianceCPU as shipped: E + A^T M^-1 A (wrong sign)."""
    a, b, c, e = (np.ascontiguousarray(v, dtype=np.float32).copy() for v in (a, b, c, e))
    batch = b.size // (n * n)
    out = np.zeros(batch, dtype=np.float32)
    ref().calcluateVarianceCPU(C.c_int(n), _p(a), _p(b), _p(c), _p(e), _p(out), C.c_int(batch))
    return out
