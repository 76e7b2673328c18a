"""Drop-in boundary on a GPU: the reference's UNMODIFIED bench programs relinked against libinvgpu.so
(oracle/Makefile `relink`), direct comparisons with the reference's own CPU path compiled unmodified
(oracle/_ref/libref_cpu.so), the multi-GPU CLI mode, the detailed-logging mode and the transfer probe.
Nothing here reads /root/reference: the relinked binaries and libref_cpu.so are prebuilt and travel."""
import ctypes as C
import os
import re
import subprocess
import threading

import numpy as np
import pytest

import oracle as orc
from tests.util import TOL, normwise_err, residual_inf, spd_batch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
REF = os.path.join(ROOT, "oracle", "_ref")
FIX = os.path.join(ROOT, "tests", "golden", "reference")
FLT = r"[-+]?\d\.\d+e[-+]\d+"


@pytest.fixture(scope="module")
def api():
    from cuda_matrix_inversion_b200 import api as _api
    assert _api.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return _api


def _run(*args, env=None, **kw):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(list(args), capture_output=True, text=True, cwd=ROOT, env=e, **kw)


# ------------------------------------------------------------------ GPU result vs the reference's own CPU code
needs_ref = pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref/libref_cpu.so not built")


@needs_ref
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128])
def test_spd_inverse_vs_reference_cpu_direct(api, n):
    """inverse_chol_blas_omp (src/inverse.c:100, LAPACK spotrf/spotri) on the same input: upper triangle."""
    a = spd_batch(n, 41, np.float32, seed=900 + n)
    flat = orc.to_colmajor(a)
    got, info = api.spd_inverse_host(flat, n)
    assert not info.any()
    ref = orc.from_colmajor(orc.ref_chol_inverse_upper(flat.copy(), n), n)
    got3 = orc.from_colmajor(got, n)
    iu = np.triu_indices(n)
    err = np.abs(got3[:, iu[0], iu[1]] - ref[:, iu[0], iu[1]]).max() / np.abs(ref[:, iu[0], iu[1]]).max()
    print(f"[vs _ref] spd n={n}: {err:.2e}")
    assert err <= 1e-5


@needs_ref
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128])
def test_general_inverse_vs_reference_cpu_direct(api, fixtures_dir, n):
    """inverse_lu_blas_omp (src/inverse.c:71, LAPACK sgetrf/sgetri) on the reference's square_5_* fixtures."""
    a = orc.read_mats(os.path.join(fixtures_dir, f"square_5_{n}_{n}.mats"), np.float32)
    flat = orc.to_colmajor(a)
    got, info = api.general_inverse_host(flat, n)
    assert not info.any()
    ref = orc.from_colmajor(orc.ref_lu_inverse(flat.copy(), n), n)
    exact = np.linalg.inv(a.astype(np.float64))
    e_gpu, e_ref = normwise_err(orc.from_colmajor(got, n), exact), normwise_err(ref, exact)
    print(f"[vs _ref] general n={n}: gpu {e_gpu:.2e} reference LAPACK {e_ref:.2e}")
    assert e_gpu <= 1e-4
    assert e_gpu <= max(1e-5, 4 * e_ref)              # as accurate as the reference's own CPU path
    assert normwise_err(orc.from_colmajor(got, n), ref) <= 1e-4


@needs_ref
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128])
def test_gp_mean_vs_reference_cpu_direct(api, n):
    """calcluateMeanCPU (src/gauss_cpu.c:23) on the same input (it destroys Bs / Cs: pass copies)."""
    from tests.util import gp_batch
    g = gp_batch(n, 53, np.float32, seed=70 + n)
    flat = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1).copy() for k, v in g.items()}
    means, _, info = api.gp_host(n, flat["a"], flat["b"], flat["c"], Ds=flat["d"])
    assert not info.any()
    ref = orc.ref_gp_mean(n, flat["a"].copy(), flat["b"].copy(), flat["c"].copy(), flat["d"].copy())
    err = np.abs(means - ref).max() / max(1.0, np.abs(ref).max())
    print(f"[vs _ref] gp mean n={n}: {err:.2e}")
    assert err <= 1e-5


# ------------------------------------------------------------------ relinked reference programs
def _have(name):
    return os.path.exists(os.path.join(REF, name))


@pytest.mark.skipif(not _have("relink_inverse_bench"), reason="oracle/_ref/relink_inverse_bench not built (make -C oracle relink)")
def test_reference_inverse_bench_relinked_against_libinvgpu():
    """reference src/inverse_bench.c + src/inverse.c, unmodified, compiled against include/*.h of this repo and
    linked with -linvgpu instead of the reference's GPU objects: all six rows, errors vs the MATLAB goldens."""
    r = _run(os.path.join(REF, "relink_inverse_bench"), os.path.join(FIX, "inverse_100_16x16"), "3", "2", "-csv",
             env={"OMP_NUM_THREADS": "4", "OPENBLAS_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr
    rows = [ln for ln in r.stdout.splitlines() if ln.strip()]
    names = ["lu_blas_cpu", "lu_blas_omp_cpu", "chol_gpu", "chol_mm2_gpu", "gauss_batched_gpu", "lu_cuda_batched_gpu"]
    assert len(rows) == 6, r.stdout
    for row, name in zip(rows, names):
        m = re.fullmatch(rf"200 16 3 {name} ({FLT}) ({FLT}) ({FLT}) ({FLT})\s*", row)
        assert m, row
        assert float(m.group(4)) < 5e-3, row


@pytest.mark.skipif(not _have("relink_gauss_bench"), reason="oracle/_ref/relink_gauss_bench not built (make -C oracle relink)")
def test_reference_gauss_bench_relinked_against_libinvgpu():
    """reference src/gauss_bench.cu, unmodified: its own add / gemv / dot kernels and cuBLAS calls around OUR
    inverse_lu_cuda_batched_device on batchedCudaMalloc'd pitched pointer arrays in pinned host memory."""
    r = _run(os.path.join(REF, "relink_gauss_bench"), os.path.join(FIX, "gaussian_100_32x32"), "2", "2", "-csv",
             env={"OMP_NUM_THREADS": "4", "OPENBLAS_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr
    rows = {ln.split()[3]: ln.split() for ln in r.stdout.splitlines() if len(ln.split()) >= 6}
    assert {"means_cpu", "means_gpu", "variances_cpu", "variances_gpu"} <= set(rows), r.stdout
    assert float(rows["means_gpu"][-1]) < 1e-3, rows["means_gpu"]          # mean |gpu - golden| per evaluation
    assert float(rows["variances_gpu"][-1]) < 1e-3, rows["variances_gpu"]


# ------------------------------------------------------------------ CLI: --gpus N, detailed logging
def test_inverse_bench_two_gpus_bit_identical_and_faster(api, tmp_path):
    """`bin/inverse_bench --gpus 2`: one host thread per device, each through its own device's pipeline (per-device engine
    lock).  Output bit-identical to --gpus 1, and the two devices really run side by side: the speed-up must reach 85 % of what
    the BOX allows for two concurrent host<->device streams -- measured on the spot with the copy-only probe
    (tools/bin/xfer_bench, `pipeline` rows; 1.77x on the pool's 2-GPU boxes, only 1.2x on its 8-GPU boxes, whose aggregate
    host path saturates near 105-130 GB/s: profiles/r2_xfer_*.csv)."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = os.path.join(FIX, "inverse_100_32x32")
    outs = {}
    for g in (1, 2):
        r = _run(os.path.join(BIN, "inverse_bench"), d, "3", "500", "-csv", "--gpus", str(g), "--json", "--dump", str(tmp_path / f"inv{g}.bin"))
        assert r.returncode == 0, r.stderr
        outs[g] = r.stdout
    a = np.fromfile(tmp_path / "inv1.bin", dtype=np.float32)
    b = np.fromfile(tmp_path / "inv2.bin", dtype=np.float32)
    assert a.size == 50000 * 32 * 32 and np.array_equal(a, b)             # same kernels, disjoint shards: bit-identical
    rate = {g: float(re.search(r'"inversions_per_s": ([-+.e\d]+)', outs[g]).group(1)) for g in (1, 2)}
    xb = _run(os.path.join(ROOT, "tools", "bin", "xfer_bench"), "--gpus", "2", "--quick", "--mb", "512")
    assert xb.returncode == 0, xb.stderr
    ceil = {}
    for ln in xb.stdout.splitlines():
        f = ln.split(",")
        if len(f) == 8 and f[4] == "pipeline" and f[2] == "1":
            ceil[int(f[0])] = float(f[7])
    allowed = ceil[2] / ceil[1]
    print(f"[--gpus] 1 GPU {rate[1]:.3e} inv/s, 2 GPUs {rate[2]:.3e} inv/s ({rate[2] / rate[1]:.2f}x); the box's copy-only pipeline: "
          f"{ceil[1]:.1f} -> {ceil[2]:.1f} GB/s ({allowed:.2f}x)")
    assert rate[2] / rate[1] >= 0.85 * allowed
    assert rate[2] / rate[1] >= 1.05                                       # and never serialised


def test_detailed_logging_phase_lines():
    """`make log=1` / INVGPU_DETAILED_LOGGING=1: TIMER_LOG lines name,batch,n,ms,ns (reference include/timer.h:8-9)."""
    r = _run(os.path.join(BIN, "inverse_bench"), os.path.join(FIX, "inverse_100_16x16"), "1", "1", "-csv",
             env={"INVGPU_DETAILED_LOGGING": "1"})
    assert r.returncode == 0, r.stderr
    for timer in ("decompose_cholesky_batched_gpu", "cholesky_mm2_batched_gpu", "inverse_gauss_batched_gpu",
                  "inverse_lu_cuda_batched_gpu"):
        for phase in ("mem_htod", "ker", "mem_dtoh"):
            assert re.search(rf"^{timer}_{phase},100,16,\d+\.\d{{4}},\d+\r?$", r.stdout, re.M), (timer, phase, r.stdout)
    r = _run(os.path.join(BIN, "gauss_bench"), os.path.join(FIX, "gaussian_100_16x16"), "1", "1", "-csv",
             env={"INVGPU_DETAILED_LOGGING": "1"})
    assert r.returncode == 0, r.stderr
    for phase in ("mem_htod", "add", "inv", "mul", "dot", "mem_dtoh"):
        assert re.search(rf"^calculate_mean_gpu_{phase},100,16,", r.stdout, re.M), (phase, r.stdout)


# ------------------------------------------------------------------ transfer probe, batchedCudaMalloc
def test_xfer_roundtrip_probe_copies_exactly():
    from cuda_matrix_inversion_b200 import lib
    rng = np.random.default_rng(3)
    src = rng.integers(0, 255, size=(3000, 4096), dtype=np.uint8)            # pageable: staged through the ring
    dst = np.zeros_like(src)
    rc = lib.invgpu_xfer_roundtrip_host(src.ctypes.data, dst.ctypes.data, 4096, 3000)
    assert rc == 0
    np.testing.assert_array_equal(src, dst)
    nbytes = 40 << 20
    p_in, p_out = lib.invgpu_host_alloc(nbytes), lib.invgpu_host_alloc(nbytes)   # pinned: DMA'd directly
    assert p_in and p_out
    a = np.ctypeslib.as_array(C.cast(p_in, C.POINTER(C.c_uint8)), shape=(nbytes,))
    b = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_uint8)), shape=(nbytes,))
    a[:] = rng.integers(0, 255, size=nbytes, dtype=np.uint8)
    b[:] = 0
    assert lib.invgpu_xfer_roundtrip_host(p_in, p_out, 1 << 20, 40) == 0
    assert np.array_equal(a, b)
    lib.invgpu_host_free(p_in); lib.invgpu_host_free(p_out)
    assert lib.invgpu_device_numa_node(0) >= -1


def test_batched_cuda_malloc_pitched_pointer_arrays(api):
    """The allocation pattern of every upstream `_device` caller (src/gauss_bench.cu:160-167): pointer arrays in PINNED
    host memory (cudaHostAlloc upstream, batched_invert.cu:120-121) filled by batchedCudaMalloc, 512-byte pitch, lda = n."""
    import torch
    from cuda_matrix_inversion_b200 import lib
    n, batch = 24, 50
    lib.batchedCudaMalloc.restype = C.c_int
    lib.batchedCudaMalloc.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_size_t, C.c_int]
    ins = torch.zeros(batch, dtype=torch.int64).pin_memory()
    outs = torch.zeros(batch, dtype=torch.int64).pin_memory()
    pitch = C.c_size_t(0)
    assert lib.batchedCudaMalloc(ins.data_ptr(), C.byref(pitch), n * n * 4, batch) == 0
    assert lib.batchedCudaMalloc(outs.data_ptr(), C.byref(pitch), n * n * 4, batch) == 0
    p = int(pitch.value)
    assert p >= n * n * 4 and p % 512 == 0 and int(ins[1] - ins[0]) == p
    a = spd_batch(n, batch, np.float32, seed=5)
    # the second runtime instance shares the primary context: plain cudaMemcpy on the raw pitched pointers
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rt.cudaMemcpy2D.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
    rt.cudaFree.argtypes = [C.c_void_p]
    flat = np.ascontiguousarray(orc.to_colmajor(a).reshape(batch, n * n))
    assert rt.cudaMemcpy2D(int(ins[0]), p, flat.ctypes.data, n * n * 4, n * n * 4, batch, 1) == 0     # host -> device
    got = np.empty((batch, n * n), dtype=np.float32)

    def fetch():
        torch.cuda.synchronize()
        assert rt.cudaMemcpy2D(got.ctypes.data, n * n * 4, int(outs[0]), p, n * n * 4, batch, 2) == 0  # device -> host
        return orc.from_colmajor(got.reshape(-1), n)

    lib.inverse_cholesky_batched_device(None, n, ins.data_ptr(), outs.data_ptr(), batch)
    assert residual_inf(a, fetch()) <= 1e-4
    lib.inverse_gauss_batched_device(None, n, ins.data_ptr(), outs.data_ptr(), batch)
    assert residual_inf(a, fetch()) <= 1e-4
    assert rt.cudaFree(int(ins[0])) == 0 and rt.cudaFree(int(outs[0])) == 0


# ------------------------------------------------------------------ n up to 256 in every family, both dtypes
@pytest.mark.parametrize("n", [129, 170, 241, 256])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_family_serves_n_up_to_256(api, n, dtype):
    """INVGPU_MAX_N_* = 256 for both dtypes: where the working copy exceeds one CTA's shared memory (general fp32
    n > 240, fp64 n > 169; SPD / GP fp64 n > 236) the any-n kernels run on a global scratch slab."""
    tol = TOL[np.dtype(dtype)]
    a = spd_batch(n, 3, dtype, seed=n)
    flat = orc.to_colmajor(a)
    got, info = api.spd_inverse_host(flat, n)
    assert not info.any()
    assert residual_inf(a, orc.from_colmajor(got, n)) <= tol
    rng = np.random.default_rng(n)
    g = (rng.random((3, n, n)) + np.eye(n) * n / 4).astype(dtype)
    g[:, 0, 0] = 0.0
    got, info = api.general_inverse_host(orc.to_colmajor(g), n)
    assert not info.any()
    want, _ = orc.gauss_jordan_inverse(orc.to_colmajor(g), n)
    assert normwise_err(orc.from_colmajor(got, n), orc.from_colmajor(want, n)) <= tol
    assert residual_inf(g, orc.from_colmajor(got, n)) <= tol
    av, cv, dv = (rng.random(3 * n).astype(dtype) for _ in range(3))
    ev = rng.random(3).astype(dtype)
    means, var, info = api.gp_host(n, av, flat, cv, dv, ev)
    assert not info.any()
    om, _ = orc.gp_mean(n, av, flat, cv, dv)
    ov, _ = orc.gp_variance(n, av, flat, cv, ev)
    assert np.abs(means - om).max() <= tol and np.abs(var - ov).max() <= tol
    # beyond 256: a documented error code, not a crash
    from cuda_matrix_inversion_b200 import lib
    big = np.eye(257, dtype=dtype).reshape(-1)
    out = np.empty_like(big)
    fn = getattr(lib, "invgpu_general_inverse_host_" + ("f32" if dtype == np.float32 else "f64"))
    assert fn(big.ctypes.data, out.ctypes.data, 257, 1, None) == -2                # INVGPU_EUNSUPPORTED


# ------------------------------------------------------------------ mixed scheduler: concurrent calls
def test_mixed_scheduler_concurrent_streams_and_threads(api):
    """Two host threads, each on its own stream, issue overlapping mixed-dimension calls: every call owns its work
    list (ring of buffers guarded by a done-event), so results must equal the oracle's for both."""
    import torch
    rng = np.random.default_rng(11)
    results = {}

    def worker(tid):
        torch.cuda.set_device(0)
        st = torch.cuda.Stream()
        for rep in range(3):
            ns = rng.integers(2, 97, size=400).astype(np.int32) if tid == 0 else rng.integers(2, 40, size=900).astype(np.int32)
            mats = [spd_batch(int(n), 1, np.float32, seed=1000 * tid + 10 * rep + i)[0] for i, n in enumerate(ns)]
            offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)])
            buf = torch.from_numpy(np.concatenate([m.T.reshape(-1) for m in mats])).cuda()
            out = torch.zeros_like(buf)
            info = torch.full((len(ns),), -1, dtype=torch.int32, device="cuda")
            pin = (buf.data_ptr() + offs[:-1] * 4).astype(np.uint64)
            pout = (out.data_ptr() + offs[:-1] * 4).astype(np.uint64)
            st.wait_stream(torch.cuda.current_stream())
            api.mixed_spd_inverse_device(pin, pout, ns, np.float32, info.data_ptr(), st.cuda_stream)
            st.synchronize()
            o = out.cpu().numpy()
            worst = 0.0
            for i, n in enumerate(ns):
                inv = o[offs[i]:offs[i + 1]].reshape(n, n).T
                worst = max(worst, np.abs(mats[i].astype(np.float64) @ inv - np.eye(n)).sum(-1).max())
            results[(tid, rep)] = (worst, int(info.abs().max()))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(results) == 6
    for key, (worst, flagged) in results.items():
        assert flagged == 0 and worst <= 1e-4, (key, worst, flagged)


# ------------------------------------------------------------------ ill-conditioned fp32 general inverse
@pytest.mark.parametrize("n", [12, 16, 32, 64, 72, 128])
def test_general_inverse_ill_conditioned_fp32(api, n):
    """The lean lane = row kernels use rcp.approx + one Newton step and the multiplier form z = -a * r: on matrices
    with cond 1e4..1e5 the result must stay as close to the fp64 truth as the (division-based) oracle's."""
    rng = np.random.default_rng(n)
    u, _ = np.linalg.qr(rng.standard_normal((17, n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((17, n, n)))
    s = np.logspace(0, -4.5, n)
    a = ((u * s) @ v.transpose(0, 2, 1)).astype(np.float32)
    flat = orc.to_colmajor(a)
    got, info = api.general_inverse_host(flat, n)
    want, oinfo = orc.gauss_jordan_inverse(flat, n)
    assert not info.any() and not oinfo.any()
    exact = np.linalg.inv(a.astype(np.float64))
    e_gpu = normwise_err(orc.from_colmajor(got, n), exact)
    e_orc = normwise_err(orc.from_colmajor(want, n), exact)
    print(f"[ill-conditioned n={n}] gpu {e_gpu:.2e} oracle {e_orc:.2e}")
    assert e_gpu <= max(1e-4, 4 * e_orc)


# ------------------------------------------------------------------ LU factors with pivots / inverse from factors / solve
@pytest.mark.parametrize("n", [1, 3, 8, 16, 31, 32, 33, 64, 100, 128, 200, 256])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_getrf_getri_gesv_vs_oracle(api, n, dtype):
    """invgpu_getrf / getri / gesv (cublasSgetrfBatched semantics, reference src/gauss/inverse_gpu.cu:24-50) against the
    oracle's sgetf2 / sgetrs restatement (itself pinned to LAPACK pivots in tests/test_oracle.py): pivot indices and
    info bit-exact, factors / inverse / solutions within the tolerance."""
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    tol = TOL[np.dtype(dtype)]
    batch, nrhs = (23, 3) if n <= 64 else (5, 2)
    rng = np.random.default_rng(1000 + n)
    a = (rng.random((batch, n, n)) + np.eye(n) * max(1.0, n / 8)).astype(dtype)   # cond ~ 1e1..1e2
    a[:, 0, 0] = 0.0 if n > 1 else 2.0                     # force an interchange in step 1
    flat = orc.to_colmajor(a)
    st = torch.cuda.current_stream().cuda_stream
    d_a = torch.from_numpy(flat.copy()).cuda()
    d_piv = torch.zeros(batch * n, dtype=torch.int32, device="cuda")
    d_info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    api.getrf_device(d_a.data_ptr(), n, batch, dtype, d_piv.data_ptr(), d_info.data_ptr(), st)
    torch.cuda.synchronize()
    lu_o, ipiv_o, info_o = orc.getrf(flat, n)
    np.testing.assert_array_equal(d_info.cpu().numpy(), info_o)
    np.testing.assert_array_equal(d_piv.cpu().numpy().reshape(batch, n), ipiv_o)
    assert normwise_err(orc.from_colmajor(d_a.cpu().numpy(), n), orc.from_colmajor(lu_o, n)) <= tol
    # inverse from the factors
    d_inv = torch.empty_like(d_a)
    d_info.fill_(-1)
    api.getri_device(d_a.data_ptr(), d_piv.data_ptr(), d_inv.data_ptr(), n, batch, dtype, d_info.data_ptr(), st)
    torch.cuda.synchronize()
    assert not d_info.cpu().numpy().any()
    want, _ = orc.lu_inverse(flat, n)
    got3, want3 = orc.from_colmajor(d_inv.cpu().numpy(), n), orc.from_colmajor(want, n)
    exact = np.linalg.inv(a.astype(np.float64))
    e_gpu, e_orc = normwise_err(got3, exact), normwise_err(want3, exact)
    print(f"[getri n={n} {np.dtype(dtype).name}] gpu {e_gpu:.2e} oracle {e_orc:.2e}")
    assert e_gpu <= max(tol, 4 * e_orc)
    # multi-RHS solve: A := LU, B := X
    b = rng.random((batch, nrhs, n)).astype(dtype)
    d_a2 = torch.from_numpy(flat.copy()).cuda()
    d_b = torch.from_numpy(b.reshape(-1).copy()).cuda()
    d_piv2 = torch.zeros_like(d_piv)
    api.gesv_device(d_a2.data_ptr(), d_b.data_ptr(), n, nrhs, batch, dtype, d_piv2.data_ptr(), d_info.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(d_piv2, d_piv) and torch.equal(d_a2, d_a)          # same factorisation, bit for bit
    x_o = orc.getrs(lu_o, ipiv_o, b.reshape(-1), n, nrhs).reshape(batch, nrhs, n)
    x = d_b.cpu().numpy().reshape(batch, nrhs, n)
    xe = np.linalg.solve(a.astype(np.float64), b.transpose(0, 2, 1).astype(np.float64)).transpose(0, 2, 1)
    e_gpu, e_orc = np.abs(x - xe).max() / np.abs(xe).max(), np.abs(x_o - xe).max() / np.abs(xe).max()
    assert e_gpu <= max(tol, 4 * e_orc), (e_gpu, e_orc)
    assert tdt == d_b.dtype


def test_getrf_singular_flags_like_lapack(api):
    import torch
    a = np.stack([np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [1.0, 1.0, 1.0]]), np.eye(3) * 2, np.zeros((3, 3))])
    flat = orc.to_colmajor(a.astype(np.float32))
    d_a = torch.from_numpy(flat.copy()).cuda()
    d_piv = torch.zeros(9, dtype=torch.int32, device="cuda")
    d_info = torch.full((3,), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    api.getrf_device(d_a.data_ptr(), 3, 3, np.float32, d_piv.data_ptr(), d_info.data_ptr(), st)
    lu_o, ipiv_o, info_o = orc.getrf(flat, 3)
    np.testing.assert_array_equal(d_info.cpu().numpy(), info_o)
    assert list(info_o) == [3, 0, 1]
    np.testing.assert_array_equal(d_piv.cpu().numpy().reshape(3, 3), ipiv_o)
    np.testing.assert_allclose(d_a.cpu().numpy(), lu_o, atol=1e-6)
    d_inv = torch.empty_like(d_a)
    api.getri_device(d_a.data_ptr(), d_piv.data_ptr(), d_inv.data_ptr(), 3, 3, np.float32, d_info.data_ptr(), st)
    torch.cuda.synchronize()
    inv = d_inv.cpu().numpy().reshape(3, 9)
    assert list(d_info.cpu().numpy()) == [3, 0, 1]
    assert np.isnan(inv[0]).all() and np.isnan(inv[2]).all() and np.allclose(inv[1], (np.eye(3) / 2).reshape(-1))


def test_legacy_lu_device_leaves_factors_in_devAs(api):
    """Upstream inverse_lu_cuda_batched_device factors devAs in place (cublasSgetrfBatched, src/gauss/inverse_gpu.cu:24-33)
    and writes the inverse to devAInvs; the drop-in keeps that side effect."""
    import torch
    from cuda_matrix_inversion_b200 import lib
    n, batch = 16, 12
    rng = np.random.default_rng(4)
    a = (rng.random((batch, n, n)) + 4 * np.eye(n)).astype(np.float32)
    flat = orc.to_colmajor(a)
    d_a = torch.from_numpy(flat.copy()).cuda()
    d_inv = torch.zeros_like(d_a)
    # pointer arrays in pinned host memory, as upstream allocates them (cudaHostAlloc, src/gauss/batched_invert.cu:120-121)
    pin = torch.tensor([d_a.data_ptr() + k * n * n * 4 for k in range(batch)], dtype=torch.int64).pin_memory()
    pout = torch.tensor([d_inv.data_ptr() + k * n * n * 4 for k in range(batch)], dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()
    lib.inverse_lu_cuda_batched_device(None, n, pin.data_ptr(), pout.data_ptr(), batch)
    torch.cuda.synchronize()
    lu_o, _, _ = orc.getrf(flat, n)
    assert normwise_err(orc.from_colmajor(d_a.cpu().numpy(), n), orc.from_colmajor(lu_o, n)) <= 1e-4
    assert residual_inf(a, orc.from_colmajor(d_inv.cpu().numpy(), n)) <= 1e-4


# ------------------------------------------------------------------ GP host call: only the upper triangle of B crosses the bus
@pytest.mark.parametrize("n,dtype,batch", [(128, np.float32, 1337), (64, np.float32, 4500), (100, np.float32, 700), (64, np.float64, 1200), (40, np.float64, 900)])
def test_gp_host_sends_upper_triangle_only(api, n, dtype, batch):
    """invgpu_gp_host_* sends the column prefixes of B (strided 3-D copies, capi.cu
    h2d_upper_triangle) instead of whole matrices.  Several pipeline chunks + a ragged tail, pinned and pageable B: the result
    must be bit-identical to the device-resident call on the full matrices, and a B whose strictly lower triangle is NaN on the
    host must give the same answer (the lower triangle is neither read nor needed)."""
    import torch
    from cuda_matrix_inversion_b200 import lib
    tdt = torch.float32 if dtype == np.float32 else torch.float64   # 32 MiB chunks: the batch spans several, with a ragged tail
    esz = np.dtype(dtype).itemsize
    assert lib.invgpu_gp_upper_h2d(n, esz) == 1 and lib.invgpu_gp_upper_h2d(32, 4) == 0 and lib.invgpu_gp_upper_h2d(32, 8) == 1
    rng = np.random.default_rng(21)
    r = rng.random((batch, n, n)).astype(dtype)
    b = (r + r.transpose(0, 2, 1) + n * np.eye(n)).astype(dtype)
    a, c, d = (rng.random((batch, n)).astype(dtype) for _ in range(3))
    flat_b = orc.to_colmajor(b)
    # device-resident reference run on whole matrices
    tb, ta, tc, td = (torch.from_numpy(x.reshape(-1).copy()).cuda() for x in (flat_b, a, c, d))
    out_m = torch.zeros(batch, device="cuda", dtype=tdt)
    d_info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    api.gp_device(n, ta.data_ptr(), tb.data_ptr(), tc.data_ptr(), td.data_ptr(), 0, out_m.data_ptr(), 0, batch, dtype,
                  d_info.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = out_m.cpu().numpy()
    assert int(d_info.abs().max()) == 0
    # host call, pageable B (staged through the ring) and pinned B, lower triangle poisoned on the host
    poisoned = flat_b.reshape(batch, n, n).copy()                  # [matrix, column, row]
    ci, ri = np.arange(n).reshape(n, 1), np.arange(n).reshape(1, n)
    poisoned[:, ri > ci] = np.nan
    for name, hb in (("pageable", poisoned.reshape(-1)), ("pinned", torch.from_numpy(poisoned.reshape(-1)).pin_memory().numpy())):
        means, _, info = api.gp_host(n, a.reshape(-1), hb, c.reshape(-1), Ds=d.reshape(-1))
        assert not info.any(), name
        assert np.array_equal(means, want), (name, np.abs(means - want).max())
    om, _ = orc.gp_mean(n, a.reshape(-1)[: 8 * n], flat_b[: 8 * n * n], c.reshape(-1)[: 8 * n], d.reshape(-1)[: 8 * n])
    assert np.abs(want[:8] - om).max() <= (1e-4 if dtype == np.float32 else 1e-10) * max(1.0, np.abs(om).max())

