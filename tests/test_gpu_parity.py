"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs, against the reference's golden fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties.  Tolerances are the north_star's: normwise relative error
and residual ||A A^-1 - I||_inf <= 1e-4 in fp32 and <= 1e-10 in fp64 (see tests/util.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as orc
from tests.util import (TOL, assert_general_parity, general_batch, gp_batch, load_fixture, normwise_err, residual_inf, spd_batch)

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


@pytest.fixture(scope="module")
def api():
    from cuda_matrix_inversion_b200 import api as _api
    assert _api.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return _api


@pytest.fixture(scope="module")
def torch():
    import torch as _t
    assert _t.cuda.is_available()
    return _t


# --------------------------------------------------------------------------------------- SPD inverse
@pytest.mark.parametrize("n", [8, 16, 32, 64])
@pytest.mark.parametrize("dtype", DTYPES)
def test_spd_inverse_reference_fixtures(api, fixtures_dir, n, dtype):
    a = load_fixture(fixtures_dir, f"inverse_100_{n}x{n}/a.mats", dtype)
    flat = orc.to_colmajor(a)
    launches = api.launch_count()
    got, info = api.spd_inverse_host(flat, n)
    assert api.launch_count() > launches, "no CUDA kernel was launched"
    assert not info.any()
    want, oinfo = orc.chol_inverse(flat, n)
    assert not oinfo.any()
    tol = TOL[np.dtype(dtype)]
    got3, want3 = orc.from_colmajor(got, n), orc.from_colmajor(want, n)
    assert normwise_err(got3, want3) <= tol
    assert residual_inf(a, got3) <= tol
    assert np.abs(got3 - got3.transpose(0, 2, 1)).max() <= tol * np.abs(got3).max()   # both triangles written
    gold = os.path.join(fixtures_dir, f"inverse_100_{n}x{n}/aInv.mats")
    if os.path.exists(gold):   # 4-digit MATLAB goldens pin to ~1e-4 absolute
        assert np.abs(got3 - orc.read_mats(gold, np.float64)).max() < 2e-4


def test_spd_inverse_known_answer(api, fixtures_dir):
    a = load_fixture(fixtures_dir, "simpleMean/chol.mats", np.float64)
    want = load_fixture(fixtures_dir, "simpleMean/cholinv.mats", np.float64)
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 2e-4)):      # cond ~ 1185
        got, info = api.spd_inverse_host(orc.to_colmajor(a.astype(dtype)), 4)
        assert info[0] == 0
        assert np.abs(orc.from_colmajor(got, 4)[0] - want[0]).max() <= max(tol, 1e-5)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 8, 13, 16, 24, 31, 32, 33, 47, 64, 96, 100, 128])
@pytest.mark.parametrize("dtype", DTYPES)
def test_spd_inverse_sizes(api, n, dtype):
    batch = 67 if n <= 32 else 19
    a = spd_batch(n, batch, dtype, seed=100 + n)
    flat = orc.to_colmajor(a)
    got, info = api.spd_inverse_host(flat, n)
    want, _ = orc.chol_inverse(flat, n)
    assert not info.any()
    tol = TOL[np.dtype(dtype)]
    assert normwise_err(orc.from_colmajor(got, n), orc.from_colmajor(want, n)) <= tol
    assert residual_inf(a, orc.from_colmajor(got, n)) <= tol


@pytest.mark.parametrize("n", [160, 200, 256])
def test_spd_inverse_256_bucket_fp32(api, n):
    a = spd_batch(n, 5, np.float32, seed=7)
    flat = orc.to_colmajor(a)
    got, info = api.spd_inverse_host(flat, n)
    want, _ = orc.chol_inverse(flat, n)
    assert not info.any()
    assert normwise_err(orc.from_colmajor(got, n), orc.from_colmajor(want, n)) <= 1e-4
    assert residual_inf(a, orc.from_colmajor(got, n)) <= 1e-4


def test_spd_reads_upper_triangle_only(api):
    """spotrf_("U") semantics (reference src/inverse.c:92): garbage below the diagonal is ignored."""
    n = 16
    a = spd_batch(n, 9, np.float64, seed=3)
    dirty = a.copy()
    il = np.tril_indices(n, -1)
    dirty[:, il[0], il[1]] = 1e30
    got, info = api.spd_inverse_host(orc.to_colmajor(dirty), n)
    want, _ = api.spd_inverse_host(orc.to_colmajor(a), n)
    assert not info.any()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_spd_flags_match_oracle(api, fixtures_dir, dtype):
    n = 8
    a = spd_batch(n, 12, np.float64, seed=11)
    a[2] = np.diag([4.0, 3.0, -1.0, 2.0, 1, 1, 1, 1])            # indefinite: info 3
    a[5, 6, 6] = np.nan                                          # NaN pivot: info 7
    a[7] = 1.1                                                    # rank one: info 2
    a[9] = -np.eye(n)                                             # negative definite: info 1
    flat = orc.to_colmajor(a.astype(dtype))
    sentinel = np.full_like(flat, 777.0)
    got, info = api.spd_inverse_host(flat, n, out=sentinel.copy())
    want, oinfo = orc.chol_inverse(flat, n)
    np.testing.assert_array_equal(info, oinfo)
    assert info[2] == 3 and info[5] == 7 and info[7] == 2 and info[9] == 1
    good = info == 0
    tol = TOL[np.dtype(dtype)]
    assert normwise_err(orc.from_colmajor(got, n)[good], orc.from_colmajor(want, n)[good]) <= tol
    # flagged matrices come back as NaN, the rest of the batch is processed
    assert np.isnan(orc.from_colmajor(got, n)[~good]).all()
    # the reference's own singular fixture (simpleMean/b.mats, all 1.1)
    b = load_fixture(fixtures_dir, "simpleMean/b.mats", dtype)
    _, info = api.spd_inverse_host(orc.to_colmajor(b), 2)
    assert info[0] == 2


@pytest.mark.parametrize("n", [16, 32, 64, 128])
@pytest.mark.parametrize("dtype", DTYPES)
def test_spd_flags_sweep_tiers(api, n, dtype):
    """The sweep tiers eliminate in a permuted order; a flagged matrix must still report LAPACK's
    natural-order spotrf info (reference src/inverse.c:92-95) and come back as NaN."""
    a = spd_batch(n, 10, np.float64, seed=21 + n)
    k0 = n // 2 + 3
    d = np.ones(n)
    d[k0 - 1] = -1.0
    a[1] = np.diag(d)                                             # indefinite: info k0
    a[3, n - 1, n - 1] = np.nan                                   # NaN in the last pivot: info n
    a[4] = 1.1                                                    # rank one: info 2
    a[6] = -np.eye(n)                                             # negative definite: info 1
    w = np.linalg.eigvalsh(a[8])
    a[8] -= (w[0] + 0.5 * (w[1] - w[0])) * np.eye(n)              # one clearly negative eigenvalue: the oracle says where
    flat = orc.to_colmajor(a.astype(dtype))
    got, info = api.spd_inverse_host(flat, n, out=np.full_like(flat, 777.0))
    want, oinfo = orc.chol_inverse(flat, n)
    np.testing.assert_array_equal(info, oinfo)
    assert info[1] == k0 and info[3] == n and info[4] == 2 and info[6] == 1 and info[8] > 0
    good = info == 0
    assert good.sum() == 5
    assert normwise_err(orc.from_colmajor(got, n)[good], orc.from_colmajor(want, n)[good]) <= TOL[np.dtype(dtype)]
    assert np.isnan(orc.from_colmajor(got, n)[~good]).all()


@pytest.mark.parametrize("n", [4, 16, 33, 64])
@pytest.mark.parametrize("dtype", DTYPES)
def test_spd_factor(api, torch, n, dtype):
    a = spd_batch(n, 21, dtype, seed=5)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    d_a = torch.from_numpy(orc.to_colmajor(a)).cuda()
    d_l = torch.empty_like(d_a)
    d_info = torch.full((21,), -1, dtype=torch.int32, device="cuda")
    api.spd_factor_device(d_a.data_ptr(), d_l.data_ptr(), n, 21, dtype, d_info.data_ptr(),
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert not d_info.cpu().numpy().any()
    l = orc.from_colmajor(d_l.cpu().numpy(), n).astype(np.float64)
    assert np.abs(np.triu(l, 1)).max() == 0
    want = np.linalg.cholesky(a.astype(np.float64))
    assert normwise_err(l, want) <= TOL[np.dtype(dtype)]
    assert tdt == d_l.dtype


# --------------------------------------------------------------------------------------- general inverse
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128])
@pytest.mark.parametrize("dtype", DTYPES)
def test_general_inverse_reference_fixtures(api, fixtures_dir, n, dtype):
    a = load_fixture(fixtures_dir, f"square_5_{n}_{n}.mats", dtype)
    flat = orc.to_colmajor(a)
    got, info = api.general_inverse_host(flat, n)
    want, oinfo = orc.gauss_jordan_inverse(flat, n)
    assert not info.any() and not oinfo.any()
    got3, want3 = orc.from_colmajor(got, n), orc.from_colmajor(want, n)
    # cond up to 9e3 here; the tolerance is nevertheless asserted outright (tests/util.py::assert_general_parity)
    e_gpu, _ = assert_general_parity(a, got3, want3, dtype, f"square_5_{n}")
    assert e_gpu <= TOL[np.dtype(dtype)]


def test_general_inverse_square_100_64_64_regenerated(api, golden_dir):
    """BASELINE config 2: square_100_64_64.mats is missing from the reference mount
    (.MISSING_LARGE_BLOBS); tests/golden/make_square_100_64_64.py regenerates it."""
    p = os.path.join(golden_dir, "square_100_64_64.npz")
    a = np.load(p)["a"]                         # [100, 64, 64] float32
    from tests.golden.make_square_100_64_64 import generate
    np.testing.assert_array_equal(a, generate())       # the committed fixture is what the script makes
    flat = orc.to_colmajor(a)
    got, info = api.general_inverse_host(flat, 64)
    assert not info.any()
    got3 = orc.from_colmajor(got, 64)
    want, _ = orc.gauss_jordan_inverse(flat, 64)
    want3 = orc.from_colmajor(want, 64)
    # per matrix (cond of U(0,1) 64x64 ranges 1e2..1e5): error vs the fp64 truth within 1e-4 or 4x the oracle's own
    inv64 = np.linalg.inv(a.astype(np.float64))
    den = np.abs(inv64).reshape(100, -1).max(1)
    e_gpu = np.abs(got3 - inv64).reshape(100, -1).max(1) / den
    e_orc = np.abs(want3 - inv64).reshape(100, -1).max(1) / den
    print(f"[parity square_100_64_64] gpu err max {e_gpu.max():.2e} median {np.median(e_gpu):.2e}; oracle max {e_orc.max():.2e}")
    assert (e_gpu <= np.maximum(1e-4, 4 * e_orc)).all()
    assert np.median(e_gpu) <= 1e-4
    res = lambda x: np.abs(a.astype(np.float64) @ x.astype(np.float64) - np.eye(64)).sum(-1).max(-1)
    assert (res(got3) <= np.maximum(1e-4, 4 * res(want3))).all()


@pytest.mark.parametrize("n", [1, 2, 3, 7, 13, 31, 33, 50, 100])
@pytest.mark.parametrize("dtype", DTYPES)
def test_general_inverse_sizes(api, n, dtype):
    rng = np.random.default_rng(n)
    a = (rng.random((23, n, n)) + np.eye(n) * 0.5).astype(dtype)
    a[:, 0, 0] = 0.0          # force a row interchange in step 0 (when n > 1)
    if n == 1:
        a[:, 0, 0] = 2.0
    flat = orc.to_colmajor(a)
    got, info = api.general_inverse_host(flat, n)
    want, oinfo = orc.gauss_jordan_inverse(flat, n)
    np.testing.assert_array_equal(info, oinfo)
    assert not info.any()
    assert_general_parity(a, orc.from_colmajor(got, n), orc.from_colmajor(want, n), dtype, f"sizes n={n}")


@pytest.mark.parametrize("n", [6, 8])                           # 8: the thread-per-matrix kernel (explicit row swaps)
@pytest.mark.parametrize("dtype", DTYPES)
def test_general_flags_match_oracle(api, dtype, n):
    rng = np.random.default_rng(2)
    a = rng.random((8, n, n)) + np.eye(n)
    a[1, :, 2] = 0.0                 # zero column: pivot 3 is exactly zero
    a[4, 3] = a[4, 1]                # duplicate row: exactly singular
    a[6] = 0.0                       # zero matrix: info 1
    flat = orc.to_colmajor(a.astype(dtype))
    got, info = api.general_inverse_host(flat, n)
    want, oinfo = orc.gauss_jordan_inverse(flat, n)
    assert info[1] == 3 and info[6] == 1
    assert info[1] == oinfo[1] and info[6] == oinfo[6]
    # A duplicated row is singular in exact arithmetic only: in floating point the last pivot is either
    # exactly 0 (flag) or a rounding residue (then the "inverse" is astronomically large).  Which of the
    # two happens depends on fused vs. unfused multiply-add (nvcc contracts a - f*p into one FFMA, in
    # this engine as in the reference's own kernel src/gauss/batched_invert.cu:75-80; the C oracle
    # rounds twice), so only the dichotomy itself can be asserted, for both.
    for flag, res in ((info[4], orc.from_colmajor(got, n)[4]), (oinfo[4], orc.from_colmajor(want, n)[4])):
        assert flag != 0 or np.abs(res).max() > 1e5
    good = (info == 0) & (oinfo == 0) & (np.arange(8) != 4)
    assert normwise_err(orc.from_colmajor(got, n)[good], orc.from_colmajor(want, n)[good]) <= TOL[np.dtype(dtype)]


@pytest.mark.parametrize("n", [40, 64, 100, 128])
@pytest.mark.parametrize("dtype", DTYPES)
def test_general_flags_tile_tiers(api, n, dtype):
    """Orders above 32 run on the lane = row kernel with two rows per lane (fp32 up to 64), on the CTA-per-matrix rolled
    kernel (gj_roll2d_kernels.cuh, fp32 65 .. 128) or on the 2-D register-tile Gauss-Jordan kernel (gj_tile_kernels.cuh: fp64;
    fp32 n = 48 / 64 / 100 / 128 through INVGPU_GJ_KERNEL=tile below): flags are
    sgetrf's (first column without a non-zero pivot), a NaN column counts as singular, outputs of flagged
    matrices are NaN and the rest of the batch is unaffected."""
    rng = np.random.default_rng(n)
    a = rng.random((6, n, n)) + np.eye(n)
    a[1, :, n // 2] = 0.0            # zero column: pivot n/2 + 1 is exactly zero
    a[3] = 0.0                       # zero matrix: info 1
    a[4, :, n - 1] = np.nan          # NaN column: singular at the last step
    flat = orc.to_colmajor(a.astype(dtype))
    got, info = api.general_inverse_host(flat, n)
    want, oinfo = orc.gauss_jordan_inverse(flat, n)
    np.testing.assert_array_equal(info, oinfo)
    assert info[1] == n // 2 + 1 and info[3] == 1 and info[4] == n and (info != 0).sum() == 3
    good = info == 0
    assert_general_parity(a[good], orc.from_colmajor(got, n)[good], orc.from_colmajor(want, n)[good], dtype, f"flags n={n}")
    assert np.isnan(orc.from_colmajor(got, n)[~good]).all()
    want_tier = "warp-rowlane" if (dtype == np.float32 and n <= 64) else "cta-roll2d" if dtype == np.float32 else "gj-tile"
    assert api.tier_name("general", n, dtype).startswith(want_tier)


# --------------------------------------------------------------------------------------- GP mean / variance
@pytest.mark.parametrize("n", [8, 16, 32, 64])
@pytest.mark.parametrize("dtype", DTYPES)
def test_gp_reference_fixtures(api, fixtures_dir, n, dtype):
    g = {k: orc.to_colmajor(load_fixture(fixtures_dir, f"gaussian_100_{n}x{n}/{k}.mats", dtype))
         for k in ("a", "b", "c", "d", "e", "means", "variances")}
    keep = {k: v.copy() for k, v in g.items()}
    means, var, info = api.gp_host(n, g["a"], g["b"], g["c"], g["d"], g["e"])
    assert not info.any()
    for k in ("a", "b", "c", "d", "e"):                       # inputs are NOT destroyed (gauss_cpu.h:42 is)
        np.testing.assert_array_equal(g[k], keep[k])
    om, _ = orc.gp_mean(n, g["a"], g["b"], g["c"], g["d"])
    ov, _ = orc.gp_variance(n, g["a"], g["b"], g["c"], g["e"])
    tol = TOL[np.dtype(dtype)]
    assert np.abs(means - om).max() <= tol * max(1.0, np.abs(om).max())
    assert np.abs(var - ov).max() <= tol * max(1.0, np.abs(ov).max())
    assert np.abs(means - g["means"]).max() < 2e-4            # MATLAB goldens, 4 digits
    assert np.abs(var - g["variances"]).max() < 2e-4
    # mean-only and variance-only calls give the same numbers as the shared-factor call
    m2, _, _ = api.gp_host(n, g["a"], g["b"], g["c"], Ds=g["d"])
    _, v2, _ = api.gp_host(n, g["a"], g["b"], g["c"], Es=g["e"])
    np.testing.assert_array_equal(m2, means)
    np.testing.assert_array_equal(v2, var)


@pytest.mark.parametrize("n", [1, 3, 13, 33, 100, 128])
@pytest.mark.parametrize("dtype", DTYPES)
def test_gp_sizes(api, n, dtype):
    g = gp_batch(n, 37, dtype, seed=n)
    flat = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1) for k, v in g.items()}
    means, var, info = api.gp_host(n, flat["a"], flat["b"], flat["c"], flat["d"], flat["e"])
    assert not info.any()
    om, _ = orc.gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    ov, _ = orc.gp_variance(n, flat["a"], flat["b"], flat["c"], flat["e"])
    tol = TOL[np.dtype(dtype)]
    assert np.abs(means - om).max() <= tol
    assert np.abs(var - ov).max() <= tol


def test_gp_128_fixture_vectors_with_synthetic_b(api, fixtures_dir):
    """gaussian_100_128x128/b.mats is missing from the reference mount: use its a,c,d,e with a
    generated B (generate_gaussian_matrices.m:20-23) and check against the oracle."""
    n = 128
    g = {k: orc.to_colmajor(load_fixture(fixtures_dir, f"gaussian_100_128x128/{k}.mats", np.float32))
         for k in ("a", "c", "d", "e")}
    b = orc.to_colmajor(spd_batch(n, 100, np.float32, seed=128))
    means, var, info = api.gp_host(n, g["a"], b, g["c"], g["d"], g["e"])
    om, _ = orc.gp_mean(n, g["a"], b, g["c"], g["d"])
    ov, _ = orc.gp_variance(n, g["a"], b, g["c"], g["e"])
    assert not info.any()
    assert np.abs(means - om).max() <= 1e-4 and np.abs(var - ov).max() <= 1e-4


def test_gp_flags(api):
    n = 8
    g = gp_batch(n, 6, np.float64, seed=9)
    g["b"][3] = -g["b"][3]
    flat = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1) for k, v in g.items()}
    means, var, info = api.gp_host(n, flat["a"], flat["b"], flat["c"], flat["d"], flat["e"])
    _, oinfo = orc.gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    np.testing.assert_array_equal(info, oinfo)
    assert info[3] == 1 and info.sum() == 1


@pytest.mark.parametrize("n", [64, 128])
def test_gp_flags_cta_tiers(api, n):
    """fp32 n = 128 runs on the tcgen05 tier (natural pivot order), n = 64 on the tile kernel: flags must be LAPACK's."""
    g = gp_batch(n, 5, np.float32, seed=19)
    g["b"][1] = -g["b"][1]
    d = np.ones(n, dtype=np.float32)
    d[n // 3] = -5.0
    g["b"][3] = np.diag(d)                                        # with C >= 0 added the pivot n//3 stays negative
    g["c"][3] = 0.5
    flat = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1) for k, v in g.items()}
    means, var, info = api.gp_host(n, flat["a"], flat["b"], flat["c"], flat["d"], flat["e"])
    om, oinfo = orc.gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    np.testing.assert_array_equal(info, oinfo)
    assert info[1] == 1 and info[3] == n // 3 + 1 and (info != 0).sum() == 2
    good = info == 0
    assert np.abs(means[good] - om[good]).max() <= 1e-4
    assert np.isnan(means[~good]).all()


def test_gp_128_tensor_core_tier(api, torch):
    """fp32 n = 128 runs on the tcgen05 tier (tc_kernels.cuh: blocked Cholesky, 3xTF32 trailing update with the matrix as
    the TMEM accumulator).  Against the oracle and fp64 truth on (i) the reference generator's matrices, (ii) SPD matrices
    with cond ~1e3 (the 3xTF32 split must hold the 1e-4 bar there, a single TF32 pass does not), (iii) a batch that makes
    every CTA of the persistent grid loop (2 400 evaluations > 4 x 148), (iv) only the UPPER triangle may be read."""
    n = 128
    assert api.tier_name("gp", n, np.float32) == "tcgen05-blocked"
    rng = np.random.default_rng(5)
    # (ii) cond ~ 1e3: B = Q diag(s) Q^T
    q, _ = np.linalg.qr(rng.standard_normal((24, n, n)))
    sv = np.logspace(0, 3, n)
    b = ((q * sv) @ q.transpose(0, 2, 1))
    b = (0.5 * (b + b.transpose(0, 2, 1))).astype(np.float32)
    g = gp_batch(n, 24, np.float32, seed=3)
    g["b"] = b
    g["c"] = (0.01 * g["c"]).astype(np.float32)
    flat = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1) for k, v in g.items()}
    means, var, info = api.gp_host(n, flat["a"], flat["b"], flat["c"], flat["d"], flat["e"])
    assert not info.any()
    om, _ = orc.gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    ov, _ = orc.gp_variance(n, flat["a"], flat["b"], flat["c"], flat["e"])
    m64 = g["b"].astype(np.float64) + np.stack([np.diag(c) for c in g["c"].astype(np.float64)])
    tm = np.einsum("bi,bi->b", g["a"].astype(np.float64), np.linalg.solve(m64, g["d"].astype(np.float64)[..., None])[..., 0])
    scale = max(1.0, np.abs(tm).max())
    e_gpu, e_orc = np.abs(means - tm).max() / scale, np.abs(om - tm).max() / scale
    print(f"[tc gp128 cond 1e3] gpu {e_gpu:.2e} oracle {e_orc:.2e} (|mean| max {np.abs(tm).max():.2f})")
    assert e_gpu <= max(1e-4, 4 * e_orc)
    assert np.abs(var - ov).max() <= max(1e-4, 4 * e_orc) * max(1.0, np.abs(ov).max())
    # (iii) + (iv): many evaluations, lower triangle of B poisoned
    batch = 2400
    gen = torch.Generator(device="cuda").manual_seed(11)
    r = torch.rand((batch, n, n), generator=gen, device="cuda")
    bsym = r + r.transpose(1, 2) + n * torch.eye(n, device="cuda")
    a, c, d = (torch.rand((batch, n), generator=gen, device="cuda") for _ in range(3))
    e = torch.rand(batch, generator=gen, device="cuda")
    # column-major storage: element (row r, col c) at [c, r]; poison r > c (the lower triangle)
    colmajor = bsym.transpose(1, 2).contiguous()
    ci, ri = torch.arange(n, device="cuda").view(n, 1), torch.arange(n, device="cuda").view(1, n)
    colmajor_p = colmajor.clone()
    colmajor_p[:, ri > ci] = float("nan")                                  # [c, r] with r > c: strictly lower triangle
    assert colmajor_p.is_contiguous() and bool(torch.isnan(colmajor_p[0, 0, 1])) and not bool(torch.isnan(colmajor_p[0, 1, 0]))
    out_m, out_v = torch.zeros(batch, device="cuda"), torch.zeros(batch, device="cuda")
    d_info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    api.gp_device(n, a.data_ptr(), colmajor_p.data_ptr(), c.data_ptr(), d.data_ptr(), e.data_ptr(), out_m.data_ptr(), out_v.data_ptr(),
                  batch, np.float32, d_info.data_ptr(), st)
    torch.cuda.synchronize()
    assert int(d_info.abs().max()) == 0
    m64 = bsym.double() + torch.diag_embed(c.double())
    want_m = (a.double().unsqueeze(1) @ torch.linalg.solve(m64, d.double().unsqueeze(2))).reshape(-1)
    want_v = e.double() - (a.double().unsqueeze(1) @ torch.linalg.solve(m64, a.double().unsqueeze(2))).reshape(-1)
    assert float((out_m.double() - want_m).abs().max()) <= 1e-5
    assert float((out_v.double() - want_v).abs().max()) <= 1e-5


# --------------------------------------------------------------------------------------- legacy symbols
def test_legacy_host_entry_points(api, fixtures_dir):
    n = 16
    a = load_fixture(fixtures_dir, "inverse_100_16x16/a.mats", np.float32)
    flat = orc.to_colmajor(a)
    want, _ = orc.chol_inverse(flat, n)
    keep = flat.copy()
    for fn in (api.inverse_cholesky_batched_gpu, api.inverse_cholesky_mm_batched_gpu,
               api.inverse_cholesky_mm2_batched_gpu, api.inverse_cholesky_stride_batched_gpu,
               api.inverse_gauss_batched_gpu, api.inverse_lu_cuda_batched_gpu):
        got = fn(n, flat)
        assert normwise_err(orc.from_colmajor(got, n), orc.from_colmajor(want, n)) <= 1e-4
        np.testing.assert_array_equal(flat, keep)             # As is const (see include/inverse_gpu.h)
    g = {k: orc.to_colmajor(load_fixture(fixtures_dir, f"gaussian_100_16x16/{k}.mats", np.float32))
         for k in ("a", "b", "c", "d", "e", "means", "variances")}
    assert np.abs(api.calcluateMeanGPU(16, g["a"], g["b"], g["c"], g["d"]) - g["means"]).max() < 2e-4
    assert np.abs(api.calcluateVarianceGPU(16, g["a"], g["b"], g["c"], g["e"]) - g["variances"]).max() < 2e-4


@pytest.mark.parametrize("ptrs_on", ["pinned_host", "device"])
def test_legacy_device_entry_points(api, torch, ptrs_on):
    """`Array *devAs` flavour: pitched per-matrix device pointers like batchedCudaMalloc
    (reference src/helper.cu:103-118), pointer array in pinned host memory (as upstream) or on the device."""
    from cuda_matrix_inversion_b200 import lib
    n, batch, pitch = 12, 40, 640            # 12*12*4 = 576 bytes, rows padded to 640
    a = spd_batch(n, batch, np.float32, seed=77)
    d_a = torch.zeros(batch * pitch // 4, dtype=torch.float32, device="cuda")
    d_o = torch.zeros_like(d_a)
    d_a.view(batch, pitch // 4)[:, : n * n] = torch.from_numpy(orc.to_colmajor(a)).view(batch, n * n).cuda()
    pa = torch.tensor([d_a.data_ptr() + k * pitch for k in range(batch)], dtype=torch.int64)
    po = torch.tensor([d_o.data_ptr() + k * pitch for k in range(batch)], dtype=torch.int64)
    if ptrs_on == "pinned_host":
        pa, po = pa.pin_memory(), po.pin_memory()
    else:
        pa, po = pa.cuda(), po.cuda()

    def out():
        torch.cuda.synchronize()
        return orc.from_colmajor(d_o.view(batch, pitch // 4)[:, : n * n].contiguous().cpu().numpy().reshape(-1), n)

    want = np.linalg.inv(a.astype(np.float64))
    for name in ("inverse_cholesky_batched_device", "inverse_cholesky_mm_batched_device",
                 "inverse_cholesky_mm2_batched_device", "inverse_gauss_batched_device",
                 "inverse_lu_cuda_batched_device"):
        d_o.zero_()
        getattr(lib, name)(None, n, pa.data_ptr(), po.data_ptr(), batch)
        assert normwise_err(out(), want) <= 1e-4, name
    # inverse_lu_cuda_batched_device leaves the LU factors in devAs like upstream (cublasSgetrfBatched in place,
    # src/gauss/inverse_gpu.cu:24-33): restore the input for the calls below
    d_a.view(batch, pitch // 4)[:, : n * n] = torch.from_numpy(orc.to_colmajor(a)).view(batch, n * n).cuda()
    # stride family: in place on devAInvs, staged == fused
    d_o.copy_(d_a)
    lib.decompose_cholesky_stride_batched_device(None, n, pa.data_ptr(), po.data_ptr(), batch)
    l = out().astype(np.float64)
    assert normwise_err(l, np.linalg.cholesky(a.astype(np.float64))) <= 1e-4
    lib.inverse_upper_stride_batched_device(None, n, pa.data_ptr(), po.data_ptr(), batch)
    assert normwise_err(out(), np.linalg.inv(np.linalg.cholesky(a.astype(np.float64)))) <= 1e-4
    lib.multiply_upper_stride_batched_device(None, n, pa.data_ptr(), po.data_ptr(), batch)
    staged = out().copy()
    assert normwise_err(staged, want) <= 1e-4
    d_o.copy_(d_a)
    lib.inverse_cholesky_stride_batched_device(None, n, pa.data_ptr(), po.data_ptr(), batch)
    assert normwise_err(out(), want) <= 1e-4
    # decompose_cholesky_batched_device factors devAs in place
    d_o.copy_(d_a)
    lib.decompose_cholesky_batched_device(None, n, po.data_ptr(), pa.data_ptr(), batch)
    assert normwise_err(out(), np.linalg.cholesky(a.astype(np.float64))) <= 1e-4


# --------------------------------------------------------------------------------------- mixed dimensions
@pytest.mark.parametrize("dtype", DTYPES)
def test_mixed_dimension_scheduler(api, torch, dtype):
    """BASELINE config 5 in miniature: sizes drawn from the 32 / 128 / 256 buckets, one call."""
    rng = np.random.default_rng(777)
    top = 256 if dtype == np.float32 else 200
    ns = np.concatenate([rng.integers(1, 33, 150), rng.integers(33, 129, 40), rng.integers(129, top + 1, 10)])
    rng.shuffle(ns)
    mats = [spd_batch(int(n), 1, dtype, seed=int(1000 + i))[0] for i, n in enumerate(ns)]
    offs = np.concatenate([[0], np.cumsum([m.size for m in mats])])
    flat = np.concatenate([m.reshape(-1) for m in mats])
    bad = 17                                                   # one non-SPD matrix in the middle
    n_bad = int(ns[bad])
    flat[offs[bad]:offs[bad + 1]] = -np.eye(n_bad, dtype=dtype).reshape(-1)
    d_in = torch.from_numpy(flat).cuda()
    d_out = torch.zeros_like(d_in)
    d_info = torch.full((len(ns),), -1, dtype=torch.int32, device="cuda")
    esz = flat.itemsize
    pin = np.array([d_in.data_ptr() + int(o) * esz for o in offs[:-1]], dtype=np.uint64)
    pout = np.array([d_out.data_ptr() + int(o) * esz for o in offs[:-1]], dtype=np.uint64)
    api.mixed_spd_inverse_device(pin, pout, ns, dtype, d_info.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    info = d_info.cpu().numpy()
    out = d_out.cpu().numpy()
    assert info[bad] == 1 and (np.delete(info, bad) == 0).all()
    assert np.isnan(out[offs[bad]:offs[bad + 1]]).all()
    tol = TOL[np.dtype(dtype)]
    for i, n in enumerate(ns):
        if i == bad:
            continue
        n = int(n)
        want, _ = orc.chol_inverse(flat[offs[i]:offs[i + 1]], n)
        got = out[offs[i]:offs[i + 1]]
        assert np.abs(got - want).max() <= tol * np.abs(want).max(), (i, n)


def test_mixed_dimension_scheduler_threaded_planning(api, torch):
    """Batches of >= 2^16 matrices are planned by several host threads (per-thread histograms + scatter):
    every matrix must still land in the right tier and report into its own info slot."""
    rng = np.random.default_rng(99)
    cnt = 70_000
    ns = rng.integers(1, 41, cnt).astype(np.int32)              # tiers 16 / 24 / 32 / 48
    offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)])
    flat = np.zeros(int(offs[-1]), dtype=np.float32)
    diag_val = rng.uniform(1.0, 3.0, cnt).astype(np.float32)    # A_i = d_i * I  ->  inverse (1 / d_i) * I
    for i in np.nonzero(ns <= 40)[0]:
        n = int(ns[i])
        flat[offs[i]:offs[i + 1]:n + 1] = diag_val[i]
    bad = [5, 33_333, 69_999]
    for b in bad:
        flat[offs[b]] = -1.0                                     # first pivot negative: info 1
    d_in = torch.from_numpy(flat).cuda()
    d_out = torch.zeros_like(d_in)
    d_info = torch.full((cnt,), -1, dtype=torch.int32, device="cuda")
    pin = (d_in.data_ptr() + offs[:-1] * 4).astype(np.uint64)
    pout = (d_out.data_ptr() + offs[:-1] * 4).astype(np.uint64)
    api.mixed_spd_inverse_device(pin, pout, ns, np.float32, d_info.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    info = d_info.cpu().numpy()
    out = d_out.cpu().numpy()
    want_info = np.zeros(cnt, dtype=np.int32)
    want_info[bad] = 1
    np.testing.assert_array_equal(info, want_info)
    good = np.ones(cnt, dtype=bool)
    good[bad] = False
    d0 = out[offs[:-1]]                                          # element (0, 0) of every result
    np.testing.assert_allclose(d0[good], 1.0 / diag_val[good], rtol=1e-5)
    for i in (0, 1, 12_345, 69_998):                             # whole matrices of a few
        n = int(ns[i])
        np.testing.assert_allclose(out[offs[i]:offs[i + 1]].reshape(n, n), np.eye(n) / diag_val[i], rtol=1e-5, atol=1e-7)
    assert np.isnan(out[offs[5]:offs[6]]).all()


# --------------------------------------------------------------------------------------- host pipeline
def test_host_pipeline_chunking_and_pinned_path(api, torch, monkeypatch):
    """Many chunks, ragged last chunk, pageable and pinned user buffers give identical results."""
    n, batch = 32, 3001
    a = spd_batch(n, batch, np.float32, seed=31)
    flat = orc.to_colmajor(a)
    base, info0 = api.spd_inverse_host(flat, n)
    monkeypatch.setenv("INVGPU_CHUNK_MB", "1")
    small, info1 = api.spd_inverse_host(flat, n)
    np.testing.assert_array_equal(base, small)
    pin_in = torch.from_numpy(flat).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    got, info2 = api.spd_inverse_host(pin_in.numpy(), n, out=pin_out.numpy())
    np.testing.assert_array_equal(got, base)
    assert not info0.any() and not info1.any() and not info2.any()
    want, _ = orc.chol_inverse(flat[: 64 * n * n], n)
    assert normwise_err(orc.from_colmajor(base[: 64 * n * n], n), orc.from_colmajor(want, n)) <= 1e-4


# --------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("dtype", DTYPES)
def test_full_size_1m_32x32_properties(api, torch, dtype):
    """BASELINE config 3 (2^20 x 32x32): residual on a sample, inverse-of-inverse, scaling law."""
    n, batch = 32, 1 << 20
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    gen = torch.Generator(device="cuda").manual_seed(1234)
    r = torch.rand((batch, n, n), generator=gen, device="cuda", dtype=tdt)
    a = r + r.transpose(1, 2) + n * torch.eye(n, device="cuda", dtype=tdt)
    del r
    inv = torch.empty_like(a)
    info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    api.spd_inverse_device(a.data_ptr(), inv.data_ptr(), n, batch, dtype, info.data_ptr(), st)
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0
    tol = TOL[np.dtype(dtype)]
    idx = torch.randint(0, batch, (4096,), device="cuda")
    res = (a[idx].double() @ inv[idx].double() - torch.eye(n, device="cuda", dtype=torch.float64)).abs().sum(-1).max()
    assert float(res) <= tol
    # oracle on a slice of the very same device-generated inputs
    sl = a[:256].cpu().numpy()
    want, _ = orc.chol_inverse(orc.to_colmajor(sl), n)
    assert normwise_err(inv[:256].cpu().numpy(), orc.from_colmajor(want, n)) <= tol
    # inverse of the inverse returns A
    back = torch.empty_like(a)
    api.spd_inverse_device(inv.data_ptr(), back.data_ptr(), n, batch, dtype, info.data_ptr(), st)
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0
    rel = ((back - a).abs().amax(dim=(1, 2)) / a.abs().amax(dim=(1, 2))).max()
    assert float(rel) <= 10 * tol
    # inv(c A) = inv(A) / c, bit-for-bit for a power of two
    a.mul_(4.0)
    api.spd_inverse_device(a.data_ptr(), back.data_ptr(), n, batch, dtype, info.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(back * 4.0, inv)


def test_full_size_gp_25k_128(api, torch):
    """One GPU's shard of BASELINE config 4 (200k x 128x128 over 8 GPUs = 25k per GPU)."""
    n, batch = 128, 25000
    gen = torch.Generator(device="cuda").manual_seed(4321)
    r = torch.rand((batch, n, n), generator=gen, device="cuda")
    b = r + r.transpose(1, 2) + n * torch.eye(n, device="cuda")
    del r
    a, c, d = (torch.rand((batch, n), generator=gen, device="cuda") for _ in range(3))
    e = torch.rand((batch,), generator=gen, device="cuda")
    means = torch.empty(batch, device="cuda")
    var = torch.empty(batch, device="cuda")
    info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    api.gp_device(n, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), e.data_ptr(), means.data_ptr(),
                  var.data_ptr(), batch, np.float32, info.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0
    k = 128
    om, _ = orc.gp_mean(n, a[:k].cpu().numpy(), b[:k].cpu().numpy(), c[:k].cpu().numpy(), d[:k].cpu().numpy())
    ov, _ = orc.gp_variance(n, a[:k].cpu().numpy(), b[:k].cpu().numpy(), c[:k].cpu().numpy(), e[:k].cpu().numpy())
    assert np.abs(means[:k].cpu().numpy() - om).max() <= 1e-4
    assert np.abs(var[:k].cpu().numpy() - ov).max() <= 1e-4
    # linearity in D: mean(A,B,C,2D) = 2 mean(A,B,C,D), exactly
    d2 = d * 2
    m2 = torch.empty_like(means)
    api.gp_device(n, a.data_ptr(), b.data_ptr(), c.data_ptr(), d2.data_ptr(), 0, m2.data_ptr(), 0, batch, np.float32,
                  0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(m2, means * 2)
    # fp64 cross-check on a sample through torch (an independent implementation)
    idx = torch.arange(0, batch, 97, device="cuda")
    mm = b[idx].double() + torch.diag_embed(c[idx].double())
    x = torch.linalg.solve(mm, d[idx].double().unsqueeze(-1)).squeeze(-1)
    ref = (a[idx].double() * x).sum(-1)
    assert float((means[idx].double() - ref).abs().max()) <= 1e-4


# --------------------------------------------------------------------------------------- ragged batches
@pytest.mark.parametrize("batch", [1, 3, 31, 32, 33, 127, 129, 515])
def test_ragged_batches_bulk_copy_kernels(api, batch):
    """The thread-per-matrix / TMA / bulk-copy kernels move whole warp tiles; partial last tiles (and batches smaller
    than one tile) must arm their mbarriers with the right byte counts and never touch memory beyond the batch:
    guard bands around the outputs stay intact, every result matches the oracle."""
    rng = np.random.default_rng(batch)
    for n, dt in ((8, np.float64), (16, np.float64)):               # fp64: single-box TMA kernel, column-split GJ, GP thread kernel
        r = rng.random((batch, n, n))
        flat = orc.to_colmajor((r + r.transpose(0, 2, 1) + n * np.eye(n)).astype(dt))
        got, info = api.spd_inverse_host(flat, n)
        want, _ = orc.chol_inverse(flat, n)
        assert not info.any() and np.abs(got - want).max() <= 1e-10 * np.abs(want).max(), ("spd f64", n, batch)
        flat = orc.to_colmajor((rng.random((batch, n, n)) - 0.5 + 0.5 * n * np.eye(n)).astype(dt))
        got, info = api.general_inverse_host(flat, n)
        want, _ = orc.gauss_jordan_inverse(flat, n)
        assert not info.any() and np.abs(got - want).max() <= 1e-10 * np.abs(want).max(), ("general f64", n, batch)
        g = gp_batch(n, batch, dt, seed=batch + n)
        fl = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1) for k, v in g.items()}
        means, var, info = api.gp_host(n, fl["a"], fl["b"], fl["c"], fl["d"], fl["e"])
        om, _ = orc.gp_mean(n, fl["a"], fl["b"], fl["c"], fl["d"])
        assert not info.any() and np.abs(means - om).max() <= 1e-10, ("gp f64", n, batch)
    for n in (8, 16, 32):
        r = rng.random((batch, n, n))
        spd = (r + r.transpose(0, 2, 1) + n * np.eye(n)).astype(np.float32)
        flat = orc.to_colmajor(spd)
        got, info = api.spd_inverse_host(flat, n)
        want, _ = orc.chol_inverse(flat, n)
        assert not info.any()
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), ("spd", n, batch)
        gen = (rng.random((batch, n, n)) - 0.5 + 0.5 * n * np.eye(n)).astype(np.float32)   # diagonally dominant: cond = O(1)
        flat = orc.to_colmajor(gen)
        got, info = api.general_inverse_host(flat, n)
        want, _ = orc.gauss_jordan_inverse(flat, n)
        assert not info.any()
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), ("general", n, batch)
    for n in (8, 16):
        g = gp_batch(n, batch, np.float32, seed=batch)
        fl = {k: orc.to_colmajor(v) if v.ndim == 3 else v.reshape(-1) for k, v in g.items()}
        means, var, info = api.gp_host(n, fl["a"], fl["b"], fl["c"], fl["d"], fl["e"])
        om, _ = orc.gp_mean(n, fl["a"], fl["b"], fl["c"], fl["d"])
        ov, _ = orc.gp_variance(n, fl["a"], fl["b"], fl["c"], fl["e"])
        assert not info.any()
        assert np.abs(means - om).max() <= 1e-4 and np.abs(var - ov).max() <= 1e-4, ("gp", n, batch)


def test_device_outputs_have_intact_guard_bands(api, torch):
    """Device flavour: outputs embedded in a larger buffer filled with a sentinel; nothing outside may change."""
    for n, batch in ((8, 45), (16, 45), (32, 45), (32, 1)):
        rng = np.random.default_rng(n + batch)
        r = rng.random((batch, n, n))
        spd = torch.from_numpy((r + r.transpose(0, 2, 1) + n * np.eye(n)).astype(np.float32)).cuda()
        pad = 4096
        buf = torch.full((pad + batch * n * n + pad,), 777.0, dtype=torch.float32, device="cuda")
        out = buf[pad:pad + batch * n * n]
        info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        api.spd_inverse_device(spd.data_ptr(), out.data_ptr(), n, batch, np.float32, info.data_ptr(), st)
        api.general_inverse_device(spd.data_ptr(), out.data_ptr(), n, batch, np.float32, info.data_ptr(), st)
        torch.cuda.synchronize()
        assert bool((buf[:pad] == 777.0).all()) and bool((buf[pad + batch * n * n:] == 777.0).all()), (n, batch)
        want = torch.linalg.inv(spd.double().transpose(1, 2)).transpose(1, 2).float().reshape(-1)
        assert float((out - want).abs().max()) <= 1e-4 * float(want.abs().max())


# --------------------------------------------------------------------------------------- kernel variants
_VARIANT_SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import oracle as orc
from cuda_matrix_inversion_b200 import api
n, batch = {n}, {batch}
rng = np.random.default_rng(5)
r = rng.random((batch, n, n))
a = (r + r.transpose(0, 2, 1) + n * np.eye(n)).astype(np.float32)
a[3] = -np.eye(n)                                   # flagged: info 1
a[batch - 2, n - 1, n - 1] = -5.0                   # flagged in the last (partial) warp tile: info n
flat = orc.to_colmajor(a)
got, info = api.spd_inverse_host(flat, n)
want, oinfo = orc.chol_inverse(flat, n)
assert (info == oinfo).all(), (info[info != oinfo], oinfo[info != oinfo])
good = info == 0
g, w = orc.from_colmajor(got, n), orc.from_colmajor(want, n)
err = np.abs(g[good] - w[good]).max() / np.abs(w[good]).max()
assert err <= 1e-4, err
assert np.isnan(g[~good]).all() and (~good).sum() == 2
print("variant ok", err)
"""


@pytest.mark.parametrize("variant,n", [(0, 32), (6, 32), (7, 32), (9, 32), (9, 8), (9, 16), (4, 32), (4, 64), (4, 128), (1, 32), (3, 64), (3, 128)])
def test_sweep_kernel_variants(variant, n):
    """The non-default sweep configurations (n = 32: 0 = default = TMA tile I/O with interleaved lanes, 6 = TMA without
    interleaving, 7 = TMA load with prefetch + direct stores, 9 = direct global access; 2x2 block pivots: 4; other grids) are selected
    per process with INVGPU_SWEEP_VARIANT, so each runs in a child process: oracle parity, flags, ragged tail."""
    import subprocess
    import sys
    from cuda_matrix_inversion_b200 import lib
    if variant not in (0, 9) and not lib.invgpu_has_lab():
        pytest.skip("non-default kernel configuration: built only with `make lab=1`")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, INVGPU_SWEEP_VARIANT=str(variant), INVGPU_TRACE="1")
    batch = 1027 if n <= 32 else 131
    p = subprocess.run([sys.executable, "-c", _VARIANT_SNIPPET.format(root=root, n=n, batch=batch)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "variant ok" in p.stdout
    if n == 32 and variant in (0, 6, 7):
        assert "sweep_spd_tma_kernel" in p.stderr      # the TMA kernel really ran (no silent fallback)
    if variant == 9:
        assert "sweep_spd_tma_kernel" not in p.stderr  # the direct-access kernel


_GJ_VARIANT_SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import oracle as orc
from cuda_matrix_inversion_b200 import api
n, batch, dt = {n}, {batch}, np.{dtype}
rng = np.random.default_rng(7)
a = (rng.random((batch, n, n)) - 0.5 + n * 0.25 * np.eye(n)).astype(dt)
a[2, :, n // 2] = 0.0                                # zero column: info n/2 + 1
a[5] = 0.0                                           # zero matrix: info 1
a[batch - 1, :, n - 1] = np.nan                      # NaN column in the last (partial) warp tile: info n
flat = orc.to_colmajor(a)
got, info = api.general_inverse_host(flat, n)
want, oinfo = orc.gauss_jordan_inverse(flat, n)
assert (info == oinfo).all(), (info[info != oinfo], oinfo[info != oinfo])
assert info[2] == n // 2 + 1 and info[5] == 1 and info[batch - 1] == n and (info != 0).sum() == 3
good = info == 0
g, w = orc.from_colmajor(got, n), orc.from_colmajor(want, n)
err = np.abs(g[good] - w[good]).max() / np.abs(w[good]).max()
assert err <= {tol}, err
assert np.isnan(g[~good]).all()
print("variant ok", err)
"""


@pytest.mark.parametrize("kernel,n,dtype", [("colsplit", 16, "float32"), ("colsplit", 32, "float32"), ("rowlane", 16, "float64"),
                                            ("rowlane", 32, "float64"), ("rowlane", 8, "float32"), ("rowlane", 24, "float32"),
                                            ("generic", 32, "float32"), ("colsplit", 16, "float64"), ("tile", 32, "float64"), ("tile", 24, "float64"), ("tile", 64, "float32"), ("tile", 48, "float32"),
                                            ("tile", 128, "float32"), ("tile", 100, "float32")])
def test_general_kernel_variants(kernel, n, dtype):
    """The general-inverse tiers that are not the default for a shape (INVGPU_GJ_KERNEL = colsplit: column-split lanes
    rowlane: lane = row also at n = 8, tile: the 2-D tile kernel also at 17 <= n <= 32, generic: shared-memory tier) keep their parity tests:
    one child process each, oracle parity, sgetrf flags (zero column, zero matrix, NaN column), ragged tail."""
    import subprocess
    import sys
    from cuda_matrix_inversion_b200 import lib
    if kernel in ("colsplit", "rowlane") and not lib.invgpu_has_lab():
        pytest.skip("non-default kernel generation: built only with `make lab=1`")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, INVGPU_GJ_KERNEL=kernel)
    tol = 1e-4 if dtype == "float32" else 1e-10
    p = subprocess.run([sys.executable, "-c", _GJ_VARIANT_SNIPPET.format(root=root, n=n, batch=1027, dtype=dtype, tol=tol)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "variant ok" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("n", [128, 100])
def test_general_warp_specialised_variant(n):
    """INVGPU_GJR2_WS=1 (lab builds): the warp-specialised form of the CTA-per-matrix Gauss-Jordan kernel -- four FMA warps plus a
    pivot warp, registers handed over with setmaxnreg (gj_roll2d_ws_kernel).  Measured slower than the four-warp kernel, kept
    with the same parity test: oracle parity, sgetrf flags, ragged tail."""
    import subprocess
    import sys
    from cuda_matrix_inversion_b200 import lib
    if not lib.invgpu_has_lab():
        pytest.skip("non-default kernel generation: built only with `make lab=1`")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, INVGPU_GJR2_WS="1")
    p = subprocess.run([sys.executable, "-c", _GJ_VARIANT_SNIPPET.format(root=root, n=n, batch=331, dtype="float32", tol=1e-4)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "variant ok" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [8, 16, 24, 32, 64, 100, 128, 160])
def test_gp_reads_upper_triangle_only(api, n, dtype):
    """Every GP tier (thread-per-evaluation, tile, sweep, tcgen05, shared-memory generic) reads only the UPPER triangle of B
    (rows 0 .. c of column c; the reference's CPU code does the same: potrf 'U', src/gauss_cpu.c:54): a batch whose strictly
    lower triangle is NaN gives bit-identical means and variances.  This is what lets the host call send column prefixes."""
    import torch
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    batch = 300
    gen = torch.Generator(device="cuda").manual_seed(n)
    r = torch.rand((batch, n, n), generator=gen, device="cuda", dtype=tdt)
    colmajor = (r + r.transpose(1, 2) + n * torch.eye(n, device="cuda", dtype=tdt)).contiguous()    # symmetric: [c, r] == [r, c]
    a, c, d = (torch.rand((batch, n), generator=gen, device="cuda", dtype=tdt) for _ in range(3))
    e = torch.rand(batch, generator=gen, device="cuda", dtype=tdt)
    ci, ri = torch.arange(n, device="cuda").view(n, 1), torch.arange(n, device="cuda").view(1, n)
    poisoned = colmajor.clone()
    poisoned[:, ri > ci] = float("nan")                            # [c, r] with r > c: strictly lower triangle
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for bmat in (colmajor, poisoned):
        m, v = torch.zeros(batch, device="cuda", dtype=tdt), torch.zeros(batch, device="cuda", dtype=tdt)
        info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
        api.gp_device(n, a.data_ptr(), bmat.data_ptr(), c.data_ptr(), d.data_ptr(), e.data_ptr(), m.data_ptr(), v.data_ptr(),
                      batch, dtype, info.data_ptr(), st)
        torch.cuda.synchronize()
        assert int(info.abs().max()) == 0
        outs.append((m.cpu().numpy(), v.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), api.tier_name("gp", n, dtype)
    assert np.isfinite(outs[1][0]).all() and np.isfinite(outs[1][1]).all()

