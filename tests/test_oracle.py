"""Pins the CPU oracle (oracle/oracle.c) before anything trusts it.

1. against every golden vector the reference ships for the path
   (SURVEY.md 8c): inverse_100_{8,16,32}/aInv.mats, gaussian_100_*/{means,variances}.mats,
   simpleMean/chol.mats -> cholinv.mats;
2. against the reference's own CPU sources compiled unmodified (oracle/_ref),
   when that library is present;
3. fp64 instantiation against numpy/LAPACK fp64 (the reference has no fp64).
"""
import os

import numpy as np
import pytest

import oracle as orc

SPD_SIZES = [8, 16, 32, 64]
GP_SIZES = [8, 16, 32, 64]


def _load_inverse(fixtures_dir, n, dtype):
    a = orc.read_mats(os.path.join(fixtures_dir, f"inverse_100_{n}x{n}", "a.mats"), dtype)
    p = os.path.join(fixtures_dir, f"inverse_100_{n}x{n}", "aInv.mats")
    ainv = orc.read_mats(p, dtype) if os.path.exists(p) else None
    return a, ainv


def _load_gp(fixtures_dir, n, dtype):
    d = os.path.join(fixtures_dir, f"gaussian_100_{n}x{n}")
    return {k: orc.read_mats(os.path.join(d, f"{k}.mats"), dtype)
            for k in ("a", "b", "c", "d", "e", "means", "variances")}


def test_mats_reader_matches_c_reader(fixtures_dir):
    for rel in ("square_5_8_8.mats", "gaussian_100_8x8/b.mats", "simpleMean/b.mats", "gaussian_100_8x8/c.mats"):
        path = os.path.join(fixtures_dir, rel)
        a = orc.read_mats(path)
        flat = orc.read_mats_c(path)
        np.testing.assert_array_equal(orc.to_colmajor(a), flat)


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built")
def test_mats_reader_matches_reference_reader(fixtures_dir):
    import ctypes as C
    io = orc.ref_io()
    if io is None:
        pytest.skip("libref_io.so not built")
    path = os.path.join(fixtures_dir, "square_5_16_16.mats")
    k, m, n = C.c_int(), C.c_int(), C.c_int()
    ptr = C.POINTER(C.c_float)()
    io.readMatricesFile(path.encode(), C.byref(k), C.byref(m), C.byref(n), C.byref(ptr))
    got = np.ctypeslib.as_array(ptr, shape=(k.value * m.value * n.value,)).copy()
    want = orc.to_colmajor(orc.read_mats(path, np.float32))
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n", [8, 16, 32])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_chol_inverse_vs_golden_ainv(fixtures_dir, n, dtype):
    a, ainv = _load_inverse(fixtures_dir, n, dtype)
    got, info = orc.chol_inverse(orc.to_colmajor(a), n)
    assert not info.any()
    got = orc.from_colmajor(got, n)
    # goldens were computed from untruncated inputs and written with 4 digits:
    # they pin results to ~1e-4 absolute.
    assert np.abs(got - ainv).max() < 2e-4
    assert np.abs(got - ainv).sum(axis=(1, 2)).mean() < 1e-3 * n / 8


def test_chol_inverse_known_answer_simplemean(fixtures_dir):
    a = orc.read_mats(os.path.join(fixtures_dir, "simpleMean", "chol.mats"), np.float64)
    want = orc.read_mats(os.path.join(fixtures_dir, "simpleMean", "cholinv.mats"), np.float64)
    got, info = orc.chol_inverse(orc.to_colmajor(a), 4)
    assert info[0] == 0
    got = orc.from_colmajor(got, 4)[0]
    exact = np.array([[2.515625, 0.484375, -1.296875, 0.359375],
                      [0.484375, 0.140625, -0.328125, 0.140625],
                      [-1.296875, -0.328125, 1.015625, -0.578125],
                      [0.359375, 0.140625, -0.578125, 0.515625]])
    np.testing.assert_allclose(got, exact, rtol=0, atol=1e-11)
    np.testing.assert_allclose(got, want[0], rtol=0, atol=1e-5)   # file holds an fp32 run
    got32, _ = orc.chol_inverse(orc.to_colmajor(a.astype(np.float32)), 4)
    np.testing.assert_allclose(orc.from_colmajor(got32, 4)[0], exact, rtol=0, atol=2e-4)


@pytest.mark.parametrize("n", SPD_SIZES)
def test_chol_inverse_fp64_vs_numpy(fixtures_dir, n):
    a, _ = _load_inverse(fixtures_dir, n, np.float64)
    got, info = orc.chol_inverse(orc.to_colmajor(a), n)
    assert not info.any()
    got = orc.from_colmajor(got, n)
    want = np.linalg.inv(a)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max() * n
    assert np.abs(got - got.transpose(0, 2, 1)).max() == 0      # both triangles written


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", SPD_SIZES)
def test_chol_inverse_vs_reference_cpu(fixtures_dir, n):
    a, _ = _load_inverse(fixtures_dir, n, np.float32)
    flat = orc.to_colmajor(a)
    ref = orc.from_colmajor(orc.ref_chol_inverse_upper(flat, n), n)
    got, _ = orc.chol_inverse(flat, n, full=False)
    got = orc.from_colmajor(got, n)
    iu = np.triu_indices(n)
    scale = np.abs(ref[:, iu[0], iu[1]]).max()
    assert np.abs(got[:, iu[0], iu[1]] - ref[:, iu[0], iu[1]]).max() <= 1e-5 * scale
    # strictly-lower triangle is the untouched input in both (spotri_("U"))
    il = np.tril_indices(n, -1)
    np.testing.assert_array_equal(got[:, il[0], il[1]], a[:, il[0], il[1]])
    np.testing.assert_array_equal(ref[:, il[0], il[1]], a[:, il[0], il[1]])


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128])
@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-4), (np.float64, 1e-10)])
@pytest.mark.parametrize("algo", ["gauss_jordan", "lu"])
def test_general_inverse_residual(fixtures_dir, n, dtype, tol, algo):
    a = orc.read_mats(os.path.join(fixtures_dir, f"square_5_{n}_{n}.mats"), dtype)
    fn = orc.gauss_jordan_inverse if algo == "gauss_jordan" else orc.lu_inverse
    got, info = fn(orc.to_colmajor(a), n)
    assert not info.any()
    got = orc.from_colmajor(got, n).astype(np.float64)
    want = np.linalg.inv(a.astype(np.float64))
    # normwise tolerance scaled by the conditioning (cond up to 9e3 for these fixtures)
    cond = np.linalg.cond(a.astype(np.float64))
    eps = np.finfo(dtype).eps
    err = np.abs(got - want).max(axis=(1, 2)) / np.abs(want).max(axis=(1, 2))
    assert (err <= 8 * eps * cond * np.sqrt(n)).all()
    if dtype == np.float64:
        res = np.abs(a @ got - np.eye(n)).sum(axis=2).max(axis=1)
        assert (res <= tol).all()


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [8, 16, 32, 64])
def test_lu_inverse_vs_reference_cpu(fixtures_dir, n):
    a = orc.read_mats(os.path.join(fixtures_dir, f"square_5_{n}_{n}.mats"), np.float32)
    flat = orc.to_colmajor(a)
    ref = orc.ref_lu_inverse(flat, n)
    got, info = orc.lu_inverse(flat, n)
    assert not info.any()
    scale = np.abs(ref).max()
    cond = np.linalg.cond(a.astype(np.float64)).max()
    assert np.abs(got - ref).max() <= 4 * np.finfo(np.float32).eps * cond * scale


@pytest.mark.parametrize("n", GP_SIZES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gp_mean_variance_vs_golden(fixtures_dir, n, dtype):
    g = _load_gp(fixtures_dir, n, dtype)
    flat = {k: orc.to_colmajor(v) for k, v in g.items()}
    mean, info = orc.gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    assert not info.any()
    var, info = orc.gp_variance(n, flat["a"], flat["b"], flat["c"], flat["e"])
    assert not info.any()
    # 4-digit fixtures (inputs AND goldens truncated) pin to ~1e-4 absolute
    assert np.abs(mean - flat["means"]).max() < 2e-4
    assert np.abs(var - flat["variances"]).max() < 2e-4
    msolve, _ = orc.gp_mean_solve(n, flat["a"], flat["b"], flat["c"], flat["d"])
    assert np.abs(msolve - mean).max() < (1e-5 if dtype == np.float32 else 1e-13)


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", GP_SIZES)
def test_gp_mean_vs_reference_cpu(fixtures_dir, n):
    g = _load_gp(fixtures_dir, n, np.float32)
    flat = {k: orc.to_colmajor(v) for k, v in g.items()}
    ref = orc.ref_gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    got, _ = orc.gp_mean(n, flat["a"], flat["b"], flat["c"], flat["d"])
    assert np.abs(got - ref).max() <= 2e-6
    # the shipped CPU variance has the wrong sign (App. A-1): E + q instead of E - q
    raw = orc.ref_gp_variance_raw(n, flat["a"], flat["b"], flat["c"], flat["e"])
    var, _ = orc.gp_variance(n, flat["a"], flat["b"], flat["c"], flat["e"])
    assert np.abs((2 * flat["e"] - raw) - var).max() <= 2e-6


def test_flags_match_lapack_semantics(fixtures_dir):
    # singular 2x2 from the reference's own fixtures (all 1.1): spotrf info = 2, getrf info = 2
    b = orc.read_mats(os.path.join(fixtures_dir, "simpleMean", "b.mats"), np.float64)
    assert b.shape == (1, 2, 2)
    _, info = orc.chol_inverse(orc.to_colmajor(b), 2)
    assert info[0] == 2
    _, info = orc.gauss_jordan_inverse(orc.to_colmajor(b), 2)
    assert info[0] == 2
    _, info = orc.lu_inverse(orc.to_colmajor(b), 2)
    assert info[0] == 2
    # indefinite: leading minor of order 3 fails
    a = np.diag([4.0, 3.0, -1.0, 2.0])[None]
    for dt in (np.float32, np.float64):
        _, info = orc.chol_inverse(orc.to_colmajor(a.astype(dt)), 4)
        assert info[0] == 3
    # NaN pivot is flagged at its own position
    a = np.eye(5)[None].copy(); a[0, 1, 1] = np.nan
    _, info = orc.chol_inverse(orc.to_colmajor(a), 5)
    assert info[0] == 2
    import scipy.linalg as sl
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(2, 12))
        r = rng.standard_normal((n, n)); s = r + r.T
        s += np.eye(n) * rng.uniform(-1, 3)
        _, info = orc.chol_inverse(orc.to_colmajor(s[None]), n)
        _, li = sl.lapack.dpotrf(s, lower=0)
        assert info[0] == li


# --------------------------------------------------------------------------------------- LU factors / solve (section 8 f4)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 5, 16, 33, 64])
def test_getrf_matches_lapack_pivots_and_factors(n, dtype):
    """orc_getrf (sgetf2 restated) against LAPACK getrf through scipy: pivot indices bit-exact, factors to rounding,
    and getrs against numpy's fp64 solve -- this is what pins the GPU getrf / gesv tests."""
    import scipy.linalg as sl
    rng = np.random.default_rng(n)
    a = rng.random((9, n, n)).astype(dtype)
    lu, ipiv, info = orc.getrf(orc.to_colmajor(a), n)
    assert not info.any() and ipiv.min() >= 1 and ipiv.max() <= n
    lu3 = orc.from_colmajor(lu, n)
    tol = 2e-4 if dtype == np.float32 else 1e-11
    for k in range(9):
        lu_s, piv_s = sl.lu_factor(a[k])
        np.testing.assert_array_equal(piv_s + 1, ipiv[k])
        assert np.abs(lu3[k] - lu_s).max() <= tol * max(1.0, np.abs(lu_s).max())
    b = rng.random((9, 3, n)).astype(dtype)                       # three right-hand sides, column-major n x 3
    x = orc.getrs(lu, ipiv, b.reshape(-1), n, 3).reshape(9, 3, n)
    want = np.linalg.solve(a.astype(np.float64), b.transpose(0, 2, 1).astype(np.float64)).transpose(0, 2, 1)
    cond = np.linalg.cond(a.astype(np.float64)).max()
    assert np.abs(x - want).max() <= 8 * np.finfo(dtype).eps * cond * max(1.0, np.abs(want).max())


def test_getrf_zero_pivot_info_and_completion():
    """sgetf2 semantics: an exactly-zero pivot column is recorded (info = first k) and the factorisation goes on."""
    import scipy.linalg as sl
    a = np.array([[[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [1.0, 1.0, 1.0]]])     # rank 2, a zero pivot appears at step 3
    lu, ipiv, info = orc.getrf(orc.to_colmajor(a), 3)
    lu_s, piv_s = sl.lu_factor(a[0], check_finite=False)
    np.testing.assert_array_equal(piv_s + 1, ipiv[0])
    assert info[0] == 3
    np.testing.assert_allclose(orc.from_colmajor(lu, 3)[0], lu_s, atol=1e-15)
