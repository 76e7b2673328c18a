"""Shared helpers for the parity tests (synthetic inputs follow the reference's MATLAB
generators, tests/generate_inverse_matrices.m:8-21 and tests/generate_gaussian_matrices.m:14-37)."""
import os

import numpy as np

import oracle as orc

TOL = {np.dtype(np.float32): 1e-4, np.dtype(np.float64): 1e-10}   # BASELINE.json north_star tolerances


def spd_batch(n, batch, dtype=np.float32, seed=1234):
    """A = R + R^T + n*I, R ~ U(0,1): math-indexed [batch, n, n] (symmetric, so layout-agnostic)."""
    rng = np.random.default_rng(seed)
    r = rng.random((batch, n, n))
    a = r + r.transpose(0, 2, 1) + n * np.eye(n)
    return a.astype(dtype)


def general_batch(n, batch, dtype=np.float32, seed=20260101):
    rng = np.random.default_rng(seed)
    return rng.random((batch, n, n)).astype(dtype)


def gp_batch(n, batch, dtype=np.float32, seed=4321):
    rng = np.random.default_rng(seed)
    b = spd_batch(n, batch, np.float64, seed + 1)
    g = dict(a=rng.random((batch, n)), b=b, c=rng.random((batch, n)), d=rng.random((batch, n)),
             e=rng.random(batch))
    return {k: v.astype(dtype) for k, v in g.items()}


def load_fixture(fixtures_dir, rel, dtype):
    return orc.read_mats(os.path.join(fixtures_dir, rel), dtype)


def residual_inf(a, ainv):
    """max over the batch of || A * Ainv - I ||_inf, evaluated in fp64."""
    a = np.asarray(a, dtype=np.float64)
    x = np.asarray(ainv, dtype=np.float64)
    n = a.shape[-1]
    r = a @ x - np.eye(n)
    return np.abs(r).sum(axis=-1).max()


def normwise_err(got, want):
    """max over the batch of max|got - want| / max|want| ('tolerance wording')."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    num = np.abs(got - want).reshape(got.shape[0], -1).max(axis=1)
    den = np.abs(want).reshape(want.shape[0], -1).max(axis=1)
    return (num / den).max()
