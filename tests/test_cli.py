"""The reference-compatible CLIs (bin/inverse_bench, bin/gauss_bench): argv contract and stdout
formats of SURVEY.md Appendix C (reference src/inverse_bench.c:54-71, 276-303; src/gauss_bench.cu:504-529,
577-702)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
FIX = os.path.join(ROOT, "tests", "golden", "reference")
FLT = r"[-+]?\d\.\d+e[-+]\d+"


def _run(*args, **kw):
    return subprocess.run(list(args), capture_output=True, text=True, cwd=ROOT, **kw)


def _have_bins():
    return os.path.exists(os.path.join(BIN, "inverse_bench")) and os.path.exists(os.path.join(BIN, "gauss_bench"))


@pytest.mark.skipif(not _have_bins(), reason="CLIs not built (make cli)")
def test_usage_and_no_gpu_is_loud():
    r = _run(os.path.join(BIN, "inverse_bench"))
    assert r.returncode != 0 and "Usage: inverse_bench TEST_FOLDER TEST_REPLICATIONS MATRIX_DUPLICATES [-csv]" in r.stderr
    r = _run(os.path.join(BIN, "gauss_bench"))
    assert r.returncode != 0 and "Usage: gauss_bench TEST_FOLDER TEST_REPLICATIONS MATRIX_DUPLICATES [-csv]" in r.stderr
    import cuda_matrix_inversion_b200 as pkg
    if pkg.api.device_count() == 0:      # no device: refuse, never compute on the CPU
        r = _run(os.path.join(BIN, "inverse_bench"), os.path.join(FIX, "inverse_100_8x8"), "1", "1")
        assert r.returncode != 0 and "no CPU path" in r.stderr and r.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("csv", [True, False])
def test_inverse_bench_output(csv):
    args = [os.path.join(BIN, "inverse_bench"), os.path.join(FIX, "inverse_100_16x16"), "3", "2"] + (["-csv"] if csv else [])
    r = _run(*args)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    names = ["chol_gpu", "chol_mm2_gpu", "gauss_batched_gpu", "lu_cuda_batched_gpu"]
    assert len(lines) == 4
    for line, name in zip(lines, names):
        if csv:
            m = re.fullmatch(rf"200 16 3 {name} ({FLT}) ({FLT}) ({FLT}) ({FLT})", line)
            assert m, line
            assert float(m.group(4)) < 5e-3          # avg L1 error vs the 4-digit MATLAB goldens
        else:
            assert re.fullmatch(rf"{name} - 200 16x16 matrices, replicated 3 times, runtime \d+\.\d{{4}} ms "
                                rf"\(\d+\.\d{{4}} ms average, \d+\.\d{{4}} ms variance\), average error {FLT}", line), line
    # single repetition: short form
    r = _run(os.path.join(BIN, "inverse_bench"), os.path.join(FIX, "inverse_100_8x8"), "1", "1", "-csv")
    assert r.returncode == 0
    assert re.fullmatch(rf"100 8 1 chol_gpu {FLT} {FLT}", r.stdout.splitlines()[0])


@pytest.mark.gpu
def test_gauss_bench_output_and_error_column():
    r = _run(os.path.join(BIN, "gauss_bench"), os.path.join(FIX, "gaussian_100_64x64"), "2", "16", "-csv", "--json")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    m = re.fullmatch(rf"1600 64 2 means_gpu ({FLT}) ({FLT}) ({FLT}) ({FLT})", lines[0])
    v = re.fullmatch(rf"1600 64 2 variances_gpu ({FLT}) ({FLT}) ({FLT}) ({FLT})", lines[1])
    assert m and v, lines
    assert float(m.group(4)) < 1e-4 and float(v.group(4)) < 1e-4     # mean |out - golden| per evaluation
    assert lines[2].startswith('{"bench": "gauss_bench"')


@pytest.mark.gpu
def test_singular_input_aborts_like_reference(tmp_path):
    d = tmp_path / "sing"
    d.mkdir()
    (d / "a.mats").write_text("1 2 2\n1.1 1.1\n1.1 1.1\n")          # reference tests/simpleMean/b.mats
    (d / "aInv.mats").write_text("1 2 2\n0 0\n0 0\n")
    r = _run(os.path.join(BIN, "inverse_bench"), str(d), "1", "1")
    assert r.returncode != 0
    assert "Error code 2 in cholesky factorization" in r.stderr
