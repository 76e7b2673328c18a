"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
headers declare, and the host-only helpers (.mats I/O) behave like the reference's.
No compute entry point is called here (there is no GPU in this container)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def _declared_symbols():
    names = set()
    src = open(os.path.join(INCLUDE, "inverse_gpu.h")).read()
    names |= set(re.findall(r"INVGPU_(?:HOST|DEVICE)_ENTRY\((\w+)\);", src))
    for hdr in ("gauss_gpu.h", "helper_cpu.h", "invgpu.h"):
        src = open(os.path.join(INCLUDE, hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"^[ \t]*#[ \t]*define(?:\\\n|[^\n])*", "", src, flags=re.M)     # macro bodies (fail / ensure / div_ceil)
        names |= set(re.findall(r"\b(\w+)\s*\([^;{]*\)\s*;", src))
    # helper_gpu.h needs <cuda_runtime.h>: its one exported function is listed by name
    assert "batchedCudaMalloc(" in open(os.path.join(INCLUDE, "helper_gpu.h")).read()
    names.add("batchedCudaMalloc")
    return names - {"defined"}


def test_library_exports_every_declared_symbol():
    from cuda_matrix_inversion_b200 import LIB_PATH
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    declared = _declared_symbols()
    assert len(declared) >= 17 + 4 + 5 + 20
    missing = sorted(declared - exported)
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # the 17 reference prototypes (reference include/inverse_gpu.h:7-31), by name
    ref17 = """inverse_gauss_batched_gpu inverse_lu_cuda_batched_gpu inverse_gauss_batched_device
        inverse_lu_cuda_batched_device inverse_cholesky_stride_batched_gpu inverse_cholesky_stride_batched_device
        decompose_cholesky_stride_batched_device inverse_upper_stride_batched_device
        multiply_upper_stride_batched_device inverse_cholesky_batched_device decompose_cholesky_batched_device
        inverse_cholesky_mm_batched_device decompose_cholesky_mm_batched_device inverse_cholesky_batched_gpu
        inverse_cholesky_mm_batched_gpu inverse_cholesky_mm2_batched_device inverse_cholesky_mm2_batched_gpu""".split()
    assert len(ref17) == 17 and not (set(ref17) - exported)


def test_headers_compile_as_c_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "types.h"\n#include "inverse_gpu.h"\n#include "gauss_gpu.h"\n#include "helper_cpu.h"\n'
                   '#include "invgpu.h"\nint main(void){return 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INCLUDE, "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)
    subprocess.run(["g++", "-x", "c++", "-Wall", "-Werror", "-I", INCLUDE, "-c", str(src), "-o", str(tmp_path / "t2.o")], check=True)
    # the reference's include order (src/gauss_bench.cu:6-18): cuda_runtime.h, cublas_v2.h, types.h, helper_cpu.h, helper_gpu.h
    cuda_inc = "/usr/local/cuda/include"
    if os.path.exists(os.path.join(cuda_inc, "cublas_v2.h")):
        src2 = tmp_path / "u.c"
        src2.write_text('#include <stdio.h>\n#include <stdlib.h>\n#include <errno.h>\n#include <cuda_runtime.h>\n#include "cublas_v2.h"\n'
                        '#include "types.h"\n#include "helper_cpu.h"\n#include "helper_gpu.h"\n#include "inverse_gpu.h"\n'
                        'int main(void){ size_t p; Array d[2]; gpuErrchk(batchedCudaMalloc(d, &p, 64, 2)); cublasErrchk(CUBLAS_STATUS_SUCCESS);\n'
                        ' ensure(div_ceil(5, 2) == 3, "div_ceil"); if (p == 1) { fail("unreachable %d", 1); } return 0; }\n')
        for cc, extra in (("gcc", ["-std=gnu99"]), ("g++", ["-x", "c++"])):
            subprocess.run([cc, *extra, "-Wall", "-Werror", "-Wno-unused-function", "-I", INCLUDE, "-I", cuda_inc, "-c", str(src2),
                            "-o", str(tmp_path / "u.o")], check=True)


def test_ctypes_mirror_binds_everything():
    from cuda_matrix_inversion_b200 import lib
    assert len(lib._declared) >= 50
    assert lib.invgpu_version().startswith(b"invgpu")


def test_no_device_means_loud_failure():
    """Without a GPU the compute entry points return an error -- they never fall back to the CPU."""
    from cuda_matrix_inversion_b200 import api
    if api.device_count() > 0:
        pytest.skip("GPU present")
    a = np.eye(4, dtype=np.float32).reshape(-1)
    with pytest.raises(api.InvGpuError):
        api.spd_inverse_host(a, 4)


def test_read_mats_file_matches_oracle_reader(fixtures_dir, tmp_path):
    from cuda_matrix_inversion_b200 import api, lib
    for rel in ("square_5_8_8.mats", "gaussian_100_8x8/b.mats", "gaussian_100_8x8/c.mats", "simpleMean/b.mats"):
        path = os.path.join(fixtures_dir, rel)
        flat, shape = api.read_mats_file(path)
        want = orc.read_mats(path, np.float32)
        assert shape == want.shape
        np.testing.assert_array_equal(flat, orc.to_colmajor(want))
    # write -> read round trip
    a = orc.to_colmajor(orc.read_mats(os.path.join(fixtures_dir, "square_5_8_8.mats"), np.float32))
    p = str(tmp_path / "rt.mats")
    assert lib.writeMatricesFile(p.encode(), 5, 8, 8, a.ctypes.data_as(C.c_void_p), 0) == 0
    back, shape = api.read_mats_file(p)
    assert shape == (5, 8, 8)
    np.testing.assert_array_equal(back, a)


def test_replicate_matrices(fixtures_dir):
    from cuda_matrix_inversion_b200 import lib
    path = os.path.join(fixtures_dir, "simpleMean", "chol.mats").encode()
    k, m, n = C.c_int(), C.c_int(), C.c_int()
    ptr = C.c_void_p()
    lib.readMatricesFile(path, C.byref(k), C.byref(m), C.byref(n), C.byref(ptr))
    lib.replicateMatrices(C.byref(ptr), 4, 4, 1, 3)
    got = np.frombuffer((C.c_float * 48).from_address(ptr.value), dtype=np.float32)
    np.testing.assert_array_equal(got[:16], got[16:32])
    np.testing.assert_array_equal(got[:16], got[32:])
    assert got[1] == 22 and got[4] == 22 and got[0] == 18
