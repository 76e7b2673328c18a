#!/usr/bin/env python
"""End-to-end rate of invgpu_spd_inverse_host_f32 (pinned host buffers):  tools/spd_e2e.py [n] [batch]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda_matrix_inversion_b200 import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 20) * 32 * 32 // (n * n)
hin = torch.empty(batch * n * n, dtype=torch.float32).pin_memory()
hout = torch.empty(batch * n * n, dtype=torch.float32).pin_memory()
r = torch.rand((1024, n, n))
blk = (r + r.transpose(1, 2) + n * torch.eye(n)).reshape(-1)
for i in range(0, batch, 1024):
    k = min(1024, batch - i)
    hin[i * n * n:(i + k) * n * n] = blk[: k * n * n]
info = np.zeros(batch, dtype=np.int32)
def step():
    rc = lib.invgpu_spd_inverse_host_f32(hin.data_ptr(), hout.data_ptr(), n, batch, info.ctypes.data)
    assert rc == 0, rc
step()
ts = []
for _ in range(4):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
ref = torch.linalg.inv((r[:4] + r[:4].transpose(1, 2) + n * torch.eye(n)).double())
got = hout[: 4 * n * n].reshape(4, n, n).double()
err = float((got - ref).abs().max() / ref.abs().max())
print(f"n={n} batch={batch} {batch / min(ts):.4e} inv/s ({min(ts) * 1e3:.1f} ms; {2 * batch * n * n * 4 / min(ts) / 1e9:.1f} GB/s of whole matrices both ways; err {err:.1e}, info max {int(np.abs(info).max())})")
