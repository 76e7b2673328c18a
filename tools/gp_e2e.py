#!/usr/bin/env python
"""End-to-end rate of invgpu_gp_host_f32 at n = 128:  tools/gp_e2e.py [batch] [--pageable]   (default: pinned host buffers)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda_matrix_inversion_b200 import lib

n = 128
pageable = '--pageable' in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith('--')]
gb = int(args[0]) if args else 100000
hB = torch.empty(gb * n * n, dtype=torch.float32)
if not pageable:
    hB = hB.pin_memory()
r = torch.rand((1000, n, n))
blk = (r + r.transpose(1, 2) + n * torch.eye(n)).reshape(-1)
for i in range(0, gb, 1000):
    hB[i * n * n:(i + 1000) * n * n] = blk[: min(1000, gb - i) * n * n]
hA, hC, hD = (torch.rand(gb * n).pin_memory() for _ in range(3))
means = np.zeros(gb, dtype=np.float32); info = np.zeros(gb, dtype=np.int32)
def step():
    rc = lib.invgpu_gp_host_f32(n, hA.data_ptr(), hB.data_ptr(), hC.data_ptr(), hD.data_ptr(), None, means.ctypes.data, None, gb, info.ctypes.data)
    assert rc == 0, rc
step()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
print(f"{'pageable' if pageable else 'pinned'} B, W={os.environ.get('INVGPU_GP_UPPER_W', 'default')} upper={lib.invgpu_gp_upper_h2d(n, 4)} chunk={os.environ.get('INVGPU_CHUNK_MB', '32')}MB: {gb / min(ts):.4e} eval/s  ({min(ts) * 1e3:.1f} ms, info max {int(np.abs(info).max())}, checksum {float(means.sum()):.6e})")
