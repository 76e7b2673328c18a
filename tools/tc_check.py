#!/usr/bin/env python
"""Correctness + timing of the tcgen05 GP-mean tier (n = 128 fp32) against fp64 torch and the CUDA-core sweep kernel.
   tools/tc_check.py [batch]"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cuda_matrix_inversion_b200 import api


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    n = 128
    gen = torch.Generator(device="cuda").manual_seed(7)
    r = torch.rand((batch, n, n), generator=gen, device="cuda")
    b = r + r.transpose(1, 2) + n * torch.eye(n, device="cuda")
    a, c, d = (torch.rand((batch, n), generator=gen, device="cuda") for _ in range(3))
    e = torch.rand(batch, generator=gen, device="cuda")
    means = torch.zeros(batch, device="cuda")
    var = torch.zeros(batch, device="cuda")
    info = torch.full((batch,), -1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    print("tier:", api.tier_name("gp", n), "INVGPU_GP_KERNEL =", os.environ.get("INVGPU_GP_KERNEL"))
    fn = lambda: api.gp_device(n, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), e.data_ptr(), means.data_ptr(),
                               var.data_ptr(), batch, np.float32, info.data_ptr(), st)
    fn()
    torch.cuda.synchronize()
    k = min(batch, 2000)
    idx = torch.cat([torch.arange(k // 2), torch.arange(batch - k // 2, batch)]).cuda()
    m64 = b[idx].double() + torch.diag_embed(c[idx].double())
    sol_d = torch.linalg.solve(m64, d[idx].double().unsqueeze(2)).squeeze(2)
    sol_a = torch.linalg.solve(m64, a[idx].double().unsqueeze(2)).squeeze(2)
    want_m = (a[idx].double() * sol_d).sum(1)
    want_v = e[idx].double() - (a[idx].double() * sol_a).sum(1)
    em = (means[idx].double() - want_m).abs().max().item()
    ev = (var[idx].double() - want_v).abs().max().item()
    print(f"batch {batch}: max |mean - fp64| = {em:.3e}, max |var - fp64| = {ev:.3e}, info nonzero = {int((info != 0).sum())}, "
          f"mean scale {want_m.abs().max().item():.3f}")
    ts = []
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gb = ((n * n + 3 * n) * 4 + 4) * batch / ms / 1e6
    print(f"median {ms:.4f} ms -> {batch / ms * 1e3:.4e} eval/s, {gb:.1f} GB/s algorithmic = {gb / 6542.1:.3f} of the HBM roofline")
    # non-SPD flag: make matrix 3 indefinite at pivot 70
    b2 = b[:8].clone()
    b2[3, 69, 69] = -5.0
    m2 = torch.zeros(8, device="cuda"); i2 = torch.zeros(8, dtype=torch.int32, device="cuda")
    api.gp_device(n, a.data_ptr(), b2.data_ptr(), c.data_ptr(), d.data_ptr(), 0, m2.data_ptr(), 0, 8, np.float32, i2.data_ptr(), st)
    torch.cuda.synchronize()
    print("flag test: info =", i2.cpu().tolist(), "mean[3] nan:", bool(torch.isnan(m2[3])), "others ok:",
          float((m2[[0, 1, 2, 4, 5, 6, 7]] - means[[0, 1, 2, 4, 5, 6, 7]]).abs().max()))


if __name__ == "__main__":
    main()
