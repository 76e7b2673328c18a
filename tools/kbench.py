#!/usr/bin/env python
"""Quick kernel-only timing + oracle check of one shape:  tools/kbench.py spd 32 f32 [batch]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import oracle as orc
from cuda_matrix_inversion_b200 import api


def main():
    op, n, dt = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    tdt, ndt = (torch.float32, np.float32) if dt == "f32" else (torch.float64, np.float64)
    esz = 4 if dt == "f32" else 8
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else max(1024, (1 << 32) // (n * n * esz) // 2)
    gen = torch.Generator(device="cuda").manual_seed(1)
    st = torch.cuda.current_stream().cuda_stream
    chunk = max(1, (256 << 20) // (n * n * 8))
    a = torch.empty((batch, n, n), device="cuda", dtype=tdt)
    for s in range(0, batch, chunk):
        r = torch.rand((min(chunk, batch - s), n, n), generator=gen, device="cuda", dtype=tdt)
        a[s:s + r.shape[0]] = (r + r.transpose(1, 2) + n * torch.eye(n, device="cuda", dtype=tdt)) if op != "general" else r
    o = torch.zeros_like(a)
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    if op == "spd":
        fn = lambda: api.spd_inverse_device(a.data_ptr(), o.data_ptr(), n, batch, ndt, info.data_ptr(), st)
        nbytes = 2 * n * n * esz * batch
    elif op == "general":
        fn = lambda: api.general_inverse_device(a.data_ptr(), o.data_ptr(), n, batch, ndt, info.data_ptr(), st)
        nbytes = 2 * n * n * esz * batch
    elif op == "gp":
        av, cv, dv = (torch.rand((batch, n), generator=gen, device="cuda", dtype=tdt) for _ in range(3))
        ev = torch.rand(batch, generator=gen, device="cuda", dtype=tdt)
        means = torch.zeros(batch, device="cuda", dtype=tdt)
        fn = lambda: api.gp_device(n, av.data_ptr(), a.data_ptr(), cv.data_ptr(), dv.data_ptr(), 0, means.data_ptr(), 0,
                                   batch, ndt, info.data_ptr(), st)
        nbytes = ((n * n + 3 * n) * esz + esz) * batch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    # correctness on a slice
    k = min(batch, 64)
    res = {"op": op, "n": n, "dtype": dt, "batch": batch, "ms": round(ms, 4), "units_per_s": batch / ms * 1e3,
           "GBps": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / 6542.1, "tier": api.tier_name(op, n, ndt),
           "info_nonzero": int((info != 0).sum())}
    if op == "spd":
        want, _ = orc.chol_inverse(a[:k].cpu().numpy().reshape(-1), n)
        got = o[:k].cpu().numpy().reshape(-1)
        res["err_vs_oracle"] = float(np.abs(got - want).max() / np.abs(want).max())
        last = o[-k:].cpu().numpy().reshape(-1)
        want, _ = orc.chol_inverse(a[-k:].cpu().numpy().reshape(-1), n)
        res["err_tail"] = float(np.abs(last - want).max() / np.abs(want).max())
    elif op == "general":
        want, _ = orc.gauss_jordan_inverse(orc.to_colmajor(a[:k].cpu().numpy().transpose(0, 2, 1)), n)
        got = o[:k].cpu().numpy().reshape(-1)
        res["err_vs_oracle"] = float(np.abs(got - want).max() / np.abs(want).max())
    elif op == "gp":
        want, _ = orc.gp_mean(n, av[:k].cpu().numpy().reshape(-1), a[:k].cpu().numpy().reshape(-1),
                              cv[:k].cpu().numpy().reshape(-1), dv[:k].cpu().numpy().reshape(-1))
        res["err_vs_oracle"] = float(np.abs(means[:k].cpu().numpy() - want).max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
