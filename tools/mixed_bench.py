#!/usr/bin/env python
"""Mixed-dimension batch (BASELINE config 5, one GPU's share) timing + check:  tools/mixed_bench.py [count]
   INVGPU_MIXED_TIMING=1 prints the host planning time and per-bucket kernel times on stderr."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import oracle as orc
from cuda_matrix_inversion_b200 import api


def main():
    cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    f32 = torch.float32
    rng = np.random.default_rng(777)
    u = rng.random(cnt)
    ns = np.where(u < 0.75, rng.integers(4, 33, cnt), np.where(u < 0.95, rng.integers(33, 129, cnt), rng.integers(129, 257, cnt))).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)])
    total = int(offs[-1])
    buf = torch.empty(total, device="cuda", dtype=f32)
    buf.uniform_(0.0, 1.0, generator=torch.Generator(device="cuda").manual_seed(777))
    scale = torch.from_numpy(np.repeat(1.0 / ns, ns.astype(np.int64) ** 2).astype(np.float32)).cuda()
    buf.mul_(scale)
    del scale
    starts = np.repeat(offs[:-1], ns)
    k = np.arange(int(ns.sum()), dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(ns)[:-1]]), ns)
    diag = torch.from_numpy(starts + k * (np.repeat(ns, ns).astype(np.int64) + 1)).cuda()
    buf[diag] = 2.0
    outb = torch.zeros_like(buf)
    pin = (buf.data_ptr() + offs[:-1] * 4).astype(np.uint64)
    pout = (outb.data_ptr() + offs[:-1] * 4).astype(np.uint64)
    info = torch.zeros(cnt, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: api.mixed_spd_inverse_device(pin, pout, ns, np.float32, info.data_ptr(), st)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts, wall = [], []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        wall.append((time.perf_counter() - w0) * 1e3)
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    # check a few matrices of every bucket against the oracle (symmetrised from the upper triangle)
    worst = 0.0
    for lo, hi in ((4, 32), (33, 128), (129, 256)):
        idx = np.nonzero((ns >= lo) & (ns <= hi))[0][:3]
        for i in idx:
            n = int(ns[i])
            a = buf[offs[i]:offs[i + 1]].cpu().numpy().astype(np.float64).reshape(n, n)   # column-major: a[c, r]
            full = np.triu(a.T) + np.triu(a.T, 1).T
            want = np.linalg.inv(full)
            got = outb[offs[i]:offs[i + 1]].cpu().numpy().reshape(n, n).T
            worst = max(worst, float(np.abs(got - want).max() / np.abs(want).max()))
    print(json.dumps({"count": cnt, "ms_events": ms, "ms_wall": float(np.median(wall)), "matrices_per_s": cnt / ms * 1e3,
                      "GBps": 2 * 4 * total / ms / 1e6, "hbm_frac": 2 * 4 * total / ms / 1e6 / 6542.1,
                      "flagged": int((info != 0).sum()), "max_rel_err": worst}))


if __name__ == "__main__":
    main()
