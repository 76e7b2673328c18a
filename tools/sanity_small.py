#!/usr/bin/env python
"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck):  tools/sanity_small.py [what ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import oracle as orc
from cuda_matrix_inversion_b200 import api


def main():
    what = set(sys.argv[1:]) or {"tc", "gj", "lu", "mixed"}
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    if "tc" in what:
        n, batch = 128, 700
        r = rng.random((batch, n, n))
        b = torch.from_numpy(orc.to_colmajor((r + r.transpose(0, 2, 1) + n * np.eye(n)).astype(np.float32))).cuda()
        a, c, d = (torch.from_numpy(rng.random(batch * n).astype(np.float32)).cuda() for _ in range(3))
        m = torch.zeros(batch, device="cuda"); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
        api.gp_device(n, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), 0, m.data_ptr(), 0, batch, np.float32, info.data_ptr(), st)
        torch.cuda.synchronize()
        print("tc gp128 ok", float(m.abs().max()), int(info.abs().max()))
    if "gj" in what:
        for n, dt in ((16, np.float32), (24, np.float32), (32, np.float32), (64, np.float32), (48, np.float32), (16, np.float64), (32, np.float64)):
            batch = 37
            g = (rng.random((batch, n, n)) + np.eye(n)).astype(dt)
            flat = orc.to_colmajor(g)
            got, info = api.general_inverse_host(flat, n)
            want, _ = orc.gauss_jordan_inverse(flat, n)
            print("gj", n, np.dtype(dt).name, float(np.abs(got - want).max() / np.abs(want).max()), int(info.any()))
    if "lu" in what:
        for n in (8, 33, 100):
            batch, nrhs = 9, 3
            g = (rng.random((batch, n, n)) + np.eye(n) * n / 8).astype(np.float32)
            d_a = torch.from_numpy(orc.to_colmajor(g).copy()).cuda()
            d_b = torch.from_numpy(rng.random(batch * n * nrhs).astype(np.float32)).cuda()
            d_piv = torch.zeros(batch * n, dtype=torch.int32, device="cuda"); d_info = torch.zeros(batch, dtype=torch.int32, device="cuda")
            d_inv = torch.empty_like(d_a)
            api.gesv_device(d_a.data_ptr(), d_b.data_ptr(), n, nrhs, batch, np.float32, d_piv.data_ptr(), d_info.data_ptr(), st)
            api.getri_device(d_a.data_ptr(), d_piv.data_ptr(), d_inv.data_ptr(), n, batch, np.float32, d_info.data_ptr(), st)
            torch.cuda.synchronize()
            print("lu", n, int(d_info.abs().max()))
    if "mixed" in what:
        ns = rng.integers(2, 200, size=300).astype(np.int32)
        offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)])
        mats = []
        for k in ns:
            r = rng.random((k, k)); mats.append((r + r.T + k * np.eye(k)).astype(np.float32).T.reshape(-1))
        buf = torch.from_numpy(np.concatenate(mats)).cuda(); out = torch.zeros_like(buf)
        info = torch.zeros(len(ns), dtype=torch.int32, device="cuda")
        pin = (buf.data_ptr() + offs[:-1] * 4).astype(np.uint64); pout = (out.data_ptr() + offs[:-1] * 4).astype(np.uint64)
        api.mixed_spd_inverse_device(pin, pout, ns, np.float32, info.data_ptr(), st)
        torch.cuda.synchronize()
        print("mixed ok", int(info.abs().max()))


if __name__ == "__main__":
    main()
