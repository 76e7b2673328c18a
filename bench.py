#!/usr/bin/env python
"""bench.py -- headline measurement of the batched-inversion hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]

Workload (BASELINE.json configs[2], the largest single-GPU configuration of the metric's first
term): synthetic batched SPD Cholesky inverse, 2^20 x 32x32 fp32 per GPU, A = R + R^T + n I,
R ~ U(0,1) (reference tests/generate_inverse_matrices.m:8-21).  One "step" = one pass of the
hot path over that batch.  Multi-GPU: the batch shards by rank with no data-path collective
(weak scaling: every rank inverts its own 2^20 matrices); the only exchange is the final gather
of one checksum scalar per rank.

One JSON line on stdout (rank 0):
  value     whole-job inversions/s with inputs resident in HBM (CUDA events on the launch stream,
            max over ranks); inputs (4.3 GB) and outputs (4.3 GB) are far larger than the 126 MB L2
  e2e       the same metric through the reference-facing host call (invgpu_spd_inverse_host_f32 ==
            inverse_cholesky_batched_gpu without the abort), pinned HOST buffers in, HOST buffers
            out, H2D + D2H inside the timed region
  roofline  algorithmic bytes (2 n^2 sizeof(T) per matrix) / measured kernel time vs the measured
            HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (oracle/_ref: inverse_chol_blas_omp, src/inverse.c:100,
            OpenBLAS 0.3.15) on a bounded sample, all host cores
  extra     the other BASELINE configs that fit one GPU (n sweep, fp64, Gauss-Jordan 64, fused GP
            mean 64 / 128), kernel-only, for context -- not part of the contract
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 32
BATCH = 1 << 20
METRIC = "batched SPD inversions/sec (32x32 fp32)"
UNIT = "inversions/s"
WORKLOAD = "synthetic batched SPD Cholesky inverse, 2^20 x 32x32 fp32 per GPU (BASELINE configs[2])"
# dram__bytes_read.sum + dram__bytes_write.sum of the headline kernel from the committed `ncu --set full` capture
# (2^18 matrices per launch there: 1.075257 GB read + 1.026820 GB written), scaled to the 2^20 of one bench launch
NCU_DRAM_BYTES_PER_LAUNCH = int((1.075257e9 + 1.026820e9) * 4)
NCU_TRAFFIC_SOURCE = "profiles/r1_sweep32_tma_interleaved_summary.md (ncu --set full, 2^18 matrices) x 4"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples taken before this point (warm-up) are dropped."""
        self.first = len(self.lines)

    def count_since_mark(self):
        return len(self.lines) - self.first if self.proc else 1 << 30

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _cpu_reference(sample: int, reps: int):
    """The reference's CPU SPD inverse on `sample` 32x32 matrices, all host cores."""
    import oracle as orc
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    rng = np.random.default_rng(1234)
    r = rng.random((sample, N, N), dtype=np.float32)
    a = (r + r.transpose(0, 2, 1) + N * np.eye(N, dtype=np.float32)).reshape(-1)
    if orc.ref_available():
        kind, fn = "reference", lambda: orc.ref_chol_inverse_upper(a, N)
        what = "oracle/_ref inverse_chol_blas_omp (reference src/inverse.c:100, OpenBLAS 0.3.15)"
    else:
        kind, fn = "port", lambda: orc.chol_inverse(a, N)
        what = "oracle port orc_chol_inverse_batch_f32 (OpenMP)"
    fn()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
    return kind, cores, what, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1 << 18
    kind, cores, what, times = _cpu_reference(sample, args.warmup + args.steps)
    times = times[args.warmup:]
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N, "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} matrices per step, {what}, OMP_NUM_THREADS={cores}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _time_kernel(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]


def _spd_device(torch, n, batch, dtype, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    chunk = max(1, min(batch, (256 << 20) // (n * n * 8)))
    a = torch.empty((batch, n, n), device="cuda", dtype=dtype)
    eye = n * torch.eye(n, device="cuda", dtype=dtype)
    for s in range(0, batch, chunk):
        r = torch.rand((min(chunk, batch - s), n, n), generator=gen, device="cuda", dtype=dtype)
        a[s:s + r.shape[0]] = r + r.transpose(1, 2) + eye
    return a


def _extras(torch, api, steps, warmup, hbm_peak):
    """Kernel-only numbers for the other single-GPU BASELINE configs (context, not contract)."""
    out = {}
    st = torch.cuda.current_stream().cuda_stream
    f32, f64 = torch.float32, torch.float64
    npdt = {f32: np.float32, f64: np.float64}

    def inv_case(key, n, batch, dt, general=False):
        a = _spd_device(torch, n, batch, dt, 1234 + n)
        if general:
            a = torch.rand((batch, n, n), generator=torch.Generator(device="cuda").manual_seed(20260101),
                           device="cuda", dtype=dt)
        o = torch.empty_like(a)
        call = api.general_inverse_device if general else api.spd_inverse_device
        t = _time_kernel(torch, lambda: call(a.data_ptr(), o.data_ptr(), n, batch, npdt[dt], 0, st), steps, warmup)
        ms = float(np.median(t))
        gbs = 2 * n * n * a.element_size() * batch / (ms * 1e-3) / 1e9
        out[key] = {"inversions_per_s": batch / (ms * 1e-3), "ms": ms, "algorithmic_GBps": gbs,
                    "hbm_frac": gbs / hbm_peak, "tier": api.tier_name("general" if general else "spd", n, npdt[dt])}
        del a, o

    inv_case("spd_8_f32", 8, 1 << 22, f32)
    inv_case("spd_8_f64", 8, 1 << 21, f64)
    inv_case("spd_16_f32", 16, 1 << 21, f32)
    inv_case("spd_32_f64", 32, 1 << 19, f64)
    inv_case("spd_64_f32", 64, 1 << 17, f32)
    inv_case("spd_64_f64", 64, 1 << 16, f64)
    inv_case("spd_128_f32", 128, 1 << 15, f32)
    inv_case("gauss_jordan_8_f32", 8, 1 << 22, f32, general=True)
    inv_case("gauss_jordan_32_f32", 32, 1 << 17, f32, general=True)
    inv_case("gauss_jordan_64_f32", 64, 1 << 15, f32, general=True)
    inv_case("gauss_jordan_128_f32", 128, 1 << 13, f32, general=True)

    def gp_case(key, n, batch):
        gen = torch.Generator(device="cuda").manual_seed(4321)
        b = _spd_device(torch, n, batch, f32, 4322)
        a, c, d = (torch.rand((batch, n), generator=gen, device="cuda") for _ in range(3))
        m = torch.empty(batch, device="cuda")
        t = _time_kernel(torch, lambda: api.gp_device(n, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), 0,
                                                      m.data_ptr(), 0, batch, np.float32, 0, st), steps, warmup)
        ms = float(np.median(t))
        gbs = ((n * n + 3 * n) * 4 + 4) * batch / (ms * 1e-3) / 1e9
        out[key] = {"evals_per_s": batch / (ms * 1e-3), "ms": ms, "algorithmic_GBps": gbs,
                    "hbm_frac": gbs / hbm_peak, "tier": api.tier_name("gp", n)}

    gp_case("gp_mean_8_f32", 8, 1 << 22)
    gp_case("gp_mean_16_f32", 16, 1 << 21)
    gp_case("gp_mean_32_f32", 32, 1 << 19)
    gp_case("gp_mean_64_f32", 64, 100 * 1600)
    gp_case("gp_mean_128_f32_25k", 128, 25000)

    # BASELINE config 5, one GPU's share (4 M matrices over 8 GPUs = 500 k): mixed dimensions,
    # P(n<=32)=.75 U{4..32}, P(<=128)=.20 U{33..128}, P(<=256)=.05 U{129..256}, seed 777
    rng = np.random.default_rng(777)
    cnt = 500_000
    u = rng.random(cnt)
    ns = np.where(u < 0.75, rng.integers(4, 33, cnt), np.where(u < 0.95, rng.integers(33, 129, cnt), rng.integers(129, 257, cnt))).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)])
    total = int(offs[-1])
    buf = torch.empty(total, device="cuda", dtype=f32)
    # SPD per matrix: fill with U(0,1)/n (|off-diagonal row sum| < 1) then put 2 on the diagonals -> diagonally dominant
    buf.uniform_(0.0, 1.0, generator=torch.Generator(device="cuda").manual_seed(777))
    scale = torch.from_numpy(np.repeat(1.0 / ns, ns.astype(np.int64) ** 2).astype(np.float32)).cuda()
    buf.mul_(scale)
    del scale
    starts = np.repeat(offs[:-1], ns)                          # diagonal positions of every matrix, vectorised
    k = np.arange(int(ns.sum()), dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(ns)[:-1]]), ns)
    diag = torch.from_numpy(starts + k * (np.repeat(ns, ns).astype(np.int64) + 1)).cuda()
    buf[diag] = 2.0
    # symmetrise is unnecessary: only the upper triangle is read
    outb = torch.empty_like(buf)
    pin = (buf.data_ptr() + offs[:-1] * 4).astype(np.uint64)
    pout = (outb.data_ptr() + offs[:-1] * 4).astype(np.uint64)
    info = torch.zeros(cnt, dtype=torch.int32, device="cuda")
    t = _time_kernel(torch, lambda: api.mixed_spd_inverse_device(pin, pout, ns, np.float32, info.data_ptr(), st), 3, 2)
    ms = float(np.median(t))
    gbs = 2 * 4 * total / (ms * 1e-3) / 1e9
    out["mixed_500k_f32"] = {"matrices_per_s": cnt / (ms * 1e-3), "ms": ms, "algorithmic_GBps": gbs, "hbm_frac": gbs / hbm_peak,
                             "flagged": int((info != 0).sum()), "tier": "persistent grids over nine padded sweep tiers (16/24/32/48/64/96/128/192/256), counting-sort work lists",
                             "note": "timing includes the host-side planning (~1 ms, multi-threaded counting sort) and work-list upload"}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from cuda_matrix_inversion_b200 import api, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert api.device_count() > 0, "no CUDA device: this framework has no CPU path"
    hbm_peak, peak_src = _peaks()
    stream = torch.cuda.current_stream().cuda_stream

    a = _spd_device(torch, N, BATCH, torch.float32, 1234 + rank)
    inv = torch.empty_like(a)
    info = torch.zeros(BATCH, dtype=torch.int32, device="cuda")

    def step():
        api.spd_inverse_device(a.data_ptr(), inv.data_ptr(), N, BATCH, np.float32, info.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # before the warm-up: nvidia-smi needs ~0.2 s before its first sample
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler.mark()
    launches0 = api.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    launches = api.launch_count() - launches0
    if rank == 0:
        # the timed region is tens of milliseconds, nvidia-smi samples every 100 ms: keep the same kernel running
        # (untimed) until a few samples under this load exist
        t_end = time.time() + 1.0
        while sampler.count_since_mark() < 4 and time.time() < t_end:
            step()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    assert int(info.abs().max()) == 0
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * BATCH * args.steps / (total_ms_max * 1e-3)
    kernel_ms = float(np.mean(per_step))
    algo_bytes = 2 * N * N * 4 * BATCH
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9

    # ---- end to end through the reference-facing host call: pinned HOST in -> HOST out
    h_in = torch.empty(BATCH * N * N, dtype=torch.float32).pin_memory()
    h_out = torch.empty(BATCH * N * N, dtype=torch.float32).pin_memory()
    h_in.copy_(a.reshape(-1))
    h_info = np.zeros(BATCH, dtype=np.int32)
    import ctypes as C
    pin, pout, pinfo = h_in.data_ptr(), h_out.data_ptr(), h_info.ctypes.data_as(C.c_void_p)

    def e2e_step():
        rc = lib.invgpu_spd_inverse_host_f32(pin, pout, N, BATCH, pinfo)
        assert rc == 0, rc

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(t.item())
    assert torch.equal(h_out[: 1024 * N * N].cuda(), inv.reshape(-1)[: 1024 * N * N])
    del h_in, h_out

    # final gather: one checksum scalar per rank (the only inter-GPU exchange of the workload)
    chk = inv.double().sum().reshape(1)
    if world > 1:
        from cuda_matrix_inversion_b200.sharding import gather_shards
        chk = gather_shards(chk, world).sum().reshape(1)

    if rank == 0:
        kind, cores, what, times = _cpu_reference(1 << 18, 4)
        cpu_value = (1 << 18) / min(times[1:])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": N, "batch_per_gpu": BATCH, "algorithm": "Cholesky-route inverse (potrf+trtri+lauum merged into one symmetric sweep; TMA tile I/O, sweep_kernels.cuh)",
                       "l2_policy": "inputs+outputs 8.6 GB per step >> 126 MB L2", "sharding": f"dp{world}",
                       "kernel_tier": api.tier_name("spd", N)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_TRAFFIC_SOURCE,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{1 << 18} matrices (1/4 of the GPU batch), best of 3, {what}, "
                                       f"OMP_NUM_THREADS={cores}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * N * N * 4,
                    "d2h_bytes_per_step": BATCH * N * N * 4 + BATCH * 4, "steps": e2e_steps,
                    "api": "invgpu_spd_inverse_host_f32 (pinned host buffers)"},
            "gpu_launches": launches, "clocks": clocks, "checksum": float(chk.item()),
        }
        if not args.no_extra and world == 1:
            del a, inv
            torch.cuda.empty_cache()
            lib.invgpu_release_workspace()
            line["extra"] = _extras(torch, api, 5, 3, hbm_peak)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
