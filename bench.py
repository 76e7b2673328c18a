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
            out, H2D + D2H inside the timed region; `ceiling` / `e2e_ceiling` = the same pipeline with the
            kernel replaced by a device copy (copy-only: what the box's host<->device path allows with all
            ranks transferring at once), `frac_of_ceiling` = e2e / ceiling; at N > 1 also `single_gpu_value` /
            `single_gpu_ceiling` (rank 0 alone on the same box, the other ranks idle at a barrier) and
            `weak_scaling_efficiency` / `ceiling_weak_scaling_efficiency` = N-rank number / (N x the single one)
  roofline  algorithmic bytes (2 n^2 sizeof(T) per matrix) / measured kernel time vs the measured
            HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (oracle/_ref: inverse_chol_blas_omp, src/inverse.c:100,
            OpenBLAS 0.3.15) on a bounded sample, all host cores
  gp_mean_128   the metric's second term (BASELINE configs[3]) at every N: 200 000 x 128x128 fp32 fused GP means
            sharded over the ranks (strong scaling) + the final all_gather of the scalars, with its own
            roofline ((n^2+3n+1) sizeof(T) per evaluation), e2e through invgpu_gp_host_f32 (which sends only the upper-triangle
            column prefixes of B: h2d_bytes_per_step counts what crosses the bus, h2d_bytes_whole_matrices the dense layout;
            `ceiling` = a contiguous copy of the sent bytes, `whole_matrix_copy_value` = a copy of whole matrices) and
            cpu_baseline (calcluateMeanCPU, src/gauss_cpu.c:23)
  mixed     BASELINE configs[4] at every N: 500 000 mixed-dimension matrices per GPU (4 M on 8 GPUs)
  extra     (N = 1 only) the other shapes, kernel-only, for context -- not part of the contract
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU baseline runs the reference's OpenMP loops with every host core.  libgomp latches OMP_NUM_THREADS when
# it is first loaded (torch loads it) and torchrun exports OMP_NUM_THREADS=1 to its workers, so the variable is set
# here, before numpy / torch are imported; OpenBLAS stays single-threaded inside the OMP regions.
try:
    CPU_THREADS = len(os.sched_getaffinity(0))
except AttributeError:
    CPU_THREADS = os.cpu_count() or 1
os.environ["OMP_NUM_THREADS"] = str(CPU_THREADS)
os.environ["OPENBLAS_NUM_THREADS"] = "1"

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 32
BATCH = 1 << 20
METRIC = "batched SPD inversions/sec (32x32 fp32)"
UNIT = "inversions/s"
WORKLOAD = "synthetic batched SPD Cholesky inverse, 2^20 x 32x32 fp32 per GPU (BASELINE configs[2])"
# the metric's second term (BASELINE configs[3]): fused GP mean, 200 000 x 128x128 fp32 in total, sharded over the GPUs
GP_N = 128
GP_BATCH = 200_000
GP_METRIC = "fused GP-mean evaluations/sec (128x128 fp32)"
GP_UNIT = "evaluations/s"
# dram__bytes_read.sum + dram__bytes_write.sum of the tcgen05 GP kernel per evaluation, from the committed `ncu --set full` capture
# (9 472 evaluations: 405.01 MB read -- only the upper triangle of B is fetched -- + 4.47 MB written)
GP_NCU_DRAM_BYTES_PER_EVAL = (405.006592e6 + 4.473856e6) / 9472
GP_NCU_TRAFFIC_SOURCE = "profiles/r2_tc_gp128_summary.md (ncu --set full, 9472 evaluations), scaled to the evaluations of one launch"
# BASELINE configs[4]: mixed dimensions, 4 M matrices on 8 GPUs = 500 000 per GPU (weak scaling)
MIXED_PER_GPU = 500_000
# dram__bytes_read.sum + dram__bytes_write.sum of the headline kernel from the committed `ncu --set full` capture
# (2^18 matrices per launch there: 1.075248 GB read + 1.027806 GB written), scaled to the 2^20 of one bench launch
NCU_DRAM_BYTES_PER_LAUNCH = int((1.075248e9 + 1.027806e9) * 4)
NCU_TRAFFIC_SOURCE = "profiles/r2_sweep32_tma_interleaved_summary.md (ncu --set full, 2^18 matrices) x 4"


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


def _host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cpu_reference(sample: int, reps: int):
    """The reference's CPU SPD inverse on `sample` 32x32 matrices, all host cores.  The routine works in place, so
    the input is refreshed before every repetition OUTSIDE the timed region (as src/inverse_bench.c:88-97 does)."""
    import oracle as orc
    cores = CPU_THREADS
    rng = np.random.default_rng(1234)
    r = rng.random((sample, N, N), dtype=np.float32)
    a = (r + r.transpose(0, 2, 1) + N * np.eye(N, dtype=np.float32)).reshape(-1)
    work = np.empty_like(a)
    if orc.ref_available():
        kind, fn = "reference", lambda: orc.ref_chol_inverse_inplace(work, N)
        what = "oracle/_ref inverse_chol_blas_omp (reference src/inverse.c:100, OpenBLAS 0.3.15)"
    else:
        out, info = np.empty_like(a), np.zeros(sample, dtype=np.int32)
        kind, fn = "port", lambda: orc.chol_inverse(work, N)
        what = "oracle port orc_chol_inverse_batch_f32 (OpenMP)"
    times = []
    for _ in range(reps + 1):
        np.copyto(work, a)
        t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
    return kind, cores, what, times[1:]


def _cpu_reference_gp(n: int, sample: int, reps: int):
    """The reference's CPU GP mean (calcluateMeanCPU, src/gauss_cpu.c:23) on `sample` evaluations, all host cores;
    it destroys Bs / Cs, which are refreshed outside the timed region."""
    import oracle as orc
    rng = np.random.default_rng(4321)
    r = rng.random((sample, n, n), dtype=np.float32)
    b = (r + r.transpose(0, 2, 1) + n * np.eye(n, dtype=np.float32)).reshape(-1)
    a, c, d = (rng.random(sample * n, dtype=np.float32) for _ in range(3))
    out = np.zeros(sample, dtype=np.float32)
    wb, wc = np.empty_like(b), np.empty_like(c)
    if orc.ref_available():
        kind, fn = "reference", lambda: orc.ref_gp_mean_inplace(n, a, wb, wc, d, out)
        what = "oracle/_ref calcluateMeanCPU (reference src/gauss_cpu.c:23, OpenBLAS 0.3.15)"
    else:
        kind, fn = "port", lambda: orc.gp_mean(n, a, wb, wc, d)
        what = "oracle port orc_gp_batch_f32 (OpenMP)"
    times = []
    for _ in range(reps + 1):
        np.copyto(wb, b); np.copyto(wc, c)
        t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
    return kind, CPU_THREADS, what, times[1:]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1 << 18
    kind, cores, what, times = _cpu_reference(sample, args.warmup + args.steps)
    times = times[args.warmup:]
    total = sum(times)
    value = sample * len(times) / total
    gk, _, gwhat, gtimes = _cpu_reference_gp(GP_N, 8192, 3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N, "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} matrices per step, {what}, OMP_NUM_THREADS={cores}; the in-place routine's "
                                   "input is refreshed outside the timed region"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gp_mean_128": {"metric": GP_METRIC, "value": 8192 / min(gtimes), "unit": GP_UNIT, "kind": gk, "cores": cores,
                        "sample": f"8192 evaluations, best of 3, {gwhat}"},
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

    return out


def _mixed_workload(torch, rank):
    """BASELINE config 5, one GPU's share (4 M matrices over 8 GPUs = 500 k per GPU): mixed dimensions,
    P(n<=32)=.75 U{4..32}, P(<=128)=.20 U{33..128}, P(<=256)=.05 U{129..256}, seed 777 (+ rank)."""
    f32 = torch.float32
    rng = np.random.default_rng(777 + rank)
    cnt = MIXED_PER_GPU
    u = rng.random(cnt)
    ns = np.where(u < 0.75, rng.integers(4, 33, cnt), np.where(u < 0.95, rng.integers(33, 129, cnt), rng.integers(129, 257, cnt))).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(ns.astype(np.int64) ** 2)])
    total = int(offs[-1])
    buf = torch.empty(total, device="cuda", dtype=f32)
    # SPD per matrix: fill with U(0,1)/n (|off-diagonal row sum| < 1) then put 2 on the diagonals -> diagonally dominant
    buf.uniform_(0.0, 1.0, generator=torch.Generator(device="cuda").manual_seed(777 + rank))
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
    return ns, total, buf, outb, pin, pout, info


class PinnedF32:
    """float32 host buffer from the engine's pinned allocator (invgpu_host_alloc), viewed as numpy / torch."""

    def __init__(self, lib, count):
        import ctypes as C
        self.lib, self.count = lib, count
        self.ptr = lib.invgpu_host_alloc(count * 4)
        assert self.ptr, "invgpu_host_alloc failed"
        self.np = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_float)), shape=(count,))

    def torch(self, torch):
        return torch.from_numpy(self.np)

    def free(self):
        if self.ptr:
            self.np = None
            self.lib.invgpu_host_free(self.ptr)
            self.ptr = None


def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from cuda_matrix_inversion_b200 import api, lib
    from cuda_matrix_inversion_b200.sharding import gather_shards, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert api.device_count() > 0, "no CUDA device: this framework has no CPU path"
    hbm_peak, peak_src = _peaks()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(step, steps, warmup):
        """W warm-up steps, then K steps bracketed by barrier + synchronize; CUDA events on the launch stream."""
        for _ in range(warmup):
            step()
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record()
        for i in range(steps):
            step()
            evs[i + 1].record()
        barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        return max_over_ranks(evs[0].elapsed_time(evs[-1])), per

    def timed_host(step, steps, warmup):
        """Host-flavour (synchronous) calls: wall clock around K calls after a barrier, max over ranks."""
        for _ in range(warmup):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    # ================================================================ headline: 2^20 x 32x32 fp32 SPD inverse per GPU
    a = _spd_device(torch, N, BATCH, torch.float32, 1234 + rank)
    inv = torch.empty_like(a)
    info = torch.zeros(BATCH, dtype=torch.int32, device="cuda")

    def step():
        api.spd_inverse_device(a.data_ptr(), inv.data_ptr(), N, BATCH, np.float32, info.data_ptr(), stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # before the warm-up: nvidia-smi needs ~0.2 s before its first sample
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    sampler.mark()
    launches0 = api.launch_count()
    total_ms_max, per_step = timed_device(step, args.steps, 0)
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
    value = world * BATCH * args.steps / (total_ms_max * 1e-3)
    kernel_ms = float(np.mean(per_step))
    algo_bytes = 2 * N * N * 4 * BATCH
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9

    # ---- end to end through the reference-facing host call: pinned HOST in -> HOST out
    h_in, h_out = PinnedF32(lib, BATCH * N * N), PinnedF32(lib, BATCH * N * N)
    h_in.torch(torch).copy_(a.reshape(-1))
    h_info = np.zeros(BATCH, dtype=np.int32)
    pinfo = h_info.ctypes.data_as(C.c_void_p)

    def e2e_step():
        rc = lib.invgpu_spd_inverse_host_f32(h_in.ptr, h_out.ptr, N, BATCH, pinfo)
        assert rc == 0, rc

    e2e_steps = max(2, min(args.steps, 5))
    e2e_s = timed_host(e2e_step, e2e_steps, 2)
    e2e_value = world * BATCH * e2e_steps / e2e_s
    assert torch.equal(h_out.torch(torch)[: 1024 * N * N].cuda(), inv.reshape(-1)[: 1024 * N * N])

    # ---- the ceiling of that call on this box: the same pipeline (chunks, streams, ring) with the kernel replaced by a
    # device copy (invgpu_xfer_roundtrip_host) -- copy-only, same bytes per step, same ranks concurrently
    def ceil_step():
        rc = lib.invgpu_xfer_roundtrip_host(h_in.ptr, h_out.ptr, N * N * 4, BATCH)
        assert rc == 0, rc

    ceil_s = timed_host(ceil_step, e2e_steps, 1)
    ceil_value = world * BATCH * e2e_steps / ceil_s
    # the same two calls with rank 0 ALONE on the box (the other ranks wait at the barrier): what one GPU gets when nobody else
    # uses the host path -- end-to-end weak-scaling efficiency = e2e / (N x this), and the same for the copy-only ceiling
    e2e_single = ceil_single = None
    if world > 1:
        barrier()
        if rank == 0:
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            e2e_single = BATCH * e2e_steps / (time.perf_counter() - t0)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ceil_step()
            ceil_single = BATCH * e2e_steps / (time.perf_counter() - t0)
        barrier()
    h_in.free(); h_out.free()

    # final gather: one checksum scalar per rank (the only inter-GPU exchange of this workload)
    chk = inv.double().sum().reshape(1)
    if world > 1:
        chk = gather_shards(chk, world).sum().reshape(1)
    del a, inv, info
    torch.cuda.empty_cache()
    lib.invgpu_release_workspace()

    # ================================================================ GP mean: 200 000 x 128x128 fp32 in total, sharded
    lo, hi = shard_bounds(GP_BATCH, world, rank)
    gb = hi - lo
    n = GP_N
    gen = torch.Generator(device="cuda").manual_seed(4321 + rank)
    gB = _spd_device(torch, n, gb, torch.float32, 4322 + rank)
    gA, gC, gD = (torch.rand((gb, n), generator=gen, device="cuda") for _ in range(3))
    per = -(-GP_BATCH // world)
    g_means = torch.zeros(per, device="cuda")                 # the fused kernel writes straight into the gather's send buffer
    g_info = torch.zeros(gb, dtype=torch.int32, device="cuda")

    def gp_step():
        api.gp_device(n, gA.data_ptr(), gB.data_ptr(), gC.data_ptr(), gD.data_ptr(), 0, g_means.data_ptr(), 0, gb, np.float32,
                      g_info.data_ptr(), stream)

    gp_steps = max(3, min(args.steps, 10))
    gl0 = api.launch_count()
    gp_ms_max, gp_per = timed_device(gp_step, gp_steps, 3)
    gp_launches = api.launch_count() - gl0
    assert int(g_info.abs().max()) == 0
    gp_value = GP_BATCH * gp_steps / (gp_ms_max * 1e-3)       # strong scaling: the 200 000 evaluations are the whole job
    gp_kernel_ms = float(np.mean(gp_per))
    gp_algo = ((n * n + 3 * n) * 4 + 4) * gb
    gp_ach = gp_algo / (gp_kernel_ms * 1e-3) / 1e9
    # the final gather of the scalars (the path's only collective): all_gather of ceil(B/G) floats per rank
    gather_ms = 0.0
    if world > 1:
        outs = [torch.empty_like(g_means) for _ in range(world)]
        for _ in range(2):
            dist.all_gather(outs, g_means)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather(outs, g_means)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(e0.elapsed_time(e1))
        all_means = torch.cat(outs)[:GP_BATCH]
    else:
        all_means = g_means[:GP_BATCH]
    gp_checksum = float(all_means.double().sum().item())
    # parity spot check of this run's numbers against fp64 torch on a slice
    kk = min(gb, 64)
    m64 = gB[:kk].double() + torch.diag_embed(gC[:kk].double())
    ref64 = (gA[:kk].double().unsqueeze(1) @ torch.linalg.solve(m64, gD[:kk].double().unsqueeze(2))).reshape(-1)
    assert float((g_means[:kk].double() - ref64).abs().max()) <= 1e-4
    # end to end: pinned host A, B, C, D in -> host means out through invgpu_gp_host_f32 (== calcluateMeanGPU)
    hB, hA, hC, hD = PinnedF32(lib, gb * n * n), PinnedF32(lib, gb * n), PinnedF32(lib, gb * n), PinnedF32(lib, gb * n)
    hB.torch(torch).copy_(gB.reshape(-1)); hA.torch(torch).copy_(gA.reshape(-1))
    hC.torch(torch).copy_(gC.reshape(-1)); hD.torch(torch).copy_(gD.reshape(-1))
    h_means = np.zeros(gb, dtype=np.float32)
    h_ginfo = np.zeros(gb, dtype=np.int32)

    def gp_e2e_step():
        rc = lib.invgpu_gp_host_f32(n, hA.ptr, hB.ptr, hC.ptr, hD.ptr, None, h_means.ctypes.data, None, gb, h_ginfo.ctypes.data)
        assert rc == 0, rc

    gp_e2e_steps = 3
    gp_e2e_s = timed_host(gp_e2e_step, gp_e2e_steps, 1)
    gp_e2e_value = GP_BATCH * gp_e2e_steps / gp_e2e_s
    assert np.array_equal(h_means[:1024], g_means[:1024].cpu().numpy())

    # The tcgen05 tier reads only the upper triangle of B, so the host call sends the column prefixes (csrc/capi.cu
    # h2d_upper_triangle: 32-column groups, strided 3-D copies): bytes per matrix that actually cross the bus
    gw = int(os.environ.get("INVGPU_GP_UPPER_W", "32"))      # columns per group (capi.cu default)
    b_sent = sum(min(n, (g + 1) * gw) * min(gw, n - gw * g) * 4 for g in range((n + gw - 1) // gw)) if lib.invgpu_gp_upper_h2d(n, 4) else n * n * 4
    sent_elems = gb * b_sent // 4

    # copy-only ceilings of that call: host -> device, nothing else -- (a) the bytes the call sends, contiguous; (b) whole matrices
    def gp_ceil_step():
        gB.reshape(-1)[:sent_elems].copy_(hB.torch(torch)[:sent_elems], non_blocking=True)
        gA.reshape(-1).copy_(hA.torch(torch), non_blocking=True)
        gC.reshape(-1).copy_(hC.torch(torch), non_blocking=True)
        gD.reshape(-1).copy_(hD.torch(torch), non_blocking=True)
        torch.cuda.synchronize()

    gp_ceil_s = timed_host(gp_ceil_step, gp_e2e_steps, 1)
    gp_ceil_value = GP_BATCH * gp_e2e_steps / gp_ceil_s
    sent_elems = gb * n * n
    gp_full_s = timed_host(gp_ceil_step, gp_e2e_steps, 1)
    gp_full_value = GP_BATCH * gp_e2e_steps / gp_full_s
    for h in (hB, hA, hC, hD):
        h.free()
    del gA, gB, gC, gD, g_means, g_info
    torch.cuda.empty_cache()
    lib.invgpu_release_workspace()

    # ================================================================ mixed dimensions: 500 000 matrices per GPU
    ns, mtotal, mbuf, mout, mpin, mpout, minfo = _mixed_workload(torch, rank)
    ml0 = api.launch_count()
    mixed_ms_max, mixed_per = timed_device(lambda: api.mixed_spd_inverse_device(mpin, mpout, ns, np.float32, minfo.data_ptr(), stream), 3, 2)
    mixed_launches = (api.launch_count() - ml0) // 5
    mixed_flagged = int((minfo != 0).sum())
    mixed_value = world * MIXED_PER_GPU * 3 / (mixed_ms_max * 1e-3)
    mixed_ms = float(np.mean(mixed_per))
    mixed_gbs = 2 * 4 * mtotal / (mixed_ms * 1e-3) / 1e9
    del mbuf, mout, minfo
    torch.cuda.empty_cache()
    lib.invgpu_release_workspace()

    if rank == 0:
        kind, cores, what, times = _cpu_reference(1 << 18, 3)
        cpu_value = (1 << 18) / min(times)
        gkind, _, gwhat, gtimes = _cpu_reference_gp(GP_N, 8192, 3)
        gp_cpu_value = 8192 / min(gtimes)
        h2d, d2h = BATCH * N * N * 4, BATCH * N * N * 4 + BATCH * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
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
                                       f"OMP_NUM_THREADS={cores} (set before libgomp loads); input refreshed outside the timed region"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "invgpu_spd_inverse_host_f32 (pinned host buffers from invgpu_host_alloc)",
                    "per_gpu": e2e_value / world, "host_GBps_aggregate": world * (h2d + d2h) * e2e_steps / e2e_s / 1e9,
                    "ceiling": ceil_value, "frac_of_ceiling": e2e_value / ceil_value,
                    "ceiling_what": "invgpu_xfer_roundtrip_host: the same host pipeline with the kernel replaced by a device copy, "
                                    "same bytes, all ranks concurrently",
                    "single_gpu_value": e2e_single, "single_gpu_ceiling": ceil_single,
                    "weak_scaling_efficiency": (e2e_value / (world * e2e_single)) if e2e_single else None,
                    "ceiling_weak_scaling_efficiency": (ceil_value / (world * ceil_single)) if ceil_single else None,
                    "efficiency_what": "e2e (resp. copy-only ceiling) at N ranks / (N x the same call with rank 0 alone on this box): "
                                       "how far the box's shared host<->device path lets N GPUs scale"},
            "e2e_ceiling": ceil_value,
            "gp_mean_128": {
                "metric": GP_METRIC, "value": gp_value, "unit": GP_UNIT, "scaling": "strong", "batch_total": GP_BATCH,
                "batch_per_gpu": gb, "steps": gp_steps, "ms_per_step": gp_ms_max / gp_steps, "dtype": "f32",
                "config": {"workload": "fused GP mean A^T (B + diag C)^-1 D, 200 000 x 128x128 fp32 (BASELINE configs[3]), "
                                       f"contiguous shards over {world} GPU(s), final all_gather of the scalars",
                           "kernel_tier": api.tier_name("gp", n)},
                "roofline": {"bound": "hbm", "achieved": gp_ach, "peak": hbm_peak, "unit": "GB/s", "frac": gp_ach / hbm_peak,
                             "algorithmic_bytes_per_launch": gp_algo, "kernel_ms": gp_kernel_ms,
                             "traffic": int(GP_NCU_DRAM_BYTES_PER_EVAL * gb), "traffic_source": GP_NCU_TRAFFIC_SOURCE,
                             "fp32_flops_per_eval": n ** 3 / 3 + 2 * n * n + 3 * n},
                "gather_ms": gather_ms, "checksum": gp_checksum, "gpu_launches": gp_launches,
                "e2e": {"value": gp_e2e_value, "unit": GP_UNIT, "h2d_bytes_per_step": gb * (b_sent + 3 * n * 4),
                        "h2d_bytes_whole_matrices": gb * (n * n + 3 * n) * 4,
                        "h2d_note": "only the upper triangle of B is read by the kernel: the call sends the column prefixes (strided copies), not whole matrices",
                        "d2h_bytes_per_step": gb * 8, "steps": gp_e2e_steps, "api": "invgpu_gp_host_f32 (== calcluateMeanGPU), pinned host buffers",
                        "ceiling": gp_ceil_value, "frac_of_ceiling": gp_e2e_value / gp_ceil_value,
                        "ceiling_what": "cudaMemcpyAsync of the bytes the call sends (contiguous) host -> device, all ranks concurrently",
                        "whole_matrix_copy_value": gp_full_value,
                        "whole_matrix_copy_what": "the same, whole matrices (what a caller without the prefix transfer would at best reach)"},
                "cpu_baseline": {"value": gp_cpu_value, "unit": GP_UNIT, "cores": cores, "kind": gkind,
                                 "sample": f"8192 evaluations, best of 3, {gwhat}, OMP_NUM_THREADS={cores}"},
            },
            "mixed": {
                "metric": "mixed-dimension SPD inversions/sec (n in 4..256, BASELINE configs[4])", "value": mixed_value,
                "unit": "matrices/s", "scaling": "weak", "matrices_per_gpu": MIXED_PER_GPU, "matrices_total": world * MIXED_PER_GPU,
                "ms_per_step": mixed_ms_max / 3, "algorithmic_GBps_per_gpu": mixed_gbs, "hbm_frac": mixed_gbs / hbm_peak,
                "flagged": mixed_flagged, "kernels_per_step": mixed_launches,
                "note": "timing includes the host-side planning (multi-threaded counting sort) and the work-list upload",
            },
            "gpu_launches": launches, "clocks": clocks, "checksum": float(chk.item()),
        }
        if not args.no_extra and world == 1:
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
