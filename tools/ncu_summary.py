#!/usr/bin/env python
"""Turn an .ncu-rep (read here, no GPU needed) into the markdown summary kept under profiles/.
   python tools/ncu_summary.py gpurun_out/tile32_r1.ncu-rep profiles/r1_tile32_summary.md [launches.csv] [--note "batch ..., command ..."]
The note records the workload (batch, command line) so that traffic vs algorithmic bytes can be recomputed from the summary."""
import collections
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]


def main():
    note = None
    if "--note" in sys.argv:
        i = sys.argv.index("--note")
        note = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    rep, dst = sys.argv[1], sys.argv[2]
    hdr, units, vals = raw(rep)
    get = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
    keys = [
        "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    ]
    lines = [f"# ncu summary of `{rep}`", ""]
    if note:
        lines += [f"Workload: {note}", ""]
    lines += ["| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in get:
            lines.append(f"| {k} | {get[k][0]} | {get[k][1]} |")
    lines += ["", "## warp stall reasons (average warps per issue slot)", "", "| reason | ratio |", "|---|---|"]
    st = [(h, v) for h, (v, u) in get.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for h, v in sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:10]:
        lines.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v} |")
    if len(sys.argv) > 3:
        rows = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 5]
        h = rows[0]
        ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
        agg = collections.defaultdict(list)
        for r in rows[1:]:
            try:
                agg[r[ki][:110] + " grid " + r[gi]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
        lines += ["", f"## launch list of `{sys.argv[3]}` (gpu__time_duration, cold-cache / serialised)", "",
                  "| launches | mean us | kernel |", "|---|---|---|"]
        for k, v in agg.items():
            lines.append(f"| {len(v)} | {sum(v) / len(v) / 1e3:.1f} | `{k}` |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
