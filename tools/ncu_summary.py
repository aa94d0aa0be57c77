#!/usr/bin/env python
"""Developer tool: summarise an .ncu-rep (raw page + per-opcode / per-line histogram from the
source page).  Usage: python tools/ncu_summary.py report.ncu-rep [n_points]"""
import collections
import csv
import io
import subprocess
import sys


FILTER = []
for _a in sys.argv:
    if _a.startswith("--kernel="):
        FILTER = ["-k", "regex:" + _a[len("--kernel="):]]
sys.argv = [a for a in sys.argv if not a.startswith("--kernel=")]


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"] + FILTER + list(extra),
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    npts = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads",
            "sass__inst_executed_local_stores", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.avg.per_cycle_active",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
    for r in rows[2:3]:
        print("kernel:", r[hdr.index("Kernel Name")][:90])
        for w in want:
            for i, h in enumerate(hdr):
                if h == w:
                    print("  %-70s %-14s %s" % (h, units[i], r[i]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        st.sort(reverse=True)
        print("  stalls per issue:", ", ".join("%s=%.2f" % (n, v) for v, n in st[:8]))
    rows = page(rep, "source", ["--print-source", "sass"])
    hdr = rows[1]
    iA, iI, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    iT = hdr.index("Thread Instructions Executed")
    ops, samp = collections.Counter(), collections.Counter()
    tot = thr = 0
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) <= iI or not r[iI].isdigit():
            continue
        parts = r[iA].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        ops[op] += int(r[iI])
        samp[op] += int(r[iS])
        tot += int(r[iI])
        thr += int(r[iT])
    print("warp instructions: %d  thread instructions: %d" % (tot, thr))
    if npts:
        print("per point: %.0f warp-instr, %.0f thread-instr" % (tot / npts, thr / npts))
    ssum = sum(samp.values())
    for op, n in ops.most_common(22):
        print("  %-8s %5.1f%% of instr   %5.1f%% of samples" % (op, 100.0 * n / tot, 100.0 * samp[op] / max(ssum, 1)))


if __name__ == "__main__":
    main()
