#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md argues from.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [warp_evals] > profiles/<name>.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    warp_evals = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name"))
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                print("  %-88s %s %s" % (h, v, u))
        if warp_evals:
            inst = float(d["smsp__inst_executed.sum"])
            wf = float(d.get("l1tex__data_pipe_lsu_wavefronts.sum", "nan") or "nan")
            print("  per warp-evaluation: %.1f warp instructions, %.1f LSU wavefronts" % (inst / warp_evals, wf / warp_evals))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
        op = t.split()[0].split(".")[0]
        ops[op] += int(r[iE])
        samp[op] += int(r[iN])
    tot, ts = sum(ops.values()), max(1, sum(samp.values()))
    print("  instruction mix (warp instructions executed, share; stall-sample share):")
    for op, n in ops.most_common(16):
        extra = " %7.2f/warp-eval" % (n / warp_evals) if warp_evals else ""
        print("    %-10s %5.1f%%  samples %5.1f%%%s" % (op, 100.0 * n / tot, 100.0 * samp[op] / ts, extra))


if __name__ == "__main__":
    main()
