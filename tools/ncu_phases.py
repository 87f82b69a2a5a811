#!/usr/bin/env python
"""Split an ncu source page (SASS view) of one kernel into the code regions between BAR.SYNC
instructions and report, per region, stall samples, warp instructions, shared-memory wavefronts
and global/local L1 tag requests.  usage: tools/ncu_phases.py report.ncu-rep"""
import csv
import io
import subprocess
import sys


def main():
    src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    col = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Tag Requests Global",
                                   "Address Space", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_math", "stall_mio", "stall_lg", "stall_wait")}
    def num(r, n):
        try:
            return float(r[col[n]])
        except (ValueError, IndexError):
            return 0.0
    regions, cur = [], dict(n=0, samples=0, inst=0, wf=0, tag=0, lsb=0, ssb=0, bar=0, math=0, mio=0, lg=0, wait=0, local=0, first="")
    for r in rows[hi + 1:]:
        if len(r) <= col["Instructions Executed"]:
            continue
        s = r[col["Source"]].strip()
        if not cur["first"]:
            cur["first"] = s[:40]
        cur["n"] += 1
        cur["samples"] += num(r, "# Samples")
        cur["inst"] += num(r, "Instructions Executed")
        cur["wf"] += num(r, "L1 Wavefronts Shared")
        cur["tag"] += num(r, "L1 Tag Requests Global")
        if r[col["Address Space"]].strip().lower().startswith("local"):
            cur["local"] += num(r, "Instructions Executed")
        for k, n in (("lsb", "stall_long_sb"), ("ssb", "stall_short_sb"), ("bar", "stall_barrier"), ("math", "stall_math"), ("mio", "stall_mio"), ("lg", "stall_lg"), ("wait", "stall_wait")):
            cur[k] += num(r, n)
        if "BAR.SYNC" in s:
            regions.append(cur)
            cur = dict(n=0, samples=0, inst=0, wf=0, tag=0, lsb=0, ssb=0, bar=0, math=0, mio=0, lg=0, wait=0, local=0, first="")
    regions.append(cur)
    tot = sum(x["samples"] for x in regions) or 1
    print("region  sass  samples%%  warp-inst    smem-wf      gl-tags   local-inst | stall samples: long_sb short_sb barrier math mio lg wait")
    for i, x in enumerate(regions):
        print("%3d   %5d   %6.1f  %10.3e  %10.3e  %10.3e  %10.3e | %6d %6d %6d %6d %6d %6d %6d" % (
            i, x["n"], 100 * x["samples"] / tot, x["inst"], x["wf"], x["tag"], x["local"], x["lsb"], x["ssb"], x["bar"], x["math"], x["mio"], x["lg"], x["wait"]))


if __name__ == "__main__":
    main()
