#!/usr/bin/env python
"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): launches and time per kernel,
and K1's share of the step kernels.  usage: tools/launch_summary.py launches.csv "<header line>" """
import collections
import csv
import json
import sys


def main():
    path, header = sys.argv[1], sys.argv[2]
    bench_share = None
    if len(sys.argv) > 3:
        for line in open(sys.argv[3]):
            if line.startswith("{"):
                bench_share = json.loads(line)["roofline"]["kernel_share_of_step"]["K1"]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        ns = float(r[iV].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iU], 1.0)
        tot[r[iK]] += ns
        cnt[r[iK]] += 1
    total = sum(tot.values())
    print(header)
    for k, ns in tot.most_common():
        print("%-72s n=%3d  total %10.1f us  per launch %10.1f us  %5.1f%%" % (k[:72], cnt[k], ns / 1e3, ns / 1e3 / cnt[k], 100 * ns / total))
    step = {k: v for k, v in tot.items() if any(t in k for t in ("pair_full_fast_kernel", "rhok_build", "ksum_kernel"))}
    k1 = sum(v for k, v in step.items() if "pair_full_fast" in k)
    if step:
        per = {k: tot[k] / cnt[k] for k in tot if any(t in k for t in ("pair_full_fast_kernel", "rhok_build", "ksum_kernel", "finalize_kernel"))}
        one = sum(per.values())
        print("one step = K2 rho_k rebuild + K1 pair action + K3 k-sum + finalize, per-launch averages: %.1f us under ncu (serialised, cold cache); K1 %.1f%%, K2 %.1f%%"
              % (one / 1e3, 100 * sum(v for k, v in per.items() if "pair_full_fast" in k) / one, 100 * sum(v for k, v in per.items() if "rhok_build" in k) / one))
        msg = "K1 share of the step kernels (K1 + K2 + K3, all launches of the list): %.1f%%" % (100 * k1 / sum(step.values()))
        if bench_share is not None:
            msg += " (bench.py event-timed share: %.1f%%)" % (100 * bench_share)
        print(msg)


if __name__ == "__main__":
    main()
