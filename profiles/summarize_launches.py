#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and
the kernel sequence of one rollout step (between two build_input launches)."""
import collections
import csv
import sys


def main(path, step_index=3):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in data if len(r) > vi]
    idx = [i for i, (n, _, _) in enumerate(seq) if "build_input" in n]
    if len(idx) < 2:
        raise SystemExit(f"need two build_input launches to delimit a step; the capture has {len(idx)} (skip more warm-up launches: ncu -s)")
    step_index = min(step_index, len(idx) - 2)
    step = seq[idx[step_index]:idx[step_index + 1]]
    own = [s for s in step if "pbmc::" in s[0]] or [s for s in step if "at::" not in s[0] and "nccl" not in s[0].lower()]
    tot = sum(v for _, v, _ in own)
    print(f"# one rollout step: {len(own)} pbmc kernels, sum of serialised cold-cache durations {tot/1e3:.1f} us")
    agg = collections.OrderedDict()
    for n, v, g in own:
        key = (n.split("(")[0].replace("void ", ""), g)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"{'kernel':58s} {'grid':>14s} {'n':>3s} {'us total':>9s} {'share':>6s}")
    for (n, g), (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:58]:58s} {g:>14s} {c:3d} {v/1e3:9.1f} {100*v/tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
