#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step).
usage: python tools/ncu_summary.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for row in r:
        name = row[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("qagnn::<", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(row[vi])
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"total {tot / 1e6:.3f} ms of GPU time over {sum(v[0] for v in agg.values())} launches "
          f"(cold-cache, serialised: compare shares, not absolutes)\n")
    print("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
