#!/usr/bin/env python
"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: which kernels
of one bench step are ours and what share of the serialized, cold-cache device time they take."""
import collections
import csv
import re
import sys


def main(path, last=None):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    if last:
        rows = rows[-int(last):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"]
        m = re.search(r"(stream_kernel|tma_kernel)<.*?::(\w+Op)<([^>]*)>", name)
        if m:
            key = f"dd4ml_b200 {m.group(1)}<{m.group(2)}<{m.group(3)}>>"
        elif re.search(r"<unnamed>::(state_\w+|gather_kernel|seg_copy_kernel|kspace_kernel|combine_kernel|gather_rows_kernel)", name):
            key = "dd4ml_b200 " + re.search(r"<unnamed>::(\w+)", name).group(1)
        else:
            key = "torch  " + re.sub(r"[<(].*", "", name.replace("void ", ""))[:60]
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1000.0
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1000.0
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("dd4ml"))
    print(f"| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {v[1] / tot:.3f} |")
    print(f"\nlaunches {len(rows)}, total {tot:.0f} us, dd4ml_b200 kernels {ours:.0f} us = {ours / tot:.3f} of the device time")


if __name__ == "__main__":
    main(*sys.argv[1:])
