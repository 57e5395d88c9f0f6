#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/launches_summary.txt
ncu serialises the launches with cold caches, so only the SHARES are comparable with the live CUDA-event timing."""
import collections
import csv
import re
import sys


def main(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[k]).replace("b200trl::<unnamed>::", "").replace("void ", "")
        ns = float(r[v].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[u], 1.0)
        t = agg.setdefault(name, [0.0, 0])
        t[0] += ns
        t[1] += 1
    total = sum(t[0] for t in agg.values())
    print(f"ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none : {cmd}")
    print("every kernel launched inside the profiled range; ncu serialises launches with cold caches, so compare SHARES\n")
    for name, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ns / 1e3:12.1f} us total  {100 * ns / total:6.2f}%  x{n:3d}  {ns / n / 1e3:10.1f} us each  {name[:100]}")
    print(f"{total / 1e3:12.1f} us total")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
