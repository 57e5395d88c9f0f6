#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key roofline metrics -> JSON, hot SASS lines -> text.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>
writes <name>_summary.json and <name>_hot_sass.txt (and prints the summary)."""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, prefix):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    summary = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
                d[h] = f"{vals[i]} {units[i]}".strip()
        rd, wr = d.get("dram__bytes_read.sum", "0 x").split(), d.get("dram__bytes_write.sum", "0 x").split()
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        d["dram_bytes_per_launch"] = float(rd[0]) * scale.get(rd[1], 1) + float(wr[0]) * scale.get(wr[1], 1)
        summary.append(d)
    json.dump(summary, open(prefix + "_summary.json", "w"), indent=1)
    print(json.dumps(summary, indent=1))
    src = page(rep, "source")
    h = src[1]
    isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = [r for r in src[2:] if len(r) > ismp and r[iex].isdigit()]
    tot_s = sum(int(r[ismp]) for r in data if r[ismp].isdigit()) or 1
    tot_i = sum(int(r[iex]) for r in data)
    with open(prefix + "_hot_sass.txt", "w") as f:
        f.write(f"total warp instructions {tot_i}, stall samples {tot_s}\n\ninstruction groups by execution count:\n")
        grp = collections.defaultdict(lambda: [0, 0])
        for r in data:
            g = grp[int(r[iex])]
            g[0] += 1
            g[1] += int(r[ismp]) if r[ismp].isdigit() else 0
        for c, (n, s) in sorted(grp.items(), key=lambda kv: -kv[0] * kv[1][0])[:14]:
            f.write(f"  executed {c:>10d} x {n:4d} instrs = {c * n / 1e6:8.1f} M warp-inst, {100 * s / tot_s:5.1f}% of samples\n")
        f.write("\ntop stall-sample instructions:\n")
        for r in sorted(data, key=lambda r: -(int(r[ismp]) if r[ismp].isdigit() else 0))[:40]:
            f.write(f"  {int(r[ismp]):7d} samples  exec {r[iex]:>10s}  {r[isrc][:110]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
