#!/usr/bin/env python
"""Turn ncu output into the markdown tables committed under profiles/.

    python profiles/tools/ncu_summary.py rep  gpurun_out/x.ncu-rep [more.ncu-rep ...]   # --set full captures -> per-kernel table
    python profiles/tools/ncu_summary.py list gpurun_out/launches.csv                    # --metrics gpu__time_duration.sum list

`rep` averages every metric over the launches of a kernel name inside one report (ncu replays serialise launches and run
them cold: compare SHARES and byte counts, not absolute times)."""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("fr::", "").strip()


def rep(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr, units, body = rows[start], rows[start + 1], rows[start + 2:]
        kcol = hdr.index("Kernel Name")
        cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
        acc = collections.OrderedDict()
        for r in body:
            k = short(r[kcol])
            a = acc.setdefault(k, [0, [0.0] * len(cols)])
            a[0] += 1
            for j, (_, c) in enumerate(cols):
                try:
                    a[1][j] += float(r[c].replace(",", ""))
                except ValueError:
                    pass
        print(f"\nsource: `{p}`\n")
        for k, (n, sums) in acc.items():
            print(f"**{k}** ({n} launch{'es' if n > 1 else ''} averaged)\n")
            print("| metric | value | unit |\n|---|---|---|")
            for j, (m, c) in enumerate(cols):
                print(f"| {m} | {sums[j] / n:.6g} | {units[c]} |")
            print()


def launches(path):
    tot = collections.OrderedDict()
    with open(path) as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        hdr = next(rd)
        k, v, mname = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
        for r in rd:
            if r[mname] != "gpu__time_duration.sum":
                continue
            a = tot.setdefault(short(r[k]), [0, 0.0])
            a[0] += 1
            a[1] += float(r[v].replace(",", "")) / 1e3
    total = sum(a[1] for a in tot.values())
    print(f"total {total:.1f} us over {sum(a[0] for a in tot.values())} launches\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name[:110]} | {n} | {us:.1f} | {100 * us / total:.1f}% | {us / n:.2f} |")


if __name__ == "__main__":
    {"rep": rep, "list": lambda a: launches(a[0])}[sys.argv[1]](sys.argv[2:])
