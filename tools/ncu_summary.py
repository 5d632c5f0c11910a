#!/usr/bin/env python
"""Summarise ncu output for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel time shares (markdown)
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep                -> key metrics of every captured launch

`launches` reads the CSV written by
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> <cmd>
`full` shells out to `ncu -i <rep> --page raw --csv` (works without a GPU).
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

FULL_KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__cluster_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "sm__cycles_elapsed.avg",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("usf::", "")
    return name.strip()


def launches(path: str) -> None:
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block, 0.0])
        a[0] += 1
        a[1] += ns
        if ns > a[4]:                      # report the launch geometry of the longest launch
            a[2], a[3], a[4] = grid, block, ns
    total = sum(a[1] for a in agg.values())
    print(f"launches: {len(rows)}, total device time {total / 1e6:.3f} ms (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | share | avg us | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, (n, ns, grid, block, _) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} | {grid} | {block} |")


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"### `{short(r[hdr.index('Kernel Name')])}`  (launch id {r[0]})\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("launches", "full"):
        sys.exit(__doc__)
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
