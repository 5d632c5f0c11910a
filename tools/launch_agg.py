"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (name, grid, block)."""
import collections, csv, sys

def main(path, top=20):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
        k = row["Kernel Name"][:72] + " " + row["Grid Size"] + " " + row["Block Size"]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"total {sum(a[1] for a in agg.values()):.1f} us in {sum(a[0] for a in agg.values())} launches")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[1]:9.1f} us {a[0]:4d}x {a[1] / a[0]:8.1f}  {k}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20)
