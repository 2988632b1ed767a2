"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / avg / total."""
import collections
import csv
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "ns")
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        name = row["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(t for _, t in agg.values())
    print(f"{'avg us':>10} {'count':>6} {'total ms':>10} {'share':>6}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t / n:10.1f} {n:6d} {t / 1000:10.3f} {t / tot:6.1%}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
