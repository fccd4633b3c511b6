"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / average / share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = row["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-72s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %5d %12.1f %10.1f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("total_us %.1f" % tot)


if __name__ == "__main__":
    main(sys.argv[1])
