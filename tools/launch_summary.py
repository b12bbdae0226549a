"""Aggregate an ncu launch list (gpu__time_duration.sum csv) by kernel name."""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    print("# %d launches, %.3f ms total (serialised, cold-cache ncu replay: compare SHARES)" % (len(rows), tot*1e-6))
    for name, (n, ns) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-70s n=%4d  %9.3f ms  %5.1f%%" % (name[:70], n, ns*1e-6, 100*ns/tot))


if __name__ == "__main__":
    main()
