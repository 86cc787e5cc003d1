"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        name = row[ki].split("(")[0][:70]
        t = float(row[vi].replace(",", ""))
        if row[ui] == "ns":
            t /= 1000
        elif row[ui] == "ms":
            t *= 1000
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("%d launches, total %.1f us" % (sum(a[0] for a in agg.values()), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s n=%4d total=%10.1f us avg=%8.1f share=%5.1f%%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
