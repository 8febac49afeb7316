"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", d["Kernel Name"]))[:72]
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print("total %.1f us in %d launches" % (tot, n))
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%9.1f us %5.1f%% %5d  %s" % (v, 100 * v / tot, c, k))


if __name__ == "__main__":
    main(sys.argv[1])
