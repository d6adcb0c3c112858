"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py profiles/<file>.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.OrderedDict(), collections.Counter()
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0][-48:]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    tot[k] = tot.get(k, 0.0) + v
    cnt[k] += 1
s = sum(tot.values())
print("%-50s %5s %12s %7s" % ("kernel", "n", "total us", "share"))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-50s %5d %12.1f %6.1f%%" % (k, cnt[k], v, 100 * v / s))
