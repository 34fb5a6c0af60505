"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): launches and mean duration per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > 5:
        name = r[4].split("(")[0]
        try:
            d.setdefault(name, []).append(float(r[-1].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) / len(v) for v in d.values())
for k, v in d.items():
    print("%-40s launches %3d  mean %8.1f us  share %4.1f %%" % (k[:40], len(v), sum(v) / len(v) / 1e3, 100 * (sum(v) / len(v)) / tot))
