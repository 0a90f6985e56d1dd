"""Print kernel name / grid / duration (us) for every row of an ncu --csv launch list."""
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = 0.0
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    us = float(r["Metric Value"].replace(",", "")) / 1e3
    tot += us
    print("%-34s %-18s %9.1f" % (r["Kernel Name"][:34].replace("void ", ""), r["Grid Size"], us))
print("total %.1f us" % tot)
