"""Kernel shares from an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_shares.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.Counter()
cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0][-42:]
    tot[name] += int(r[14])
    cnt[name] += 1
total = sum(tot.values())
for k, v in tot.most_common():
    print("%-42s n=%4d total_ns=%14d share=%.3f" % (k, cnt[k], v, v / total))
