"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total ms, share.
usage: python tools/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_list.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
name, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    k = re.sub(r"\(.*", "", r[name])[:64]
    tot[k] += float(r[val].replace(",", "")) / 1e6
    cnt[k] += 1
total = sum(tot.values())
print("# kernel | launches | total ms | share")
for k, v in tot.most_common():
    print("%-66s %6d %12.3f %6.2f%%" % (k, cnt[k], v, 100 * v / total))
print("# total captured: %.1f ms over %d launches" % (total, len(rows)))
