"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/summarize_launches.py f.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in data:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else v
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T / 1e3:.2f} ms over {sum(cnt.values())} launches (serialised, cold-cache: compare SHARES)")
print(f"{'ms':>9} {'share':>6} {'n':>5} {'avg us':>9}  kernel")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v / 1e3:9.3f} {100 * v / T:5.1f}% {cnt[k]:5d} {v / cnt[k]:9.1f}  {k[:100]}")
