"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list. usage: python devtools/launch_summary.py <launches.csv> "<command>" > summary.txt"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = collections.Counter(); n = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*$", "", r[4]); ns = float(r[-1].replace(",", "")) * (1e-6 if r[-2] in ("ns", "nsecond") else 1e-3 if r[-2] in ("us", "usecond") else 1.0)
    tot[name] += ns; n[name] += 1
T = sum(tot.values())
print(sys.argv[2] + " (cold-cache, serialised launches)")
for k, v in tot.most_common():
    print(f"{k[:100]:100s} launches {n[k]:4d}  total {v:10.3f} ms  share {100 * v / T:6.2f}%")
print(f"total {T:.3f} ms")
