"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Grid Size")
tot, cnt = collections.defaultdict(float), collections.Counter()
verbose = len(sys.argv) > 2
for r in data:
    name = r[ki].split("(")[0][:70]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    tot[name] += v
    cnt[name] += 1
    if verbose:
        print(f"{name:72s} grid={r[gi]:>16s} {v:9.1f} us")
s = sum(tot.values())
print(f"{'kernel':72s} {'n':>4s} {'total us':>10s} {'share':>6s}")
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
    print(f"{k:72s} {cnt[k]:4d} {v:10.1f} {100 * v / s:5.1f}%")
print(f"{'sum':72s} {sum(cnt.values()):4d} {s:10.1f}")
