"""Per-(kernel, grid) table of an `ncu --csv --metrics gpu__time_duration.sum,...` launch list: launches, us each, total, MB/launch."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, mi, vi, gi, ii = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Grid Size", "ID"))
t, n, by = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    k = r[ki].split("(")[0][-44:] + " " + r[gi]
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        t[k] += v
        n[k] += 1
    elif r[mi].startswith("dram__bytes"):
        by[k] += v
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{k:72s} {n[k]:4d} {v / 1e3 / n[k]:8.1f} us each {v / 1e3:9.1f} us {100 * v / tot:5.1f}%  {by[k] / max(n[k], 1):10.1f} (dram units)/launch")
print(f"total {tot / 1e6:.2f} ms over {sum(n.values())} launches")
