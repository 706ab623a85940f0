"""Aggregate an ncu launch list (gpu__time_duration + dram bytes per launch) per kernel.
usage: ncu_launch_summary.py launches.csv [out.json]"""
import csv, json, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
per = defaultdict(dict)
for r in rows:
    per[int(r[0])]["name"] = r[4]
    per[int(r[0])]["grid"] = r[8]
    per[int(r[0])][r[12]] = float(r[14].replace(",", ""))
agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "read_mb": 0.0, "write_mb": 0.0})
for i in sorted(per):
    d = per[i]
    n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").strip()
    a = agg[n]
    a["launches"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0) / 1e3
    a["read_mb"] += d.get("dram__bytes_read.sum", 0) / 1e6
    a["write_mb"] += d.get("dram__bytes_write.sum", 0) / 1e6
tot = sum(a["us"] for a in agg.values())
out = {}
print(f"{'kernel':50s} {'n':>5s} {'us':>10s} {'share':>7s} {'rd MB/launch':>13s} {'wr MB/launch':>13s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{n[:50]:50s} {a['launches']:5d} {a['us']:10.1f} {a['us'] / tot:7.1%} {a['read_mb'] / a['launches']:13.1f} {a['write_mb'] / a['launches']:13.1f}")
    out[n] = {"launches": a["launches"], "us_total": round(a["us"], 1), "share": round(a["us"] / tot, 4),
              "dram_read_mb_per_launch": round(a["read_mb"] / a["launches"], 2),
              "dram_write_mb_per_launch": round(a["write_mb"] / a["launches"], 2)}
print(f"total {tot:.1f} us over {sum(a['launches'] for a in agg.values())} launches")
if len(sys.argv) > 2:
    json.dump({"total_us": round(tot, 1), "kernels": out}, open(sys.argv[2], "w"), indent=1)
