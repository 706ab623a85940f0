"""Aggregate an ncu launch list per kernel: time, DRAM bytes and tensor-pipe utilisation per launch, plus the
time-weighted whole-run tensor-pipe utilisation and the hash of the sources the capture was taken on.

capture (one GPU, after the same command has exited 0 without ncu):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s <skip> -c <N> \
        --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
usage: ncu_launch_summary.py launches.csv [out.json]   (run from the repo root right after the capture)"""
import csv, json, os, re, sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
per = defaultdict(dict)
for r in rows:
    per[int(r[0])]["name"] = r[4]
    per[int(r[0])]["grid"] = r[8]
    per[int(r[0])][r[12]] = float(r[14].replace(",", ""))
agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "read_mb": 0.0, "write_mb": 0.0, "tensor_us": 0.0, "has_tensor": False})
for i in sorted(per):
    d = per[i]
    n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").strip()
    a = agg[n]
    us = d.get("gpu__time_duration.sum", 0) / 1e3
    a["launches"] += 1
    a["us"] += us
    a["read_mb"] += d.get("dram__bytes_read.sum", 0) / 1e6
    a["write_mb"] += d.get("dram__bytes_write.sum", 0) / 1e6
    if TENSOR in d:
        a["has_tensor"] = True
        a["tensor_us"] += us * d[TENSOR] / 100.0
tot = sum(a["us"] for a in agg.values())
tensor_tot = sum(a["tensor_us"] for a in agg.values())
have_tensor = any(a["has_tensor"] for a in agg.values())
out = {}
print(f"{'kernel':50s} {'n':>5s} {'us':>10s} {'share':>7s} {'rd MB/launch':>13s} {'wr MB/launch':>13s} {'tensor %':>9s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    tp = 100.0 * a["tensor_us"] / a["us"] if a["us"] and a["has_tensor"] else float("nan")
    print(f"{n[:50]:50s} {a['launches']:5d} {a['us']:10.1f} {a['us'] / tot:7.1%} {a['read_mb'] / a['launches']:13.1f} "
          f"{a['write_mb'] / a['launches']:13.1f} {tp:9.1f}")
    out[n] = {"launches": a["launches"], "us_total": round(a["us"], 1), "share": round(a["us"] / tot, 4),
              "dram_read_mb_per_launch": round(a["read_mb"] / a["launches"], 2),
              "dram_write_mb_per_launch": round(a["write_mb"] / a["launches"], 2)}
    if a["has_tensor"]:
        out[n]["tensor_pipe_pct"] = round(tp, 2)
print(f"total {tot:.1f} us over {sum(a['launches'] for a in agg.values())} launches")
if have_tensor:
    print(f"time-weighted tensor-pipe utilisation over all captured launches: {100.0 * tensor_tot / tot:.2f} % "
          f"({TENSOR}; per-launch times are cold-cache and serialised)")
if len(sys.argv) > 2:
    from bench import source_sha
    doc = {"total_us": round(tot, 1), "source_sha": source_sha(), "kernels": out}
    if have_tensor:
        doc["tensor_pipe_pct_time_weighted"] = round(100.0 * tensor_tot / tot, 2)
        doc["tensor_pipe_metric"] = TENSOR
    json.dump(doc, open(sys.argv[2], "w"), indent=1)
