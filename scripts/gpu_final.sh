mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"])
print(d["breakdown_ms_per_step"]); print(d["clocks"]); print(d.get("cpu_baseline"))
PY
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 7549 -c 504 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 40 -c 1 -o gpurun_out/prof_attn_bwd_fused -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu bwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc05 -s 40 -c 1 -o gpurun_out/prof_attn_fwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "ncu fwd rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
