mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/t_gpu.log
timeout 120 python scripts/attn_time.py 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"])
print(d["breakdown_ms_per_step"])
print({k:v for k,v in d["breakdown_detail"].items()})
print(d["clocks"]); print(d["robust"])
PY
tail -5 gpurun_out/bench.err
