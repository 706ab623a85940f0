set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/t_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"])
print("roofline",d["roofline"]["achieved"],d["roofline"]["frac"])
print(d["breakdown_ms_per_step"]); print(d["clocks"]); print(d["robust"])
PY
tail -5 gpurun_out/bench.err
