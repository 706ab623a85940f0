mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
wc -l gpurun_out/bench_default.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"], "steps", d["steps"], d["warmup"])
print(d["breakdown_ms_per_step"]); print(d["clocks"]); print(d["roofline"]); print(d["cpu_baseline"]); print(d.get("gpu_launches"))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
