mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/t_gpu.log
VITATK_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench0.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PDL=0 value',d['value'],'ms/step',d['ms_per_step'])"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"])
print(d["breakdown_ms_per_step"])
print({k:v for k,v in d["breakdown_detail"].items()})
print(d["clocks"]); print(d["robust"]); print(d["roofline"])
PY
tail -5 gpurun_out/bench.err
