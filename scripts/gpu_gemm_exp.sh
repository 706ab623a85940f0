set -x
mkdir -p gpurun_out
for d in 0 1 2 4 8 16 6 18 20; do
  VITATK_GEMM_DBG=$d timeout 120 python scripts/gemm_bench.py 20 proj,qkv,fc1,fc2,bfc2,bfc1,plain768 >> gpurun_out/gemm_exp.log 2>&1; echo "dbg $d rc=$?"
done
cat gpurun_out/gemm_exp.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms/step",d["ms_per_step"])
print(d["breakdown_ms_per_step"])
for k,v in d["breakdown_detail"].items(): print(k, v)
PY
tail -5 gpurun_out/bench.err
