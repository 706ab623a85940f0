mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | grep -v Warning | tail -3
timeout 200 python scripts/gemm_bench.py 30 2>&1 | tail -11
for v in 1 1; do
VITATK_GEMM_EPI16=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('EPI16=$v value',d['value'],'ms/step',d['ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k in ('qkv','proj','fc1','fc2','bfc2','bfc1','bproj','bqkv')})"
done
