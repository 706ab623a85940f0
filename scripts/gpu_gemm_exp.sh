mkdir -p gpurun_out
rm -f gpurun_out/gemm_exp.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | grep -v Warning | tail -3
for gl in h2 p32; do
for e in 0 1; do
for d in 0 14; do
VITATK_GELU=$gl VITATK_GEMM_EPI16=$e VITATK_GEMM_DBG=$d timeout 120 python scripts/gemm_bench.py 30 fc1 2>&1 | tail -1 | sed "s/^/GELU=$gl EPI16=$e dbg=$d /" >> gpurun_out/gemm_exp.log
done
done
done
cat gpurun_out/gemm_exp.log
for cfg in "h2 0" "p32 0" "p32 1" "h2 0" "p32 0" "p32 1"; do
set -- $cfg
VITATK_GELU=$1 VITATK_GEMM_EPI16=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('GELU=$1 EPI16=$2 value',d['value'],'ms/step',d['ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k in ('qkv','proj','fc1','fc2','bfc2','bfc1','bqkv')})"
done
