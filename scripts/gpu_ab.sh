mkdir -p gpurun_out
for v in 1024 2304 4096 1024 4096; do
VITATK_GEMM_EPI16_MAXK=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MAXK=$v value',d['value'],'ms/step',d['ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k in ('fc2','bfc1','bqkv','qkv','fc1')})"
done
