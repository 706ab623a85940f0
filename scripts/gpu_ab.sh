mkdir -p gpurun_out
for v in 0 1 0 1; do
VITATK_LN_STREAM=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('LN_STREAM=$v value',d['value'],'ms/step',d['ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k in ('layernorm_bwd','bt_fc2','bt_proj','bfc2','bproj')})"
done
