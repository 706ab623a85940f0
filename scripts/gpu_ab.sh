mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warning | tail -3
for v in 0 1 0 1; do
VITATK_GEMM_RB=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('RB=$v value',d['value'],'ms/step',d['ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k.startswith('t_') or k.startswith('bt_')})"
done
