mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | grep -v Warning | tail -2
timeout 300 python scripts/tc_const_check.py 2>&1 | tail -3
for v in 1 1; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms/step',d['ms_per_step'], d['breakdown_ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k in ('qkv','proj','fc1','fc2','t_qkv','t_proj','t_fc1','t_fc2')})"
done
