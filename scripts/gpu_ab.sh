mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/t_gpu.log
for v in 0 1 0 1; do
VITATK_ZIGZAG=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ZIGZAG=$v value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2), {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()}, d['robust']['clean_correct'], {k:v for k,v in d['breakdown_detail'].items() if k.startswith('t_') or k.startswith('bt_')})"
done
