mkdir -p gpurun_out
for v in 0 1 0 1; do
VITATK_PDL=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PDL=$v value',d['value'],'ms/step',d['ms_per_step'], d['breakdown_ms_per_step'])"
done
