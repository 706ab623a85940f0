mkdir -p gpurun_out
for v in 0 1 0 1; do
VITATK_FUSE_LN_T=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('FUSE_LN_T=$v value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2), {k:round(v,1) for k,v in d['breakdown_ms_per_step'].items()})"
done
