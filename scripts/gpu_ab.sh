mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "attention or logits_and_grad or full_size" 2>&1 | tail -1
for v in 0 1 0 1; do
VITATK_ATTN_STREAM=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ab_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ATTN_STREAM=$v value',d['value'],'ms/step',d['ms_per_step'], {k:v[0] for k,v in d['breakdown_detail'].items() if k in ('attention_fwd','attention_bwd','t_proj','proj','bt_qkv','bqkv')})"
done
