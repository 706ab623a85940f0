mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"
tail -4 gpurun_out/t_attn.log
for d in 0 64 0 64; do VITATK_ATTN_DBG=$d timeout 120 python scripts/attn_time.py 2>&1 | tail -2; done
