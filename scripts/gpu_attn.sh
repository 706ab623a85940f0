mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"
tail -25 gpurun_out/t_attn.log
