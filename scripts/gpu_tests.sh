set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
tail -30 gpurun_out/t_gemm.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "not gemm" > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"
tail -40 gpurun_out/t_kern.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q > gpurun_out/t_engine.log 2>&1; echo "engine rc=$?"
tail -60 gpurun_out/t_engine.log
