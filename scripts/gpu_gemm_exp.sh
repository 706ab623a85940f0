mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" > gpurun_out/t_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/t_gemm.log
rm -f gpurun_out/gemm_exp.log
timeout 120 python scripts/gemm_bench.py 20 fc1,plain768 >> gpurun_out/gemm_exp.log 2>&1
cat gpurun_out/gemm_exp.log
