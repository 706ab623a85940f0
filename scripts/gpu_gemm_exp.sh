mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" > gpurun_out/t_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -12 gpurun_out/t_gemm.log
rm -f gpurun_out/gemm_exp.log
for d in 0 2; do
VITATK_GEMM_DBG=$d VITATK_GEMM_2CTA=1 timeout 120 python scripts/gemm_bench.py 20 proj,qkv,fc1,fc2,bfc2,bfc1,bproj,bqkv,plain768 >> gpurun_out/gemm_exp.log 2>&1; echo "rc=$?"
done
cat gpurun_out/gemm_exp.log
