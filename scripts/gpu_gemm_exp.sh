mkdir -p gpurun_out
rm -f gpurun_out/gemm_exp.log
for d in 0 1024 2; do
VITATK_GEMM_DBG=$d timeout 120 python scripts/gemm_bench.py 20 qkv,bproj,plain768 >> gpurun_out/gemm_exp.log 2>&1
done
cat gpurun_out/gemm_exp.log
