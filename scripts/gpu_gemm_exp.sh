mkdir -p gpurun_out
rm -f gpurun_out/gemm_exp.log
for d in 0 512 0 512; do
VITATK_GEMM_DBG=$d timeout 120 python scripts/gemm_bench.py 20 qkv,proj,bproj,plain768,bfc1,bfc2 >> gpurun_out/gemm_exp.log 2>&1
done
cat gpurun_out/gemm_exp.log
