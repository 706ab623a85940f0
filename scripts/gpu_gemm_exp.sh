mkdir -p gpurun_out
rm -f gpurun_out/gemm_exp.log
for d in 4 16 2 1; do
  VITATK_GEMM_DBG=$d timeout 120 python scripts/gemm_bench.py 20 proj,qkv,fc1,fc2,bfc2,bfc1,plain768 >> gpurun_out/gemm_exp.log 2>&1; echo "dbg $d rc=$?"
done
cat gpurun_out/gemm_exp.log
