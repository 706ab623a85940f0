mkdir -p gpurun_out
timeout 120 python scripts/gemm_bench.py 3 fc1,proj > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 2 -c 1 -o gpurun_out/prof_fc1 -f python scripts/gemm_bench.py 3 fc1 > gpurun_out/ncu_fc1.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 2 -c 1 -o gpurun_out/prof_proj -f python scripts/gemm_bench.py 3 proj > gpurun_out/ncu_proj.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
