mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 24 -c 2 -o gpurun_out/prof_attn_bwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc05 -s 24 -c 1 -o gpurun_out/prof_attn_fwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 400 -c 16 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep
