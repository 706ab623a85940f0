mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 6 -c 1 -o gpurun_out/prof_engine_ew2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_engine_ew2.ncu-rep
