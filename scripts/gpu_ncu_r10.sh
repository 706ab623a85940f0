mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 8059 -c 536 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches.csv
timeout 120 python scripts/gemm_bench.py 3 fc1,qkv,bfc2 > gpurun_out/plain2.log 2>&1 || exit 1
for n in fc1 qkv bfc2; do
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 2 -c 1 -o gpurun_out/prof_$n -f python scripts/gemm_bench.py 3 $n > gpurun_out/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
done
ls -la gpurun_out/*.ncu-rep
