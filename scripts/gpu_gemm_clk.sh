mkdir -p gpurun_out
for d in 0 16 20; do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 50 > gpurun_out/clk_$d.csv &
  SMI=$!
  VITATK_GEMM_DBG=$d timeout 120 python scripts/gemm_bench.py 3000 bfc1,plain768 2>&1 | tail -2
  kill $SMI
  sort -t, -k2 -n -r gpurun_out/clk_$d.csv | head -3
done
