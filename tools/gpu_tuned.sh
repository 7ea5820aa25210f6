# tuned trees of the order-search sweep: parity tests, oracle check of the sc32 slice on the box's host, bench lines
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "tuned_tree" 2>&1 | tail -n 3
for w in n53_m20_sparse1024_sc31 n53_m20_sparse1024_sc32; do
  timeout -s KILL 600 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-half > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  echo "$w rc=$?"; cut -c1-260 gpurun_out/r02_bench_$w.json; tail -n 2 gpurun_out/r02_bench_$w.err
done
timeout -s KILL 1500 python tools/check_vs_oracle_on_box.py n53_m20_sparse1024_sc32 2>&1 | tail -n 2
timeout -s KILL 900 python tools/check_vs_oracle_on_box.py n53_m20_sparse1024_sc31 2>&1 | tail -n 1
