# end-of-round measurement on one box: both bench arms, the GPU tests + smoke, the ncu launch list of the bench command
mkdir -p gpurun_out
bash tools/gpu_bench_both.sh ${1:-8} 3
bash tools/gpu_tests.sh
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-half --no-e2e --no-reuse > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches_bench.err
python tools/launch_list_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_list.txt; head -n 8 gpurun_out/r02_launch_list.txt
for w in n53_m20_sparse1024_sc31 n53_m20_sparse1024_sc32; do
  timeout -s KILL 600 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-half > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"; cut -c1-200 gpurun_out/r02_bench_$w.json
done
