mkdir -p gpurun_out
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
# launch list of one bench step (cold-cache, serialised: shares only)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-e2e --no-cpu-baseline --no-half > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
# full captures: the fat GEMM (2-CTA, 3xf16), the pack kernel of its A operand, one stem kernel
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_2cta -s 1 -c 1 -f -o gpurun_out/prof_gemm \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 2 -c 2 -f -o gpurun_out/prof_pack \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_pack.log 2>&1
echo "pack rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 40 -c 3 -f -o gpurun_out/prof_stem \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_stem.log 2>&1
echo "stem rc=$?"
ls -la gpurun_out/*.ncu-rep
