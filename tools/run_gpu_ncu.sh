mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm3xtf32_kernelILi256 -s 3 -c 1 -f -o gpurun_out/prof_gemm \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 24 -c 4 -f -o gpurun_out/prof_stem \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_stem.log 2>&1
echo "stem rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 6 -c 2 -f -o gpurun_out/prof_pack \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_pack.log 2>&1
echo "pack rc=$?"
ls -la gpurun_out/
