mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm3xtf32 -s 5 -c 1 -f -o gpurun_out/prof_gemm \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 40 -c 1 -f -o gpurun_out/prof_stem \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_stem.log 2>&1
echo "stem rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 10 -c 1 -f -o gpurun_out/prof_pack \
    python tools/gpu_probe.py n53_m20_sparse1024 --top 3 > gpurun_out/ncu_pack.log 2>&1
echo "pack rc=$?"
