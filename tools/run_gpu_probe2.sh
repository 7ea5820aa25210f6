mkdir -p gpurun_out
timeout -s KILL 900 python tools/gpu_probe.py n53_m12_sparse1024 --check --tc-min-flops 1e11 --top 12 > gpurun_out/p2_m12.log 2>&1; echo "rc=$?" >> gpurun_out/p2_m12.log
timeout -s KILL 900 python tools/gpu_probe.py n53_m20_sparse1024 --check --tc-min-flops 2e11 --top 30 > gpurun_out/p2_m20.log 2>&1; echo "rc=$?" >> gpurun_out/p2_m20.log
timeout -s KILL 900 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 3 > gpurun_out/p2_m20b.log 2>&1
cat gpurun_out/p2_m12.log gpurun_out/p2_m20.log gpurun_out/p2_m20b.log
