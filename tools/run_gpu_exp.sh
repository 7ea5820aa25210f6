mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 4 gpurun_out/t_gpu.log
timeout -s KILL 600 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 12 2>&1 | tee gpurun_out/p_m20.log
timeout -s KILL 600 python tools/gpu_probe.py n53_m12_sparse1024 --check --top 6 2>&1 | tee gpurun_out/p_m12.log
