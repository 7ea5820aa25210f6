mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 8 gpurun_out/t_gpu.log
timeout -s KILL 400 python tools/gpu_probe.py n30_sparse10000 --check --top 4 > gpurun_out/p_n30s_m.log 2>&1
head -n 6 gpurun_out/p_n30s_m.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m12_sparse1024 --check --top 6 > gpurun_out/p_m12_m.log 2>&1
head -n 12 gpurun_out/p_m12_m.log
