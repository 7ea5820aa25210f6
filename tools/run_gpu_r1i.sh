mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 12 gpurun_out/t_gpu.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 16 > gpurun_out/p_m20_i.log 2>&1
head -n 22 gpurun_out/p_m20_i.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m12_sparse1024 --check --top 8 > gpurun_out/p_m12_i.log 2>&1
head -n 14 gpurun_out/p_m12_i.log
timeout -s KILL 400 python tools/gpu_probe.py n30_full --top 4 > gpurun_out/p_n30f_i.log 2>&1
head -n 10 gpurun_out/p_n30f_i.log
