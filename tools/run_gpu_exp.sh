mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 25 gpurun_out/t_gpu.log
timeout -s KILL 600 python tools/gpu_probe.py n30_sparse10000 --check --top 12 2>&1 | tee gpurun_out/p_n30s.log
timeout -s KILL 600 python tools/gpu_probe.py n30_full --top 8 2>&1 | tee gpurun_out/p_n30f.log
