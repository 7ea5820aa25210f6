mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 12 gpurun_out/t_gpu.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 40 > gpurun_out/p_m20_pack2.log 2>&1
head -n 46 gpurun_out/p_m20_pack2.log
