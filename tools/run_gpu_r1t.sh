mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_gpu_t.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu_t.log
tail -n 8 gpurun_out/t_gpu_t.log
