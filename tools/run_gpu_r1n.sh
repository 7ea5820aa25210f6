mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_gpu_n.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu_n.log
tail -n 8 gpurun_out/t_gpu_n.log
( time timeout -s KILL 600 python bench.py ) > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; tail -c 3000 gpurun_out/bench_n.json
