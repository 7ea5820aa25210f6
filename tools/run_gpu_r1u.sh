mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests -q -m gpu -x -k "stem" ) > gpurun_out/t_gpu_u.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_u.log; tail -n 3 gpurun_out/t_gpu_u.log
timeout -s KILL 400 python tools/run_configs.py 2> gpurun_out/configs_n1_u.err | tee gpurun_out/configs_n1_u.log
( timeout -s KILL 600 python bench.py ) > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; tail -c 1500 gpurun_out/bench_u.json
