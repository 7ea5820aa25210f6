mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 15 gpurun_out/t_gpu.log; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
