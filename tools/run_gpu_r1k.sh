mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 6 gpurun_out/t_gpu.log
timeout -s KILL 200 python tools/one_step.py 24 3 6 --algo skinny --reps 3 --shuffle 2>&1 | tail -n 1
timeout -s KILL 200 python tools/one_step.py 21 5 6 --algo skinny --reps 3 --shuffle 2>&1 | tail -n 1
timeout -s KILL 900 python tools/run_configs.py 2> gpurun_out/configs_n1.err | tee gpurun_out/configs_n1.log
tail -n 3 gpurun_out/configs_n1.err
