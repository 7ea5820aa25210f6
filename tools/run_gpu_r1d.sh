mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 25 gpurun_out/t_gpu.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 30 > gpurun_out/p_m20_skinny.log 2>&1
head -n 40 gpurun_out/p_m20_skinny.log
for S in "24 5 5" "25 5 3" "26 3 2" "26 2 4" "24 3 5" "23 7 3"; do
  for A in skinny stem; do
    timeout -s KILL 200 python tools/one_step.py $S --algo $A --reps 3 2>&1 | tail -n 2
  done
done > gpurun_out/one_step_skinny.log 2>&1
cat gpurun_out/one_step_skinny.log
