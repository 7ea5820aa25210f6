mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 8 gpurun_out/t_gpu.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 45 > gpurun_out/p_m20_skinny2.log 2>&1
head -n 52 gpurun_out/p_m20_skinny2.log
for S in "24 5 5" "26 3 2" "26 2 4" "25 3 3" "24 3 5"; do
  for A in skinny stem; do
    timeout -s KILL 200 python tools/one_step.py $S --algo $A --reps 3 2>&1 | tail -n 2
  done
done > gpurun_out/one_step_skinny.log 2>&1
cat gpurun_out/one_step_skinny.log
