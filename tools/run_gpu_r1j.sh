mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 6 gpurun_out/t_gpu.log
for S in "24 5 5" "24 4 4" "25 5 3" "24 3 6" "23 7 3" "24 6 2" "22 6 6"; do
  timeout -s KILL 200 python tools/one_step.py $S --algo skinny --reps 3 --shuffle 2>&1 | tail -n 1
done > gpurun_out/one_step_skinny3.log 2>&1
cat gpurun_out/one_step_skinny3.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 12 > gpurun_out/p_m20_j.log 2>&1
head -n 18 gpurun_out/p_m20_j.log
timeout -s KILL 400 python tools/gpu_probe.py n30_full --top 6 > gpurun_out/p_n30f_j.log 2>&1
head -n 12 gpurun_out/p_n30f_j.log
