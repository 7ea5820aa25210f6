mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "skinny or folded or forced or n53 or n30 or pairs" ) > gpurun_out/t_gpu_2e.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_2e.log; tail -n 4 gpurun_out/t_gpu_2e.log
{
for cfg in "24 3 6 2,3,10,18,22,23" "24 3 5 1,11,17,21,26" "25 2 5 0,1,24,28,29" "24 4 4 16,18,26,27" "24 5 5 2,12,18,22,25" "25 5 3 0,11,16" "23 7 3 15,22,24" "22 6 6 3,9,12,20,25,27"; do
  set -- $cfg
  echo "== skinny $1 $2 $3 ka=$4: $(timeout 120 python tools/one_step.py $1 $2 $3 --algo skinny --ka $4 --reps 3 2>&1 | tail -n 1)"
done
} > gpurun_out/one_2e.log 2>&1
cat gpurun_out/one_2e.log
for c in n53_m20_sparse1024 n30_full; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 30 --tag _2e > gpurun_out/p_${c}_2e.log 2>&1; echo "$c $(sed -n 3,5p gpurun_out/p_${c}_2e.log | tr '\n' ' ')"
done
