mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "skinny or tc or n53 or n30 or forced" ) > gpurun_out/t_gpu_2b.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_2b.log; tail -n 4 gpurun_out/t_gpu_2b.log
{
for cfg in "24 3 6 2,3,10,18,22,23" "24 3 5 1,11,17,21,26" "25 2 5 0,1,24,28,29" "24 4 4 16,18,26,27" "24 5 5 2,12,18,22,25" "25 5 3 0,11,16"; do
  set -- $cfg
  echo "== skinny $1 $2 $3 ka=$4: $(timeout 120 python tools/one_step.py $1 $2 $3 --algo skinny --ka $4 --reps 3 2>&1 | tail -n 1)"
done
echo "== fat gemm 15 13 15: $(timeout 300 python tools/one_step.py 15 13 15 --algo tc --reps 3 2>&1 | tail -n 2 | tr '\n' ' ')"
} > gpurun_out/one_2b.log 2>&1
cat gpurun_out/one_2b.log
timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 14 > gpurun_out/p_m20_2b.log 2>&1; sed -n 3,20p gpurun_out/p_m20_2b.log | cut -c1-180
