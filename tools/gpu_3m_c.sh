export TNC_EXPERIMENTS=1
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -x -k "3m" ) 2>&1 | tail -n 2
for d in 16 32; do for kc in 1 2; do
  echo "== 3M DRAIN=$d KC=$kc"; TNC_TC_DRAIN=$d TNC_TC_KC=$kc timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2
done; done
