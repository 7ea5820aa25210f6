mkdir -p gpurun_out
export TNC_EXPERIMENTS=1   # the TNC_* variant knobs below are only honoured with this (include/tnc_b200.h)
( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "3m or two_cta or long_contraction" ) > gpurun_out/t_3m.log 2>&1; echo "rc=$?" >> gpurun_out/t_3m.log
grep -E "3M max|passed|failed|rc=|Error|error" gpurun_out/t_3m.log | tail -n 12
for kc in 1 2 4; do
  echo "== 3M KC=$kc"; TNC_TC_KC=$kc timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2
done
echo "== 4M"; TNC_TC_3M=0 timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2
echo "== 3M no lockstep"; TNC_TC_SYNC=0 timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2
echo "== 3M group_m 4"; TNC_TC_GROUP_M=4 timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2
echo "== 3M group_m 16"; TNC_TC_GROUP_M=16 timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2
echo "== n30 fat 14 14 12: 3M / 4M"; timeout -s KILL 300 python tools/one_step.py 14 14 12 --reps 3 2>&1 | tail -n 2; TNC_TC_3M=0 timeout -s KILL 300 python tools/one_step.py 14 14 12 --reps 3 2>&1 | tail -n 2
