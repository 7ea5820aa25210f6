# 3M complex product: parity tests, then the fat GEMM of the n53 m20 tree (m15 n13 k15) timed with and without it
export TNC_EXPERIMENTS=1   # the TNC_* variant knobs below are only honoured with this (include/tnc_b200.h)
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "3m or two_cta or long_contraction or tc_single" ) > gpurun_out/t_3m.log 2>&1; echo "rc=$?" >> gpurun_out/t_3m.log
grep -E "3M max|passed|failed|rc=|Error|error" gpurun_out/t_3m.log | tail -n 30
for m in 1 0; do
  echo "== TNC_TC_3M=$m 3xf16"; TNC_TC_3M=$m timeout -s KILL 300 python tools/one_step.py 15 13 15 --reps 4 2>&1 | tail -n 4
done
echo "== 3M f16"; timeout -s KILL 300 python tools/one_step.py 15 13 15 --precision f16 --reps 3 2>&1 | tail -n 3
echo "== 4M f16"; TNC_TC_3M=0 timeout -s KILL 300 python tools/one_step.py 15 13 15 --precision f16 --reps 3 2>&1 | tail -n 3
echo "== short K 3M / 4M"; timeout -s KILL 300 python tools/one_step.py 21 9 7 --reps 3 2>&1 | tail -n 2; TNC_TC_3M=0 timeout -s KILL 300 python tools/one_step.py 21 9 7 --reps 3 2>&1 | tail -n 2
