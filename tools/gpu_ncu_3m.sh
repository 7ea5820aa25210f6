# ncu --set full of the 3M fat GEMM and of the 4M one (m15 n13 k15), one launch each
export TNC_EXPERIMENTS=1   # the TNC_* variant knobs below are only honoured with this (include/tnc_b200.h)
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:gemm3m -s 1 -c 1 -f -o gpurun_out/prof_3m python tools/one_step.py 15 13 15 --reps 2 > gpurun_out/ncu_3m.log 2>&1
TNC_TC_3M=0 timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:gemm_2cta -s 1 -c 1 -f -o gpurun_out/prof_4m python tools/one_step.py 15 13 15 --reps 2 > gpurun_out/ncu_4m.log 2>&1
ls -la gpurun_out/*.ncu-rep
