# CUDA-event time of the fat GEMM for several sweep-group sizes.
export TNC_EXPERIMENTS=1   # the TNC_* variant knobs below are only honoured with this (include/tnc_b200.h)
mkdir -p gpurun_out
{
for gm in 16 8 4 2 32 16; do
  echo "== fat gemm group_m=$gm: $(TNC_TC_GROUP_M=$gm timeout 300 python tools/one_step.py 15 13 15 --algo tc --reps 3 2>&1 | tail -n 2 | tr '\n' ' ')"
done
} > gpurun_out/one_2d.log 2>&1
cat gpurun_out/one_2d.log
