mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout -s KILL 1500 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
tail -n 40 gpurun_out/t_gpu.log
for P in 3xtf32 3xf16 f16; do
  timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 14 --precision $P --tag _$P > gpurun_out/p_m20_$P.log 2>&1
  echo "probe $P rc=$?"; head -n 12 gpurun_out/p_m20_$P.log
done
TNC_TC_KC=2 timeout -s KILL 400 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 6 --precision 3xf16 --tag _3xf16_kc2 > gpurun_out/p_m20_3xf16_kc2.log 2>&1
head -n 8 gpurun_out/p_m20_3xf16_kc2.log
TNC_TC_PRECISION=3xf16 TNC_TC_KC=1 timeout -s KILL 300 python tools/tc_calibrate.py > gpurun_out/calib_3xf16_kc1.log 2>&1
TNC_TC_PRECISION=3xf16 TNC_TC_KC=2 timeout -s KILL 300 python tools/tc_calibrate.py > gpurun_out/calib_3xf16_kc2.log 2>&1
tail -n 4 gpurun_out/calib_3xf16_kc1.log gpurun_out/calib_3xf16_kc2.log
