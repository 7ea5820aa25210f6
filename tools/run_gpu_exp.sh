mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_single or tc_long" > gpurun_out/t_2cta.log 2>&1; echo "2cta unit rc=$?" >> gpurun_out/t_2cta.log
tail -n 6 gpurun_out/t_2cta.log
if grep -q "rc=0" gpurun_out/t_2cta.log; then
  timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
  tail -n 4 gpurun_out/t_gpu.log
  timeout -s KILL 600 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 6 2>&1 | tee gpurun_out/p_m20.log
  TNC_TC_2CTA=0 timeout -s KILL 600 python tools/gpu_probe.py n53_m20_sparse1024 --check --top 2 2>&1 | tee gpurun_out/p_m20_1cta.log
fi
