mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_gpu_y.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu_y.log
tail -n 8 gpurun_out/t_gpu_y.log
for c in n53_m12_sparse1024 n53_m20_sparse1024; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 20 --tag _chain > gpurun_out/p_${c}_chain.log 2>&1
  sed -n 3,5p gpurun_out/p_${c}_chain.log
  TNC_NO_CHAIN=1 timeout -s KILL 400 python tools/gpu_probe.py $c --top 20 --tag _nochain > gpurun_out/p_${c}_nochain.log 2>&1
  sed -n 3,4p gpurun_out/p_${c}_nochain.log
done
timeout -s KILL 400 python tools/run_configs.py 2> gpurun_out/configs_n1_y.err | tee gpurun_out/configs_n1_y.log | cut -c1-330
