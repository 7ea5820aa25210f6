mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_gpu_2a.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu_2a.log
tail -n 8 gpurun_out/t_gpu_2a.log
for c in n53_m12_sparse1024 n53_m20_sparse1024; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 12 --tag _pairs > gpurun_out/p_${c}_pairs.log 2>&1
  sed -n 2,17p gpurun_out/p_${c}_pairs.log | cut -c1-200
done
timeout -s KILL 400 python tools/run_configs.py 2> gpurun_out/configs_n1_2a.err | tee gpurun_out/configs_n1_2a.log | cut -c1-330
