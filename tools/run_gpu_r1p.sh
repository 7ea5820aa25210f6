mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "stem or folded or forced or n30 or n53" ) > gpurun_out/t_gpu_p.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_p.log
tail -n 15 gpurun_out/t_gpu_p.log
for c in n53_m20_sparse1024 n30_sparse10000 n30_full; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 45 > gpurun_out/p_${c}_p.log 2>&1
  head -n 5 gpurun_out/p_${c}_p.log | tail -n 3
done
