mkdir -p gpurun_out
for c in n53_m20_sparse1024 n53_m12_sparse1024 n30_sparse10000 n30_full; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 45 > gpurun_out/p_${c}_o.log 2>&1
  head -n 5 gpurun_out/p_${c}_o.log
done
