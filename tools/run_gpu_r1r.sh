mkdir -p gpurun_out
for c in n53_m20_sparse1024 n30_sparse10000; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 45 > gpurun_out/p_${c}_r.log 2>&1
  head -n 5 gpurun_out/p_${c}_r.log | tail -n 3
done
grep "algo=2" gpurun_out/p_n53_m20_sparse1024_r.log | head -8 | cut -c1-170
