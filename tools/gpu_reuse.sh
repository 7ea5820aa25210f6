# slice reuse: parity tests, then measurements
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "slice_reuse" ) > gpurun_out/t_reuse.log 2>&1; echo "tests rc=$?"; grep -E "^E |FAILED|passed|failed" gpurun_out/t_reuse.log | tail -n 12
for c in n53_m20_sparse1024 n53_m12_sparse1024 n53_m20_sparse1024_sc31; do
  timeout -s KILL 900 python tools/reuse_bench.py $c --ranges 2,8,64,512 2>&1 | tee gpurun_out/reuse_$c.txt | tail -n 16
done
