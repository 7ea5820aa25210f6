# trees picked by the reuse-aware pricing of the order-search sweep: reuse measurement, then one slice against the CPU oracle
# (complex64 and complex128) on the box's host
mkdir -p gpurun_out
for c in "$@"; do
  timeout -s KILL 900 python tools/reuse_bench.py $c --ranges 2,8,64,512,2048 --check 2 2>&1 | tee gpurun_out/reuse_$c.txt | tail -n 18
done
for c in "$@"; do
  timeout -s KILL 1800 python tools/check_vs_oracle_on_box.py $c --c128 2>&1 | tail -n 1 | cut -c1-1500
done
