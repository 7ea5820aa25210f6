# one quick bench line (no CPU baseline) + the slice_reuse key
mkdir -p gpurun_out
( timeout -s KILL 900 python bench.py --steps ${1:-4} --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"
tail -n 3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','fused_amax_operands')}, d['e2e'])
print(json.dumps(d['slice_reuse'], indent=1))
PY
