# end of round: both bench arms on one box, then the bench lines of the trees of DESIGN.md 7.2 / 7.3
mkdir -p gpurun_out
bash tools/gpu_bench_both.sh ${1:-8} 3
for w in n53_m20_sparse1024_sc31_s20 n53_m20_sparse1024_sc32_s20; do
  timeout -s KILL 900 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-half --reuse-slices 1024 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']); print({k: d['slice_reuse'].get(k) for k in ('value','ms_per_slice_per_gpu','workspace_gib','keep_gib','steps_tied_to_their_reader','extrapolated_full_task_seconds','error')})" gpurun_out/r02_bench_$w.json
done
