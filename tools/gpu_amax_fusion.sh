# amax fusion: the new parity test, then the same-box A/B (alternating) of the bench slice with and without it
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "amax_reduced or tc_3m or tc_single or skinny_single or stem_single or n53_m20_one_slice or cuda_graph or alternating" ) > gpurun_out/t_amax.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/t_amax.log
for i in 1 2; do
  for v in "" "--no-fuse-amax"; do
    timeout -s KILL 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-half --no-e2e $v > gpurun_out/amax_ab.json 2> gpurun_out/amax_ab.err
    python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/amax_ab.json'))
b=d['breakdown']
print(f"{sys.argv[1] or 'fused':>16}: {d['value']:.3f} slices/s  ms/slice {d['ms_per_step']/d['config']['slices_per_step_per_gpu']:.2f}  gemm {b['gemm_ms']:.2f} pack {b['pack_ms']:.2f} stem {b['stem_ms']:.2f} skinny {b['skinny_ms']:.2f}  launches {d['gpu_launches']}")
PY
  done
done
python tools/gpu_probe.py n53_m20_sparse1024 --check > gpurun_out/probe_fused.txt 2>&1; grep -E "check|slice total|by class|step  196|step  303|step  315|step  172" gpurun_out/probe_fused.txt
