# One gpurun call at the end of a work session: GPU parity tests, smoke, bench (both arms), ncu launch list of the
# bench command, per-step probes of four configurations.  Outputs land in gpurun_out/; copy what is to be kept into profiles/.
mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_gpu_final.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu_final.log
tail -n 6 gpurun_out/t_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
( time timeout -s KILL 600 python bench.py ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -n 4 gpurun_out/bench_final.err
( time timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 700 gpurun_out/bench_ref_final.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-e2e --no-cpu-baseline --no-half > gpurun_out/ncu_list_final.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_final.csv
for c in n53_m20_sparse1024 n53_m12_sparse1024 n30_sparse10000 n30_full; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 45 > gpurun_out/p_${c}_final.log 2>&1
  sed -n 3,5p gpurun_out/p_${c}_final.log
done
