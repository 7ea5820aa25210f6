mkdir -p gpurun_out
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
timeout -s KILL 300 python tools/permute_bench.py 2>&1 | tee gpurun_out/permute_bench.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-e2e --no-cpu-baseline --no-half > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
# full captures on one-step targets (second repetition of each)
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:amax_kernel|pack2_kernel|pack_kernel|gemm_2cta' -s 4 -c 4 -f -o gpurun_out/prof_fat \
    python tools/one_step.py 15 13 15 --algo tc --reps 2 > gpurun_out/ncu_fat.log 2>&1
echo "fat rc=$?"; tail -n 3 gpurun_out/ncu_fat.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:skinny_kernel -s 1 -c 1 -f -o gpurun_out/prof_skinny \
    python tools/one_step.py 24 5 5 --algo skinny --reps 2 --shuffle > gpurun_out/ncu_skinny.log 2>&1
echo "skinny rc=$?"; tail -n 2 gpurun_out/ncu_skinny.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 1 -c 1 -f -o gpurun_out/prof_stem \
    python tools/one_step.py 26 3 2 --algo stem --reps 2 --shuffle > gpurun_out/ncu_stem.log 2>&1
echo "stem rc=$?"; tail -n 2 gpurun_out/ncu_stem.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pack2_kernel -s 14 -c 1 -f -o gpurun_out/prof_permute \
    python tools/permute_bench.py --reps 1 > gpurun_out/ncu_permute.log 2>&1
echo "permute rc=$?"
ls -la gpurun_out/*.ncu-rep
