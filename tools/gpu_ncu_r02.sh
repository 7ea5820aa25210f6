# round 2 ncu evidence: (1) --set full of the 3M fat GEMM and of its pack / amax launches (one_step 15 13 15),
# (2) launch list of the bench command
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"gemm3m|pack2_kernel|amax_kernel" -s 4 -c 4 -f -o gpurun_out/r02_fat python tools/one_step.py 15 13 15 --reps 2 > gpurun_out/r02_ncu_fat.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-half --no-e2e --no-reuse > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches_bench.err
ls -la gpurun_out/r02_*
