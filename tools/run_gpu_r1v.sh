mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu -x -k "own or stem_bulk" ) > gpurun_out/t_gpu_v.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_v.log; tail -n 3 gpurun_out/t_gpu_v.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:stem_bulk -s 2 -c 1 -f -o gpurun_out/prof_stem_bulk \
    python tools/one_step.py 25 3 3 --algo stem --ka 21,22,25 --reps 3 > gpurun_out/ncu_stem_bulk.log 2>&1
echo "ncu stem_bulk rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:skinny -s 2 -c 1 -f -o gpurun_out/prof_skinny_n3k6 \
    python tools/one_step.py 24 3 6 --algo skinny --ka 2,3,10,18,22,23 --reps 3 > gpurun_out/ncu_skinny.log 2>&1
echo "ncu skinny rc=$?"
ls -la gpurun_out/*.ncu-rep
