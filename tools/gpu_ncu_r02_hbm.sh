# round 2, ncu --set full of the HBM-bound kernels in their final form: bulk-copy fp32 streaming kernel (m25 n3 k3),
# streaming tcgen05 kernel (m23 n5 k5), pack (copy mode is exercised by one_step's tc path: m21 n7 k8 = step 303 of n53 m20)
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:stem_bulk -s 1 -c 1 -f -o gpurun_out/r02_stem_bulk python tools/one_step.py 25 3 3 --algo stem --shuffle --reps 2 > gpurun_out/r02_ncu_stem.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:skinny -s 1 -c 1 -f -o gpurun_out/r02_skinny_k5n5b python tools/one_step.py 23 5 5 --algo skinny --shuffle --reps 2 > gpurun_out/r02_ncu_skinny.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"pack2|gemm_2cta" -s 3 -c 3 -f -o gpurun_out/r02_step303 python tools/one_step.py 21 7 8 --reps 2 > gpurun_out/r02_ncu_303.log 2>&1
for f in r02_stem_bulk r02_skinny_k5n5b r02_step303; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt; head -n 6 gpurun_out/$f.summary.txt; done
