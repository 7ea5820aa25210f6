# both bench arms back to back on one box, the way the driver runs them (reference first)
mkdir -p gpurun_out
( time timeout -s KILL 1200 python bench.py --impl reference --steps ${1:-8} --warmup ${2:-3} ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
cut -c1-1500 gpurun_out/r02_bench_reference.json; tail -n 4 gpurun_out/r02_bench_reference.err
( time timeout -s KILL 900 python bench.py --steps ${1:-8} --warmup ${2:-3} ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "b200 rc=$?"
cut -c1-400 gpurun_out/r02_bench_n1.json; tail -n 4 gpurun_out/r02_bench_n1.err
