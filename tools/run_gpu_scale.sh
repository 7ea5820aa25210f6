mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 600 python bench.py --gpus 1 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_n1.json gpurun_out/bench_n$N.json gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench_n$N.err
