# multi-GPU checks (gpurun --gpus N): bench at N, the sharded full-amplitude config, optionally the five configurations
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8 > gpurun_out/gpus_n$N.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench N=$N rc=$?"; cut -c1-300 gpurun_out/r02_bench_n$N.json; tail -n 2 gpurun_out/r02_bench_n$N.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    tools/run_sharded.py 2> gpurun_out/r02_sharded_n$N.err | tail -n 1
tail -n 2 gpurun_out/r02_sharded_n$N.err
if [ "$2" = "configs" ]; then
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) tools/run_configs.py 2> gpurun_out/r02_configs_n$N.err | tee gpurun_out/r02_configs_n$N.log | cut -c1-330
  tail -n 2 gpurun_out/r02_configs_n$N.err
fi
