# gpurun --gpus 8: bench at 8 and 2 GPUs and the five BASELINE configurations at 8 GPUs (tools/run_configs.py).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
for N in 8 2; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err
  echo "N=$N rc=$?"; cut -c1-400 gpurun_out/bench_final_n$N.json; tail -n 2 gpurun_out/bench_final_n$N.err
done
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/run_configs.py 2> gpurun_out/configs_final_n8.err | tee gpurun_out/configs_final_n8.log | cut -c1-300
tail -n 3 gpurun_out/configs_final_n8.err
