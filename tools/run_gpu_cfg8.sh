mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/run_configs.py 2> gpurun_out/configs_n8.err | tee gpurun_out/configs_n8.log
tail -n 3 gpurun_out/configs_n8.err
timeout -s KILL 300 python tools/run_configs.py 2> gpurun_out/configs_n1.err | tee gpurun_out/configs_n1.log
