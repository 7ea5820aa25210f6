# GPU parity tests + a bench run without the CPU baseline: the quick check after a kernel change.
mkdir -p gpurun_out
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_gpu_re.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu_re.log
tail -n 6 gpurun_out/t_gpu_re.log
( timeout -s KILL 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_re.json 2> gpurun_out/bench_re.err; cut -c1-200 gpurun_out/bench_re.json
# NVTX ranges: the smoke contraction with TNC_NVTX=1 must behave exactly as without
TNC_NVTX=1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
