# GPU parity tests + smoke (fast check after a change)
mkdir -p gpurun_out
( time timeout -s KILL 1500 python -m pytest tests -q -m gpu -x -s ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/t_gpu.log
grep -E "relative error per amplitude|best-fit|passed|failed|rc=|^E " gpurun_out/t_gpu.log | tail -n 40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 6
