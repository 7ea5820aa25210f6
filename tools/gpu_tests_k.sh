# a subset of the GPU tests: bash tools/gpu_tests_k.sh "<pytest -k expression>"
mkdir -p gpurun_out
( timeout -s KILL 1200 python -m pytest tests -q -m gpu -x -k "$1" ) > gpurun_out/t_k.log 2>&1; echo "tests rc=$?"; grep -E "^E |FAILED|passed|failed" gpurun_out/t_k.log | tail -n 12
