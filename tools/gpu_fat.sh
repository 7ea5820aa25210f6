export TNC_EXPERIMENTS=1
python -m pytest tests/test_gpu_parity.py -q -x -k "3m or two_cta or n53_m20_one" 2>&1 | tail -n 2
python tools/one_step.py 15 13 15 --reps 4 2>&1 | tail -n 3
TNC_TC_3M=0 python tools/one_step.py 15 13 15 --reps 2 2>&1 | tail -n 1
