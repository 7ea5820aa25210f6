# same-box A/B of 3M variants on the fat GEMM of n53 m20 (box-to-box spread is +-5 %: only alternating runs on ONE box decide)
export TNC_EXPERIMENTS=1
python -m pytest tests/test_gpu_parity.py -q -x -s -k "3m_complex or two_cta or n53_m20_one" 2>&1 | grep -E "3M max|passed|failed" | tail -n 7
for r in 1 2 3; do for v in 1 0; do echo "== GAUSS=$v"; TNC_TC_GAUSS=$v python tools/one_step.py 15 13 15 --reps 3 2>&1 | tail -n 2 | cut -c1-90; done; done
