# short-K / narrow-N tensor-core steps: 3M panels (12 B per amplitude, 2.25 products) against 4M panels (8 B, 3 products)
export TNC_EXPERIMENTS=1
for sh in "21 7 8" "19 8 8" "21 9 7" "18 7 7" "12 7 12"; do
  for m in 1 0; do echo "== $sh 3M=$m: $(TNC_TC_3M=$m timeout 300 python tools/one_step.py $sh --shuffle --reps 3 2>&1 | tail -n 1 | cut -c1-110)"; done
done
