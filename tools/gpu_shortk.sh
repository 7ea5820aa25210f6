# short-K tensor-core steps of n53 m20 (steps 315 / 303 / 172 / 422 shapes): accumulation chunk length
export TNC_EXPERIMENTS=1
for sh in "21 9 7" "19 8 8" "21 7 8"; do
  for kc in 1 2 4; do echo "== $sh KC=$kc: $(TNC_TC_KC=$kc timeout 300 python tools/one_step.py $sh --shuffle --reps 3 2>&1 | tail -n 1 | cut -c1-110)"; done
done
