mkdir -p gpurun_out
{
for ka in 21,22,25 13,16,17 4,5,6 0,1,2; do
  echo "== stem 25 3 3 ka=$ka"; python tools/one_step.py 25 3 3 --algo stem --ka $ka --reps 3 | tail -n 2
done
echo "== old kernel 25 3 3 ka=21,22,25"; TNC_STEM_NO_BULK=1 python tools/one_step.py 25 3 3 --algo stem --ka 21,22,25 --reps 3 | tail -n 2
for ka in 11,12,14,28 3,4,5,6; do
  echo "== stem 26 2 4 ka=$ka"; python tools/one_step.py 26 2 4 --algo stem --ka $ka --reps 3 | tail -n 2
done
echo "== old kernel 26 2 4"; TNC_STEM_NO_BULK=1 python tools/one_step.py 26 2 4 --algo stem --ka 11,12,14,28 --reps 3 | tail -n 2
echo "== stem 26 3 2 ka=9,10"; python tools/one_step.py 26 3 2 --algo stem --ka 9,10 --reps 3 | tail -n 2
echo "== stem 26 2 2 ka=10,26"; python tools/one_step.py 26 2 2 --algo stem --ka 10,26 --reps 3 | tail -n 2
echo "== skinny 24 3 6"; python tools/one_step.py 24 3 6 --algo skinny --ka 2,3,10,18,22,23 --reps 3 | tail -n 2
echo "== skinny 24 3 5"; python tools/one_step.py 24 3 5 --algo skinny --ka 1,11,17,21,26 --reps 3 | tail -n 2
echo "== skinny 25 2 5"; python tools/one_step.py 25 2 5 --algo skinny --ka 0,1,24,28,29 --reps 3 | tail -n 2
echo "== skinny 24 4 4"; python tools/one_step.py 24 4 4 --algo skinny --reps 3 | tail -n 2
} > gpurun_out/one_q.log 2>&1
cat gpurun_out/one_q.log
