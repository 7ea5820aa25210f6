# compute-sanitizer (memcheck, racecheck) over the small streaming-kernel tests.
mkdir -p gpurun_out
# memcheck + racecheck on the small configurations (new bulk-copy kernel: mbarriers, async copies)
( timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "stem_bulk or (stem_single and not 14) or smoke" ) > gpurun_out/san_mem.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/san_mem.log
tail -n 6 gpurun_out/san_mem.log
( timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "stem_bulk" ) > gpurun_out/san_race.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/san_race.log
tail -n 6 gpurun_out/san_race.log
