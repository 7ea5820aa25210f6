# compute-sanitizer over the kernels and host paths added in round 2 (small configurations)
mkdir -p gpurun_out
( timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "amax_reduced or tc_3m_complex or (slice_reuse_is_bit and n12) or cuda_graph_replay and n12 or rowdot or queued_contractions" ) > gpurun_out/san2_mem.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/san2_mem.log
tail -n 5 gpurun_out/san2_mem.log
( timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "(amax_reduced and 3xf16 and 1.0) or (stem_bulk and not segments)" ) > gpurun_out/san2_race.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/san2_race.log
tail -n 5 gpurun_out/san2_race.log
( timeout -s KILL 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "(amax_reduced and 3xf16 and 1.0)" ) > gpurun_out/san2_sync.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/san2_sync.log
tail -n 4 gpurun_out/san2_sync.log
