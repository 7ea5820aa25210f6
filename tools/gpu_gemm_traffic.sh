# DRAM traffic of the fat GEMM (ncu) for three sweep-group sizes -> profiles/r01_gemm_sweep_group_traffic.txt
export TNC_EXPERIMENTS=1   # the TNC_* variant knobs below are only honoured with this (include/tnc_b200.h)
mkdir -p gpurun_out
for gm in 16 8 4; do
  TNC_TC_GROUP_M=$gm timeout -s KILL 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_2cta -s 1 -c 1 --csv --log-file gpurun_out/traffic_gm$gm.csv \
      python tools/one_step.py 15 13 15 --algo tc --reps 2 > gpurun_out/traffic_gm$gm.log 2>&1
  echo "group_m=$gm"; grep -v "^==" gpurun_out/traffic_gm$gm.csv | cut -d, -f5,13-15 | tail -n 5
done
