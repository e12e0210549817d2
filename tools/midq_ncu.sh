set -x
for cfg in "256 1" "256 2" "512 1" "512 2"; do
  set -- $cfg
  python tools/one_search.py 100 $1 1 $2 > gpurun_out/plain_mid_$1_$2.log 2>&1 || exit 1
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:mips_filter -c 12 --csv --log-file gpurun_out/midq_$1_$2.csv python tools/one_search.py 100 $1 1 $2 > gpurun_out/ncu_mid_$1_$2.log 2>&1
done
