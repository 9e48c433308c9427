# throughput under sustained load for schedule / L2 policy variants
python tools/l2_sweep.py 100000,300000 "1,38,0,0;1,38,0,1;1,38,4,1;1,38,8,1;1,38,4,2;1,24,4,1;1,64,4,1;2,38,0,0;2,38,4,1;2,38,8,2;2,64,4,1" 2.5 2>&1 | grep "^n="
python tools/l2_sweep.py 1000000 "1,38,0,0;1,38,4,1;2,38,4,1" 1 2>&1 | grep "^n="
# DRAM traffic of one launch (100k all-pairs) for the key variants
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
for v in "1 0 0" "1 4 1" "2 0 0" "2 4 1"; do
  set -- $v
  echo "== ncu cg=$1 window=$2 hint=$3"
  SEMGATE_RM_CAP_MB=38 SEMGATE_SYNC_WINDOW=$2 SEMGATE_L2_HINT=$3 ncu --metrics $M --clock-control none -k regex:gated_topk -c 1 python tools/ncu_target.py $1 100000 4096 1 2>&1 | grep -E "dram__|lts__|gpu__time|sm__pipe|candidates"
done
