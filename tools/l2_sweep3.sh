python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu9.log
python tools/l2_sweep.py 20000 "2,38,0,0;2,38,0,1;2,38,0,2;2,38,3,2;2,38,3,1;1,38,0,0;1,38,3,2;1,38,0,2;2,38,0,0" 2.0 2>&1 | grep "^n="
python tools/l2_sweep.py 300000 "1,38,3,2;1,80,3,2;1,60,3,2" 1.8 2>&1 | grep "^n="
python tools/l2_sweep.py 1000000 "1,38,3,2;1,38,2,2" 1 2>&1 | grep "^n="
python tools/bench_kernels.py > gpurun_out/kernels2.json 2> gpurun_out/kernels2.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second
for v in "1 3 2" "2 3 2"; do
  set -- $v
  echo "== ncu 300k cg=$1 window=$2 hint=$3"
  SEMGATE_RM_CAP_MB=38 SEMGATE_SYNC_WINDOW=$2 SEMGATE_L2_HINT=$3 ncu --metrics $M --clock-control none -k regex:gated_topk -c 1 python tools/ncu_target.py $1 300000 4096 1 2>&1 | grep -E "dram__|lts__|gpu__time|sm__|candidates"
done
