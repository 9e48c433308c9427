# final-code ncu pass (one GPU): launch list of the bench command, full capture + tensor-pipe metrics of K2 (symmetric, config 2).
# Each command first runs plain.  bash tools/round2_ncu_final.sh <tag>
T=${1:-r02c}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-c5 --no-extra > gpurun_out/${T}_ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c5 --no-extra > gpurun_out/${T}_ncu_l.log 2>&1
echo "launch list rc=$?"
python tools/ncu_target.py 2 20000 4096 2 > gpurun_out/${T}_ncu_plain_target.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gated_topk -s 2 -c 1 -f -o gpurun_out/${T}_k2_sym python tools/ncu_target.py 2 20000 4096 2 > gpurun_out/${T}_ncu_f.log 2>&1
echo "k2 full capture rc=$?"
ncu --metrics sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_tensor_op_hmma.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,sm__cycles_active.avg,sm__cycles_elapsed.avg,gpu__time_duration.sum,smsp__cycles_active.avg \
  --clock-control none -k regex:gated_topk -s 2 -c 1 --csv --log-file gpurun_out/${T}_k2_sym_tensor_metrics.csv python tools/ncu_target.py 2 20000 4096 2 > gpurun_out/${T}_ncu_t.log 2>&1
echo "k2 tensor metrics rc=$?"
ls -la gpurun_out/${T}_*ncu-rep
