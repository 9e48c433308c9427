# round-2 multi-GPU pass on one box: bash tools/round2_multi.sh <ranks> <tag> [what...]
#   what: check (tools/dist_check.py), c2 (default bench line incl. the 1M-keyframe leg), c3, c4
N=${1:-2}; T=${2:-r02m}; shift; shift
WHAT=${@:-check c2 c3}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
for w in $WHAT; do
  case $w in
    check) timeout 900 $TR --master-port 29541 tools/dist_check.py > gpurun_out/${T}_dist${N}.log 2>&1; echo "dist_check rc=$?"; grep -c "ok=True\|: True" gpurun_out/${T}_dist${N}.log; grep "False\|Error\|error" gpurun_out/${T}_dist${N}.log | head -5;;
    c2) timeout 900 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_c2_${N}gpu.json 2> gpurun_out/${T}_bench_c2_${N}gpu.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c2_${N}gpu.json; tail -3 gpurun_out/${T}_bench_c2_${N}gpu.err;;
    c3|c4|c5) timeout 900 $TR --master-port 29561 bench.py --gpus $N --workload $w --steps 5 --warmup 3 > gpurun_out/${T}_bench_${w}_${N}gpu.json 2> gpurun_out/${T}_bench_${w}_${N}gpu.err; echo "$w rc=$?"; cut -c1-300 gpurun_out/${T}_bench_${w}_${N}gpu.json; tail -3 gpurun_out/${T}_bench_${w}_${N}gpu.err;;
  esac
done
