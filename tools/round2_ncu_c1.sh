# ncu pass for K2 at config 1 (5k x 512-d: the epilogue-bound regime)
T=${1:-r02w}
mkdir -p gpurun_out
python tools/ncu_target.py 0 5000 512 3 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gated_topk -s 1 -c 1 -f -o gpurun_out/${T}_k2_c1 python tools/ncu_target.py 0 5000 512 3 > gpurun_out/${T}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${T}_ncu.log
