# K5 (re-rank) pass on one B200: forms against each other and the fp32 reference, DINOv2-shape timing, ncu capture of the pair form
T=${1:-r02v}
mkdir -p gpurun_out
timeout 500 python tools/rerank_check.py > gpurun_out/${T}_rerank_check.log 2>&1; echo "check rc=$?"; tail -2 gpurun_out/${T}_rerank_check.log | cut -c1-600
for cfg in "1 -" "0 -"; do set -- $cfg; export SEMGATE_RERANK_RING=$1; echo "ring=$1"; timeout 200 python tools/rerank_check.py 100000 timing-only 2>&1 | tail -1 | cut -c1-300; done
unset SEMGATE_RERANK_RING
python tools/ncu_rerank.py 20000 > gpurun_out/${T}_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:rerank_pair -s 1 -c 1 -f -o gpurun_out/${T}_k5_pair python tools/ncu_rerank.py 20000 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "rerank" > gpurun_out/${T}_pytest_rerank.log 2>&1; tail -3 gpurun_out/${T}_pytest_rerank.log
