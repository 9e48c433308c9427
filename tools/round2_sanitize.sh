# one compute-sanitizer tool per gpurun call: bash tools/round2_sanitize.sh memcheck|racecheck|synccheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/r02_sanitize_plain_${TOOL}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_sanitize_plain_${TOOL}.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --log-file gpurun_out/r02_sanitizer_${TOOL}.log python tools/sanitize_target.py > gpurun_out/r02_sanitizer_${TOOL}.out 2>&1
echo "compute-sanitizer $TOOL rc=$?"
tail -5 gpurun_out/r02_sanitizer_${TOOL}.out
tail -15 gpurun_out/r02_sanitizer_${TOOL}.log
