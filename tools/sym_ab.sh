# A/B of the symmetric sweep's run-table schedule at config 2 (development tool): K2 ms per variant
run() { env "$@" python bench.py --no-cpu --no-e2e --steps 20 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'step_ms', round(j['ms_per_step'], 4), 'k2_ms', round(j['roofline']['kernel_ms'], 4))"; }
run SEMGATE_SYM_TABLE=1
run SEMGATE_SYM_TABLE=1 SEMGATE_L2_HINT=1
run SEMGATE_SYM_TABLE=1 SEMGATE_L2_HINT=0
run SEMGATE_SYM_TABLE=1 SEMGATE_SYM_RM=10
run SEMGATE_SYM_TABLE=1 SEMGATE_SYM_RM=10 SEMGATE_L2_HINT=1
run SEMGATE_SYM_TABLE=1 SEMGATE_SYM_RUN=4
run SEMGATE_SYM_TABLE=1 SEMGATE_SYM_RUN=16 SEMGATE_SYM_RM=10
run SEMGATE_SYM_TABLE=0
run SEMGATE_SYM_TABLE=0 SEMGATE_WINDOW_MB=0
