timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu14.log
for e in 0 1; do
  SEMGATE_EPI_MODE=$e python bench.py --no-cpu --no-e2e > gpurun_out/ab_c2_epi$e.json 2>> gpurun_out/ab.err
  SEMGATE_EPI_MODE=$e python bench.py --no-cpu --no-e2e --workload c1 > gpurun_out/ab_c1_epi$e.json 2>> gpurun_out/ab.err
done
SEMGATE_RM_CAP_MB=80 python bench.py --no-cpu --no-e2e > gpurun_out/ab_c2_cap80.json 2>> gpurun_out/ab.err
SEMGATE_RM_CAP_MB=160 python bench.py --no-cpu --no-e2e > gpurun_out/ab_c2_cap160.json 2>> gpurun_out/ab.err
SEMGATE_WINDOW_MB=0 python bench.py --no-cpu --no-e2e > gpurun_out/ab_c2_win0.json 2>> gpurun_out/ab.err
SEMGATE_L2_HINT=0 python bench.py --no-cpu --no-e2e > gpurun_out/ab_c2_hint0.json 2>> gpurun_out/ab.err
python bench.py --no-cpu --no-e2e --cta-group 1 > gpurun_out/ab_c2_cg1.json 2>> gpurun_out/ab.err
for f in gpurun_out/ab_*.json; do python - "$f" <<'PY'
import json,sys
j=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1]); r=j['roofline']
print(sys.argv[1], round(j['ms_per_step'],4), 'K2 ms', round(r['kernel_ms'],4), 'TF/s', round(r['achieved']), 'frac', round(r['frac'],3))
PY
done
