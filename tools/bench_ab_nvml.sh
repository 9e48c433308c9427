for i in 1 2 3 4 5 6; do
  python bench.py --no-cpu --no-c5 --no-extra --no-e2e 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); d=j['step_ms_distribution']; print('nvml on ', round(j['ms_per_step'],4), round(d['median'],4), round(d['max'],3), round(d['host_enqueue_ms_per_step'],3), j['clocks']['samples'])"
  SEMGATE_BENCH_NO_NVML=1 python bench.py --no-cpu --no-c5 --no-extra --no-e2e 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); d=j['step_ms_distribution']; print('nvml off', round(j['ms_per_step'],4), round(d['median'],4), round(d['max'],3), round(d['host_enqueue_ms_per_step'],3))"
done
