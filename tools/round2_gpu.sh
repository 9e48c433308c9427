# round-2 measurement pass on one B200: tests, smoke, the default bench line (tag = $1)
T=${1:-r02a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
tail -5 gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_smoke.log
cut -c1-600 gpurun_out/${T}_bench_c2.json
tail -5 gpurun_out/${T}_bench_c2.err
