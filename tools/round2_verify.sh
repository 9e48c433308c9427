# round-2 verification pass on one B200 (tag = $1): tests, smoke, the default bench line three times (stall check)
T=${1:-r02n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err ) 2> gpurun_out/${T}_bench_time.log
for i in 2 3; do timeout 600 python bench.py --no-cpu --no-c5 --no-extra > gpurun_out/${T}_bench_c2_rep$i.json 2> gpurun_out/${T}_bench_c2_rep$i.err; done
tail -3 gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_smoke.log
cut -c1-400 gpurun_out/${T}_bench_c2.json gpurun_out/${T}_bench_c2_rep2.json gpurun_out/${T}_bench_c2_rep3.json
cat gpurun_out/${T}_bench_time.log
tail -5 gpurun_out/${T}_bench_c2.err
