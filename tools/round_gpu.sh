# round measurement pass on one B200: tests, smoke, bench lines, ncu launch list + full captures
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r01d_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r01d_smoke.log 2>&1
python bench.py > gpurun_out/r01d_bench_c2.json 2> gpurun_out/r01d_bench_c2.err
python bench.py --symmetric off --no-cpu > gpurun_out/r01d_bench_c2_full.json 2> gpurun_out/r01d_bench_c2_full.err
python bench.py --workload c1 > gpurun_out/r01d_bench_c1.json 2> gpurun_out/r01d_bench_c1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_bench_ref.json 2> gpurun_out/r01d_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01d_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gated_topk -c 1 -f -o gpurun_out/r01d_k2_sym python tools/ncu_target.py 2 20000 4096 2 > gpurun_out/r01d_ncu_f.log 2>&1
