# round measurement pass on one B200: tests, bench lines, companion kernels, ncu launch list + full capture
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c5 --steps 2 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python tools/bench_kernels.py > gpurun_out/kernels5.json 2> gpurun_out/kernels5.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gated_topk -c 1 -f -o gpurun_out/r01c_k2_cg2 python tools/ncu_target.py 2 20000 4096 2 > gpurun_out/ncu_f5.log 2>&1
python bench.py --workload c1 --cta-group 1 --no-cpu --no-e2e > gpurun_out/bench_c1_cg1.json 2> gpurun_out/bench_c1_cg1.err
