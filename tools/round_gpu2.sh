nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_probe tools/probe/cluster_probe.cu && /tmp/cluster_probe > gpurun_out/cluster_probe.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu12.log
python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c5 --steps 2 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
ncu --set full --clock-control none --import-source on -k regex:gated_topk -c 1 -f -o gpurun_out/r01b_k2_cg2_300k python tools/ncu_target.py 2 300000 4096 1 > gpurun_out/ncu_f4.log 2>&1
