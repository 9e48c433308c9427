timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu15.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python tools/l2_sweep.py 20000,100000,300000 "2,40,8,2" 2.0 2>&1 | grep "^n=" > gpurun_out/sustained_check.log
python tools/bench_kernels.py > gpurun_out/kernels4.json 2> gpurun_out/kernels4.err
