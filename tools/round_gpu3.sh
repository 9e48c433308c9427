timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu13.log
python tools/l2_sweep.py 20000,100000,300000 "2,40,8,2;4,40,8,2;4,40,4,2;4,80,8,2" 1.8 2>&1 | grep "^n=" > gpurun_out/quad_sweep.log
python bench.py --cta-group 4 --no-cpu > gpurun_out/bench_c2_quad.json 2> gpurun_out/bench_c2_quad.err
python bench.py --cta-group 2 --no-cpu > gpurun_out/bench_c2_pair.json 2> gpurun_out/bench_c2_pair.err
