python tools/l2_sweep.py 20000,100000 "2,40,8,2;4,40,8,2;4,40,0,2;4,80,8,2" 1.5 2>&1 | grep "^n=" > gpurun_out/quad_sweep2.log
python bench.py --cta-group 4 --no-cpu > gpurun_out/bench_c2_quad.json 2> gpurun_out/bench_c2_quad.err
