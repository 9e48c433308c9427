# multi-GPU pass on one 8 x B200 box: correctness, weak scaling of config 2, strong scaling of the 1M sweep
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_multi.log
timeout 600 $TR --nproc-per-node 8 --master-port 29541 tools/dist_check.py > gpurun_out/dist8.log 2>&1
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_c2_${n}gpu.json 2> gpurun_out/bench_c2_${n}gpu.err
done
for n in 8 4 2; do
  timeout 900 $TR --nproc-per-node $n --master-port 2956$n bench.py --gpus $n --workload c5 --steps 2 --warmup 3 > gpurun_out/bench_c5_${n}gpu.json 2> gpurun_out/bench_c5_${n}gpu.err
done
timeout 600 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --exchange allgather > gpurun_out/bench_c2_8gpu_allgather.json 2> gpurun_out/bench_c2_8gpu_allgather.err
timeout 600 $TR --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --workload c5 --steps 2 --exchange allgather > gpurun_out/bench_c5_8gpu_allgather.json 2> gpurun_out/bench_c5_8gpu_allgather.err
