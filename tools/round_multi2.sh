TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 tools/dist_check.py > gpurun_out/dist2b.log 2>&1
for x in allgather peer; do
  timeout 600 $TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --exchange $x > gpurun_out/bench_c2_2gpu_$x.json 2> gpurun_out/bench_c2_2gpu_$x.err
  timeout 900 $TR --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --workload c5 --steps 2 --exchange $x > gpurun_out/bench_c5_2gpu_$x.json 2> gpurun_out/bench_c5_2gpu_$x.err
done
