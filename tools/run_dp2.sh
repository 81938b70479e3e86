mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
SECONDS=0
timeout 200 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_graph.log 2>&1; echo "rc=$? after ${SECONDS}s" >> gpurun_out/bench_n2_graph.log
tail -c 200 gpurun_out/bench_n2_graph.log
grep "^{" gpurun_out/bench_n2_graph.log > gpurun_out/bench_r02_cfg2_n2_graph.json
