set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
USTRUN_DP_LANES=2 USTRUN_DP_USE_GRAPH=1 USTRUN_PRECISION=bf16 timeout 240 $TR --master-port 29541 tools/dp_check.py > gpurun_out/dpc_a.log 2>&1; echo "rc=$?" >> gpurun_out/dpc_a.log
USTRUN_DP_LANES=4 USTRUN_DP_USE_GRAPH=1 USTRUN_DP_MODEL=b_dsbn USTRUN_PRECISION=bf16 timeout 240 $TR --master-port 29542 tools/dp_check.py > gpurun_out/dpc_b.log 2>&1; echo "rc=$?" >> gpurun_out/dpc_b.log
timeout 400 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_graph.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_graph.log
timeout 400 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --lanes 1 --no-dp-parity > gpurun_out/bench_n2_graph_l1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_graph_l1.log
grep -h "dp_check\|rc=" gpurun_out/dpc_a.log gpurun_out/dpc_b.log | cut -c1-300
tail -c 300 gpurun_out/bench_n2_graph.log; tail -c 300 gpurun_out/bench_n2_graph_l1.log
