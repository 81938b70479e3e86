set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_gpu.py -x -q > gpurun_out/dp_test.log 2>&1; echo "rc=$?" >> gpurun_out/dp_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
USTRUN_DP_LANES=2 USTRUN_PRECISION=bf16 timeout 240 $TR --master-port 29541 tools/dp_check.py > gpurun_out/dpc_a.log 2>&1; echo "rc=$?" >> gpurun_out/dpc_a.log
USTRUN_DP_LANES=4 USTRUN_DP_MODEL=b_dsbn USTRUN_PRECISION=bf16 timeout 240 $TR --master-port 29542 tools/dp_check.py > gpurun_out/dpc_b.log 2>&1; echo "rc=$?" >> gpurun_out/dpc_b.log
timeout 400 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_lanes.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_lanes.log
timeout 400 $TR --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --lanes 1 --no-dp-parity > gpurun_out/bench_n2_l1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_l1.log
tail -3 gpurun_out/dp_test.log gpurun_out/dpc_a.log gpurun_out/dpc_b.log; tail -c 600 gpurun_out/bench_n2_lanes.log; tail -c 300 gpurun_out/bench_n2_l1.log
