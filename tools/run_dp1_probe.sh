# The complete data-parallel path on ONE GPU (1-rank NCCL group): tests/test_dp_gpu.py, then bench.py with USTRUN_BENCH_FORCE_DP=1 (graph + lanes, captured NCCL calls)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_gpu.py -x -q > gpurun_out/dp_test.log 2>&1; echo "rc=$?" >> gpurun_out/dp_test.log
tail -n 25 gpurun_out/dp_test.log | cut -c1-300
export USTRUN_BENCH_FORCE_DP=1
A="--steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline"
timeout 300 python bench.py $A > gpurun_out/dp1_graph.log 2>&1; echo "rc=$?" >> gpurun_out/dp1_graph.log
grep -v "^frame" gpurun_out/dp1_graph.log | grep -i "warn\|error\|rc=" | cut -c1-300 | head
