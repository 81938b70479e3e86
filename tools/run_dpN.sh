# bench.py on N GPUs of one box under torchrun, bounded by a timeout (usage: gpurun --gpus N -- bash tools/run_dpN.sh N); the JSON line lands in gpurun_out/
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SECONDS=0
timeout 170 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_graph.log 2>&1; echo "rc=$? after ${SECONDS}s" >> gpurun_out/bench_n${N}_graph.log
tail -c 200 gpurun_out/bench_n${N}_graph.log
grep "^{" gpurun_out/bench_n${N}_graph.log > gpurun_out/bench_r02_cfg2_n${N}_graph.json
