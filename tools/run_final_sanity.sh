# Round-end sanity on one B200: the tests touched last, smoke(), and the default bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_dp_gpu.py tests/test_graph_step_gpu.py tests/test_step_gpu.py -x -q > gpurun_out/final_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_tests.log
tail -n 4 gpurun_out/final_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final_smoke.log
tail -n 2 gpurun_out/final_smoke.log
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/bench_final.json
