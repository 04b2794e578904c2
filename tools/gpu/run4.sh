set -x
timeout 600 python -m pytest tests/test_ldpc_gpu.py tests/test_chain_gpu.py tests/test_ts_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02d_pytest.log
tail -5 gpurun_out/r02d_pytest.log
timeout 300 python tools/ldpc_quick_bench.py 2,0,1,3,4,5 4096 > gpurun_out/r02d_quick.log 2>&1
cat gpurun_out/r02d_quick.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -c 1500 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
