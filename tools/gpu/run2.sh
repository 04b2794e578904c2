set -x
timeout 600 python -m pytest tests/test_ldpc_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02b_pytest_ldpc.log
tail -5 gpurun_out/r02b_pytest_ldpc.log
timeout 300 python tools/ldpc_quick_bench.py 2,0,3,5,7 4096 > gpurun_out/r02b_quick.log 2>&1
cat gpurun_out/r02b_quick.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
