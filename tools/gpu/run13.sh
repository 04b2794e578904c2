set -x
timeout 600 python -m pytest tests/test_ldpc_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02m_pytest.log
tail -5 gpurun_out/r02m_pytest.log
timeout 300 python tools/ldpc_quick_bench.py 2,0,1,3,4,5,6,8,11 4096 > gpurun_out/r02m_quick.log 2>&1
cat gpurun_out/r02m_quick.log
