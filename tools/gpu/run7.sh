set -x
timeout 900 python -m pytest tests/test_bch.py tests/test_e2e_gpu.py tests/test_chain_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02g_pytest.log
tail -12 gpurun_out/r02g_pytest.log
