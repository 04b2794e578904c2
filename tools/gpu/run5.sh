set -x
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_e2e_gpu.py tests/test_fft_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02e_pytest_new.log
tail -12 gpurun_out/r02e_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02e_pytest.log
tail -6 gpurun_out/r02e_pytest.log
