set -x
timeout 600 python -m pytest tests/test_ldpc_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02n_pytest.log
tail -5 gpurun_out/r02n_pytest.log
timeout 300 python tools/ldpc_quick_bench.py 2,0,1,3,4,5,6,8 4096 > gpurun_out/r02n_quick.log 2>&1
cat gpurun_out/r02n_quick.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 1 -c 1 -o gpurun_out/r02_ldpc_v12_r23 python tools/ldpc_profile_run.py 2 576 3 2.9 > gpurun_out/r02n_ncu.log 2>&1
