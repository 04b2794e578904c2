set -x
timeout 900 python -m pytest tests/test_fec_front_gpu.py tests/test_chain_gpu.py tests/test_e2e_gpu.py tests/test_edge_cases_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02u_pytest.log
tail -4 gpurun_out/r02u_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread --clock-control none -k regex:'equalize|ti_|demap' -c 12 --csv --log-file gpurun_out/r02u_stream_kernels.csv python tools/chain_profile_run.py 20 > gpurun_out/r02u_ncu.log 2>&1
tail -2 gpurun_out/r02u_ncu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
tail -c 300 gpurun_out/r02u_bench.err
head -c 300 gpurun_out/r02u_bench.json
