set -x
timeout 300 python -m pytest tests/test_ldpc_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/ldpc_quick_bench.py 2,1,5 4096 2>&1 | grep group32 > gpurun_out/r02y_quick.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ldpc_decode -s 1 -c 1 python tools/ldpc_profile_run.py 2 576 3 2.9 2>&1 | grep -E "dram__|gpu__time|lts__" >> gpurun_out/r02y_quick.log
cat gpurun_out/r02y_quick.log
