set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02r_pytest.log
tail -5 gpurun_out/r02r_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
tail -c 600 gpurun_out/r02r_bench.err
head -c 300 gpurun_out/r02r_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread --clock-control none -k regex:'fft|equalize|ti_|demap' -c 40 --csv --log-file gpurun_out/r02r_stream_kernels.csv python tools/chain_profile_run.py 20 > gpurun_out/r02r_ncu.log 2>&1
tail -2 gpurun_out/r02r_ncu.log
