set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread --clock-control none -k regex:'fft|equalize|ti_|demap' -c 60 --csv --log-file gpurun_out/r02a_stream_kernels.csv python tools/chain_profile_run.py 20 > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_pytest.log; tail -c 600 gpurun_out/r02a_bench.json
