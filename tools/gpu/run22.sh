set -x
# LDPC r2/3: full capture (kept) + summary
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 1 -c 1 -o gpurun_out/r02_ldpc_v13_r23 python tools/ldpc_profile_run.py 2 576 3 2.9 > gpurun_out/r02v_ncu.log 2>&1
ncu -i gpurun_out/r02_ldpc_v13_r23.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02_ldpc_v13.summary.csv
# streaming kernels: one full capture of each, summaries only
timeout 900 ncu --set full --clock-control none -k regex:'fft_pass|equalize|ti_deint|demap_' -s 24 -c 8 -o gpurun_out/r02_stream_full python tools/chain_profile_run.py 20 >> gpurun_out/r02v_ncu.log 2>&1
ncu -i gpurun_out/r02_stream_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02_stream_kernels.summary.csv
rm -f gpurun_out/r02_stream_full.ncu-rep
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02v_benchncu.log 2>&1
tail -3 gpurun_out/r02v_ncu.log
wc -l gpurun_out/r02_ldpc_v13.summary.csv gpurun_out/r02_stream_kernels.summary.csv gpurun_out/r02_launches.csv
