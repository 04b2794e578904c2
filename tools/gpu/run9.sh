set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 1 -c 1 -o gpurun_out/r02_ldpc_v9_r34 python tools/ldpc_profile_run.py 3 576 3 > gpurun_out/r02i_ncu.log 2>&1
tail -3 gpurun_out/r02i_ncu.log
