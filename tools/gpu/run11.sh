set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 1 -c 1 -o gpurun_out/r02_ldpc_v10_r23 python tools/ldpc_profile_run.py 2 576 3 2.9 > gpurun_out/r02k_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 1 -c 1 -o gpurun_out/r02_ldpc_v10_r56 python tools/ldpc_profile_run.py 5 576 3 4.3 >> gpurun_out/r02k_ncu.log 2>&1
tail -6 gpurun_out/r02k_ncu.log
