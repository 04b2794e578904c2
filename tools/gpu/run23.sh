set -x
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size)
import ctypes
cudart=ctypes.CDLL('libcudart.so')
PY
timeout 300 python -m pytest tests/test_ldpc_gpu.py -m gpu -x -q 2>&1 | tail -3
for np_ in 0 1; do
echo "NO_PERSIST=$np_" >> gpurun_out/r02w_quick.log
if [ $np_ = 1 ]; then export T2B200_LDPC_NO_PERSIST=1; else unset T2B200_LDPC_NO_PERSIST; fi
timeout 300 python tools/ldpc_quick_bench.py 2,1,5 4096 2>&1 | grep group32 >> gpurun_out/r02w_quick.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ldpc_decode -s 1 -c 1 python tools/ldpc_profile_run.py 2 576 3 2.9 2>&1 | grep -E "dram__|gpu__time|lts__" >> gpurun_out/r02w_quick.log
done
cat gpurun_out/r02w_quick.log
