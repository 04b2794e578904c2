set -x
timeout 600 python -m pytest tests/test_ldpc_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02o_pytest.log
tail -5 gpurun_out/r02o_pytest.log
for t in 2 5 8 12 400; do
echo "WALK_MIN $t" >> gpurun_out/r02o_quick.log
T2B200_LDPC_WALK_MIN=$t timeout 300 python tools/ldpc_quick_bench.py 2,1,3,4,5,8 4096 2>&1 | grep group32 >> gpurun_out/r02o_quick.log
done
cat gpurun_out/r02o_quick.log
