set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02t_pytest.log
tail -8 gpurun_out/r02t_pytest.log
timeout 300 python tools/ldpc_quick_bench.py 2,1,5 4096 2>&1 | grep group32 > gpurun_out/r02t_quick.log
cat gpurun_out/r02t_quick.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
tail -c 600 gpurun_out/r02t_bench.err
head -c 300 gpurun_out/r02t_bench.json
