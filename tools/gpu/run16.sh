set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02p_pytest.log
tail -5 gpurun_out/r02p_pytest.log
timeout 300 python tools/ldpc_quick_bench.py 2,0,1,3,4,5,6,7,8,9,10,11 4096 > gpurun_out/r02p_quick.log 2>&1
cat gpurun_out/r02p_quick.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
tail -c 600 gpurun_out/r02p_bench.err
head -c 400 gpurun_out/r02p_bench.json
