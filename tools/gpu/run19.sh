set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02s_pytest.log
tail -5 gpurun_out/r02s_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
tail -c 600 gpurun_out/r02s_bench.err
head -c 300 gpurun_out/r02s_bench.json
