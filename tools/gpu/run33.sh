set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_final_pytest.log
tail -5 gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -3 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
head -c 300 gpurun_out/r02_bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_reference.json 2> gpurun_out/r02_bench_final_reference.err
head -c 400 gpurun_out/r02_bench_final_reference.json
