set -x
timeout 600 python -m pytest tests/test_sharded_gpu2.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02x_sharded.log
cat gpurun_out/r02x_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02x_bench_n2.json 2> gpurun_out/r02x_bench_n2.err
tail -c 400 gpurun_out/r02x_bench_n2.err
head -c 300 gpurun_out/r02x_bench_n2.json
