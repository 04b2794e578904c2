set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/test_sharded_gpu2.py > gpurun_out/r02f_sharded.log 2>&1
tail -12 gpurun_out/r02f_sharded.log
T2B200_BENCH_SKIP_EXTRAS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
tail -c 1800 gpurun_out/r02f_bench_n2.json; tail -5 gpurun_out/r02f_bench_n2.err
