NP=2
T2B200_SHARD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 10 --warmup 3 > gpurun_out/r02_bench_n$NP.json 2> gpurun_out/r02_bench_n$NP.err
grep "shard trace" gpurun_out/r02_bench_n$NP.err | tail -8
head -c 150 gpurun_out/r02_bench_n$NP.json
