NP=${NP:-8}
run() { echo "== $*" ; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 tools/sharded_bench.py 4032 2>&1 | grep -E "sharded|shard trace rank [01]:|rror" | tail -3; }
run T2B200_SHARD_TRACE=1
run T2B200_SHARD_TRACE=1 T2B200_NCCL_MAX_CTAS=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 tests/test_sharded_gpu2.py 2>&1 | grep -E "code|SHARDED"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NP --steps 10 --warmup 3 > gpurun_out/r02_bench_n$NP.json 2> gpurun_out/r02_bench_n$NP.err
head -c 200 gpurun_out/r02_bench_n$NP.json
