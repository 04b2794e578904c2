NP=2
run() { echo "== $*" ; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 tools/sharded_bench.py 4032 2>&1 | grep -E "sharded|shard trace rank [01]:|rror" | tail -3; }
run T2B200_SHARD_TRACE=1
run T2B200_SHARD_TRACE=1 T2B200_SHARD_GROW=1
run T2B200_SHARD_TRACE=1 T2B200_SHARD_GROW=4
run T2B200_SHARD_TRACE=1 T2B200_SHARD_WIDE=1
run T2B200_SHARD_TRACE=1 T2B200_SHARD_WIDE=1 T2B200_SHARD_GROW=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 tests/test_sharded_gpu2.py 2>&1 | grep -E "code|SHARDED"
