NP=${NP:-8}
run() { echo "== $*" ; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 tools/sharded_bench.py 4032 2>&1 | grep -E "sharded|shard trace rank [01]:" | tail -3; }
run T2B200_SHARD_TRACE=1 T2B200_NCCL_MAX_CTAS=2
run T2B200_SHARD_TRACE=1 T2B200_NCCL_MAX_CTAS=8 T2B200_SHARD_SLOTS=8
run T2B200_SHARD_TRACE=1 T2B200_NCCL_MAX_CTAS=16 T2B200_SHARD_SLOTS=8
run T2B200_SHARD_TRACE=1 T2B200_NCCL_MAX_CTAS=4 T2B200_SHARD_SLOTS=8
run T2B200_SHARD_TRACE=1 T2B200_NCCL_MAX_CTAS=16 T2B200_SHARD_SLOTS=7
