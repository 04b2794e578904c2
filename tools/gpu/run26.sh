run() { echo "== $*" ; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sharded_bench.py 4032 2>&1 | grep -E "sharded|shard trace" | tail -7; }
run T2B200_SHARD_TRACE=1
run T2B200_NCCL_MAX_CTAS=0
run T2B200_NCCL_MAX_CTAS=2
run T2B200_SHARD_CHUNK=2016
run T2B200_SHARD_CHUNK=576
