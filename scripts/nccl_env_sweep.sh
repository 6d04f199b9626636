# diagnostic (GPU box, N GPUs): the step's collectives alone under a few NCCL settings
N=${N:-8}
run() { echo "== $*"; env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/nccl_probe.py 2>/dev/null | grep -E "all_reduce|all_gather 786|reduce_scatter" ; }
run X=1
run NCCL_ALGO=NVLS
run NCCL_ALGO=Ring
run NCCL_ALGO=Tree
run NCCL_MIN_CTAS=32
run NCCL_NVLS_ENABLE=0
