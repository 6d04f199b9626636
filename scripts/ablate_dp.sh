# diagnostic: attribute the data-parallel overhead of bench.py at N GPUs (default 2)
N=${N:-2}
run() { # label, env..., then extra args after --
  label=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
"
}
run "full" X=1
run "full mixing=0" NCCL_GRAPH_MIXING_SUPPORT=0
STEPS=100 run "full steps=100" X=1
run "no-shard" MMSA_BENCH_ABLATE=shard
run "no-reduce" MMSA_BENCH_ABLATE=reduce
run "neither" MMSA_BENCH_ABLATE=shard,reduce
