# diagnostic (GPU box, N GPUs): flat one-bucket vs arena (early buckets on / off) over the NCCL CTA cap
N=${N:-8}
run() {
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 --no-extra-configs --sustained-seconds 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d['dp_parity']['ok'], d['dp_parity']['grad_rel_max'])
"
}
run MMSA_DP_REDUCER=flat NCCL_MAX_CTAS=32
run MMSA_DP_REDUCER=arena MMSA_DP_BUCKETS=0 NCCL_MAX_CTAS=32
run MMSA_DP_REDUCER=arena MMSA_DP_BUCKETS=1 NCCL_MAX_CTAS=32
run MMSA_DP_REDUCER=arena MMSA_DP_BUCKETS=1 NCCL_MAX_CTAS=16
