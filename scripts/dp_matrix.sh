# diagnostic (GPU box, N GPUs): data-parallel step time over reducer kind x NCCL CTA cap
N=${N:-2}
for R in arena flat; do for C in 0 4 8 16; do
  if [ "$C" = "0" ]; then export -n NCCL_MAX_CTAS; unset NCCL_MAX_CTAS; CE="NCCL_MAX_CTAS=32"; else CE="NCCL_MAX_CTAS=$C"; fi
  env $CE MMSA_DP_REDUCER=$R timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 --no-extra-configs --sustained-seconds 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$R', '$CE', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d['dp_parity']['ok'])
"
done; done
