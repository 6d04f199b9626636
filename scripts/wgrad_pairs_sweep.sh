# diagnostic (GPU box): step time of bench.py over the number of CTA pairs the weight gradients may occupy
for P in ${PAIRS:-74 63 54 45 36}; do
  MMSA_WGRAD_PAIRS=$P timeout 200 python bench.py --steps 50 --warmup 5 --no-extra-configs --no-torch-eager --sustained-seconds 0 --cpu-sample-batch 4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('pairs', $P, 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']))
"
done
