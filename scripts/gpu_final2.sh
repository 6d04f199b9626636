#!/bin/bash
# One gpurun call (1 GPU) at the end of a round: the -m gpu suite, smoke(), the tracked bench lines (default N=1, with the
# optimizer, the reference arm) and the ncu launch list of one eager step.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "n1 exit=$?"
timeout 200 python bench.py --optimizer --no-extra-configs --no-torch-eager > gpurun_out/final_n1_opt.json 2> gpurun_out/final_n1_opt.err; echo "opt exit=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-extra-configs --no-torch-eager > gpurun_out/final_ncu.log 2>&1; echo "launch list exit=$?"
for f in n1 n1_opt ref; do python scripts/show_bench.py gpurun_out/final_$f.json 2>/dev/null | head -1; done
