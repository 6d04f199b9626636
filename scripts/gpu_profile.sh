#!/bin/bash
# One gpurun call (1 GPU): plain short bench, then the ncu launch list of the same command and
# `--set full` captures of the dominant kernels.  Outputs land in gpurun_out/; summaries are
# distilled into profiles/ by scripts/summarize_ncu.py on the CPU box.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 1 --no-graph"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
# 104 launches per step; skip TrainStep.warmup (2 steps) and take 3 steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 320 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 92 -c 12 \
    -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gate_ln_pool|attn_|gemm_simt" -s 22 -c 16 \
    -f -o gpurun_out/prof_mem $CMD > gpurun_out/ncu_mem.log 2>&1
echo "mem capture exit=$?"
ls -la gpurun_out
