#!/bin/bash
# One gpurun call (1 GPU): plain short bench, then the ncu launch list of the same command, the DRAM traffic of every
# tcgen05 GEMM launch of one step, and `--set full` captures of the dominant kernels.  Outputs land in gpurun_out/;
# scripts/summarize_ncu.py distils them into profiles/ on the CPU box.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 1 --no-graph --sustained-seconds 0 --no-extra-configs --no-torch-eager"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
# ~110 library launches (+ ~40 torch plumbing kernels) per step; skip TrainStep.warmup (2 iterations x 2 slots) and take about 3 steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
# DRAM bytes of the 34 tcgen05 GEMM launches of one step (after the 4 warm-up steps)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_tcgen05 -s 136 -c 34 --csv --log-file gpurun_out/gemm_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "gemm traffic exit=$?"
# (gpurun brings back at most 64 MiB: ~2 MB per fully captured launch -> 10 forward + 8 backward GEMMs + 6 row/attention kernels)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 136 -c 10 \
    -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm fwd capture exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 150 -c 8 \
    -f -o gpurun_out/prof_gemm2 $CMD > gpurun_out/ncu_gemm2.log 2>&1
echo "gemm bwd capture exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gate_ln_pool|attn_" -s 16 -c 6 \
    -f -o gpurun_out/prof_mem $CMD > gpurun_out/ncu_mem.log 2>&1
echo "mem capture exit=$?"
du -sh gpurun_out
ls -la gpurun_out | head -30
