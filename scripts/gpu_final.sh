#!/bin/bash
# One gpurun call (1 GPU): the bench lines tracked under profiles/ (default N=1, with optimizer, reference arm, the two
# larger BASELINE shapes) and one `--set full` capture of backward GEMM launches (dgrad / wgrad mix).
mkdir -p gpurun_out
rm -f gpurun_out/prof_gemm2.ncu-rep
timeout 300 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "n1 exit=$?"
timeout 300 python bench.py --optimizer > gpurun_out/final_n1_opt.json 2> gpurun_out/final_n1_opt.err; echo "opt exit=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit=$?"
timeout 300 python bench.py --batch 1024 --L 128 --steps 20 > gpurun_out/final_b1024_l128.json 2> gpurun_out/final_b1024_l128.err; echo "b1024 l128 exit=$?"
timeout 300 python bench.py --batch 1024 --L 512 --steps 10 > gpurun_out/final_b1024_l512.json 2> gpurun_out/final_b1024_l512.err; echo "b1024 l512 exit=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-graph"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 174 -c 8 \
    -f -o gpurun_out/prof_gemm2 $CMD > gpurun_out/ncu_gemm2.log 2>&1
echo "gemm2 capture exit=$?"
for f in n1 n1_opt ref b1024_l128 b1024_l512; do python scripts/show_bench.py gpurun_out/final_$f.json 2>/dev/null | head -1; done
