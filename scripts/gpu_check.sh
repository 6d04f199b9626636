#!/bin/bash
# One gpurun call: kernel tests (GEMM first, isolated, so a hung tcgen05 pipeline cannot take the rest down),
# model tests, diagnostics, then the bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, args...
  local name=$1; local to=$2; shift 2
  timeout $to python -m pytest "$@" -q --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run gemm 120 tests/test_gpu_kernels.py -m gpu -k "linear"
run kernels 150 tests/test_gpu_kernels.py -m gpu -k "not linear"
run model 240 tests/test_gpu_model.py -m gpu
if [ "$1" != "nodiag" ]; then
  timeout 600 python scripts/diag_parity.py > gpurun_out/diag.log 2>&1; echo "diag exit=$?" | tee -a gpurun_out/summary.txt
fi
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
cat gpurun_out/summary.txt
