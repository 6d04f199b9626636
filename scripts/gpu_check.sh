#!/bin/bash
# One gpurun call: kernel tests (GEMM first, isolated, so a hung tcgen05 pipeline cannot take the rest down),
# model tests, then an optional bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, args...
  local name=$1; local to=$2; shift 2
  timeout $to python -m pytest "$@" -q --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run gemm 300 tests/test_gpu_kernels.py -m gpu -k "linear"
run kernels 600 tests/test_gpu_kernels.py -m gpu -k "not linear"
run model 900 tests/test_gpu_model.py -m gpu
if [ -f bench.py ] && [ "$1" != "nobench" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 2 gpurun_out/bench.log
fi
cat gpurun_out/summary.txt
