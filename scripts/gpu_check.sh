#!/bin/bash
# scripts/gpu_check.sh -- run on the B200 box via gpurun: GPU tests, smoke, a short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
cat gpurun_out/smoke.log
timeout 600 python bench.py --nu 100 --steps 5 --warmup 3 > gpurun_out/bench_nu100.json 2> gpurun_out/bench_nu100.err; echo "exit $?"
cat gpurun_out/bench_nu100.json; tail -5 gpurun_out/bench_nu100.err
