#!/bin/bash
# scripts/gpu_round_check.sh -- what the driver runs at round end, in one go: GPU tests, smoke, both bench arms,
# plus the batch workload.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"
cat gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload batch_spheres --steps 10 --warmup 3 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; echo "batch exit $?"
cat gpurun_out/bench_batch.json; tail -3 gpurun_out/bench_batch.err
