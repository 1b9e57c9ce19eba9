#!/bin/bash
# scripts/gpu_partition_check.sh -- partitioned-mode tests on one GPU (in-process transport).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "partitioned" > gpurun_out/pytest_partition.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_partition.log
grep -E "partitioned|cg iterations|passed|failed|Error|error" gpurun_out/pytest_partition.log | tail -30
tail -5 gpurun_out/pytest_partition.log
