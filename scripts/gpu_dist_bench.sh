#!/bin/bash
# scripts/gpu_dist_bench.sh -- partitioned mode over NCCL (run with gpurun --gpus N): the NCCL test, then strong scaling of
# one grid mesh over 1..N GPUs with the global hierarchy, and block-Jacobi for comparison.
#   NGPU=2 SIZES="2000 4000" bash scripts/gpu_dist_bench.sh
set -x
mkdir -p gpurun_out
NGPU=${NGPU:-2}
SIZES=${SIZES:-"2000 4000"}
nvidia-smi -L > gpurun_out/dist_gpus.txt; nproc >> gpurun_out/dist_gpus.txt; free -g >> gpurun_out/dist_gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu_nccl.py -x -q -m gpu > gpurun_out/pytest_nccl.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
tail -5 gpurun_out/pytest_nccl.log
: > gpurun_out/dist_bench.jsonl
run() {   # nproc nx env...
  local np=$1 nx=$2; shift 2
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29531 \
      tests/tools/dist_partitioned_check.py $nx $nx 5 > gpurun_out/dist_run.log 2>&1
  echo "exit $? np=$np nx=$nx $*"
  grep "^PARTITIONED " gpurun_out/dist_run.log | sed 's/^PARTITIONED //' >> gpurun_out/dist_bench.jsonl
  grep -v "^PARTITIONED" gpurun_out/dist_run.log | tail -5
}
for nx in $SIZES; do
  run 1 $nx ARAP_DIST_BLOCK_JACOBI=0
  np=2
  while [ $np -le $NGPU ]; do
    run $np $nx ARAP_DIST_BLOCK_JACOBI=0
    run $np $nx ARAP_DIST_BLOCK_JACOBI=1
    np=$((np * 2))
  done
done
cat gpurun_out/dist_bench.jsonl
