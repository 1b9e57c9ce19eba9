#!/bin/bash
# scripts/gpu_dist_quick.sh -- NCCL test + the global-hierarchy partitioned solver on NGPU GPUs for the given grid sizes.
set -x
mkdir -p gpurun_out
NGPU=${NGPU:-2}
timeout 600 python -m pytest tests/test_multi_gpu_nccl.py -x -q -m gpu > gpurun_out/pytest_nccl.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
tail -15 gpurun_out/pytest_nccl.log
: > gpurun_out/dist_quick.jsonl
for nx in ${SIZES:-2000 4000}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29541 \
      tests/tools/dist_partitioned_check.py $nx $nx 5 > gpurun_out/dist_run.log 2>&1
  echo "exit $? nx=$nx"
  grep "^PARTITIONED " gpurun_out/dist_run.log | sed 's/^PARTITIONED //' >> gpurun_out/dist_quick.jsonl
  grep -v "^PARTITIONED" gpurun_out/dist_run.log | tail -8
done
cat gpurun_out/dist_quick.jsonl
