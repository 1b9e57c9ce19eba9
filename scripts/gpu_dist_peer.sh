#!/bin/bash
# scripts/gpu_dist_peer.sh -- partitioned solver on NGPU GPUs: NCCL against the peer-memory transport.
set -x
mkdir -p gpurun_out
NGPU=${NGPU:-2}
: > gpurun_out/dist_peer.jsonl
for nx in ${SIZES:-2000 4000}; do
  for tr in peer nccl; do
    ARAP_DIST_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29551 \
        tests/tools/dist_partitioned_check.py $nx $nx ${ITERS:-5} > gpurun_out/dist_run_$tr.log 2>&1
    echo "exit $? nx=$nx transport=$tr"
    grep "^PARTITIONED " gpurun_out/dist_run_$tr.log | sed 's/^PARTITIONED //' >> gpurun_out/dist_peer.jsonl
    grep -v "^PARTITIONED" gpurun_out/dist_run_$tr.log | grep -v "OMP_NUM\|^\*\*\*\*\|^$" | tail -6
  done
done
cat gpurun_out/dist_peer.jsonl
