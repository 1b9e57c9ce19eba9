#!/bin/bash
# scripts/gpu_partitioned_bench.sh -- bench.py --workload partitioned_grid on 1..NGPU GPUs (strong scaling of one grid mesh).
set -x
mkdir -p gpurun_out
N=${NGPU:-2}
NX=${NX:-4000}
: > gpurun_out/partitioned_bench.jsonl
timeout 900 python bench.py --workload partitioned_grid --nx $NX --steps 5 --warmup 3 2> gpurun_out/pb.err | grep '^{' >> gpurun_out/partitioned_bench.jsonl; tail -2 gpurun_out/pb.err
np=2
while [ $np -le $N ]; do
  for tr in ${TRANSPORTS:-nccl}; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29571 \
        bench.py --workload partitioned_grid --nx $NX --gpus $np --steps 5 --warmup 3 --transport $tr 2> gpurun_out/pb.err | grep '^{' >> gpurun_out/partitioned_bench.jsonl
    tail -2 gpurun_out/pb.err
  done
  np=$((np * 2))
done
cat gpurun_out/partitioned_bench.jsonl
