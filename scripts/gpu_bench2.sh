#!/bin/bash
# scripts/gpu_bench2.sh -- what the driver's scaling run does at N GPUs: both bench arms under torchrun.
set -x
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
timeout 900 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench${N}_reference.json 2> gpurun_out/bench${N}_reference.err; echo "reference exit $?"
grep '^{' gpurun_out/bench${N}_reference.json | cut -c1-400
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench${N}.json 2> gpurun_out/bench${N}.err; echo "bench exit $?"
grep '^{' gpurun_out/bench${N}.json | cut -c1-900; tail -3 gpurun_out/bench${N}.err
timeout 900 $TR bench.py --workload batch_spheres --gpus $N --steps 10 --warmup 3 > gpurun_out/bench${N}_batch.json 2> gpurun_out/bench${N}_batch.err; echo "batch exit $?"
grep '^{' gpurun_out/bench${N}_batch.json | cut -c1-600; tail -3 gpurun_out/bench${N}_batch.err
