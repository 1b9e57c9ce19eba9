#!/bin/bash
# round 2, multi-GPU call: bench.py under torchrun (N = number of visible GPUs): first with small sizes (plumbing), then full size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
MODE=${2:-both}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
free -g | head -2 >> gpurun_out/multi_gpus.txt; nproc >> gpurun_out/multi_gpus.txt
run() {  # name, extra args...
  name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 "$@" \
      > gpurun_out/r02_multi_${name}_n$N.json 2> gpurun_out/r02_multi_${name}_n$N.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_multi_${name}_n$N.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
    m=d.get("multi_gpu") or {}
    for k,v in m.items():
        if isinstance(v,dict):
            print(" ", k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ("workload","limiter","preconditioner","single_gpu_workload","note")})
            if "limiter" in v: print("     limiter:", v["limiter"])
        else:
            print(" ", k, v)
except Exception as e:
    print("parse failed", e)
    import subprocess; print(subprocess.run(["tail","-30","gpurun_out/r02_multi_${name}_n$N.err"],capture_output=True,text=True).stdout)
PY
}
if [ "$MODE" = "small" ] || [ "$MODE" = "both" ]; then
  run small --part-nx 1000 --weak-verts-per-gpu 250000 --batch 512 --oracle-nx 300 --no-f32
fi
if [ "$MODE" = "full" ] || [ "$MODE" = "both" ]; then
  run full
fi
if [ "$MODE" = "peer" ]; then
  run peer --transport peer --part-nx 2000 --weak-verts-per-gpu 1000000 --batch 512 --oracle-nx 300 --no-f32
fi
if [ "$MODE" = "cmp" ]; then     # the two transports on the same mid-size workloads
  run ncclmid --transport nccl --part-nx 2000 --weak-verts-per-gpu 1000000 --batch 512 --oracle-nx 300 --no-f32 --no-cpu-baseline
  run peer --transport peer --part-nx 2000 --weak-verts-per-gpu 1000000 --batch 512 --oracle-nx 300 --no-f32 --no-cpu-baseline
fi
