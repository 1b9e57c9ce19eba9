#!/bin/bash
# round 2, first GPU call: tests, bench A/B (PDL / step graph), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests.log
tail -5 gpurun_out/r02_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
ARAP_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_nopdl.json 2> gpurun_out/r02_bench_nopdl.err
ARAP_STEP_GRAPH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_nostep.json 2> gpurun_out/r02_bench_nostep.err
ARAP_STEP_GRAPH=0 ARAP_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_neither.json 2> gpurun_out/r02_bench_neither.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/r02_a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for f in default nopdl nostep neither; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["cg"], d.get("parity"))
except Exception as e:
    print("$f failed", e)
PY
done
