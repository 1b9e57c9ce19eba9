#!/bin/bash
# round 2, second GPU call (1 GPU): all GPU tests, the new bench line, PDL mode 2, launch list with the host-driven loop
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests2.log
tail -15 gpurun_out/r02_tests2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench2_default.json 2> gpurun_out/r02_bench2_default.err; echo "bench rc=$?"
ARAP_PDL=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f32 > gpurun_out/r02_bench2_pdl2.json 2> gpurun_out/r02_bench2_pdl2.err
ARAP_STEP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/r02_b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-f32 > gpurun_out/ncu_bench2.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
for f in default pdl2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench2_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["cg"], d.get("parity"), d.get("cold_start"), d.get("f32"), d["clocks"], d["prepare_ms"])
except Exception as e:
    print("$f failed", e)
PY
done
