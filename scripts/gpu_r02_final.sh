#!/bin/bash
# round 2, final single-GPU record: the driver's own sequence (GPU tests with -x, smoke, default bench) + launch list + ncu --set full
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 --timeout-method=thread > gpurun_out/r02_final_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_c_bench.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("bench", round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["cg"], d["parity"], d["cold_start"], d["f32"], "prepare", round(d["prepare_ms"]), d["prepare_host_setup_ms"], d.get("prepare_device_setup_ms"), d["frame"]["ms"], d["clocks"])
    print({n: (round(k[n]["avg_us"],1), round(k[n].get("frac_of_peak") or 0,3)) for n in k})
    print(d["roofline"])
except Exception as e:
    print("bench parse failed", e)
PY
if [ -z "$NO_NCU" ]; then
ARAP_STEP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 1000 --csv --log-file gpurun_out/r02_c_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-f32 > gpurun_out/ncu_c_launches.log 2>&1
ARAP_STEP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"local_step_kernel|rhs_residual_kernel" \
    -s 2 -c 4 -o gpurun_out/r02_prof_local -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-f32 > gpurun_out/ncu_c_local.log 2>&1
ARAP_STEP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_spmv_z|cg_fused_update|mg_fine|mg_restrict|mg_csr|mg_prolong|mg_dense" \
    -s 60 -c 20 -o gpurun_out/r02_prof_hot -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-f32 > gpurun_out/ncu_c_hot.log 2>&1
fi
ls -la gpurun_out | tail -12
