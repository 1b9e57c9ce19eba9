#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -k "partitioned" -x -q --timeout 150 --timeout-method=thread -s > gpurun_out/r02_tests3_part.log 2>&1; echo "part tests rc=$?"
tail -60 gpurun_out/r02_tests3_part.log
for v in default "ARAP_TILES=0" "ARAP_TILES=0 ARAP_TAIL=0" "ARAP_REORDER=0" "ARAP_TILES=0 ARAP_MG_DEVICE_SETUP=0" "ARAP_TILES=0 ARAP_B200_LIB=mesh_deform_b200/variants/libarap_rhsmb2.so"; do
  name=$(echo "$v" | tr ' =/.' '____')
  if [ "$v" = "default" ]; then ev=""; else ev="$v"; fi
  env $ev timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f32 > gpurun_out/r02_bench3_$name.json 2> gpurun_out/r02_bench3_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench3_$name.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("$v", round(d["value"],1), round(d["ms_per_step"],4), d["cg"]["iterations_per_arap_iteration"], d["cg"].get("kernel_launches_per_cg_iteration"), d["config"].get("one_ring_kernels"), d["config"].get("vertex_order"), "prepare", round(d["prepare_ms"]), d["prepare_host_setup_ms"], d.get("prepare_device_setup_ms"),
          {n: round(k[n]["avg_us"],1) for n in ("local_step","rhs_residual","cg_spmv","mg_fine_residual","mg_fine_postsmooth","cg_update_mg","mg_tail") if n in k})
except Exception as e:
    print("$v failed", e)
    import subprocess; print(subprocess.run(["tail","-5","gpurun_out/r02_bench3_$name.err"],capture_output=True,text=True).stdout)
PY
done
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread --deselect tests/test_gpu_parity.py::test_partitioned_mesh_in_process_matches_oracle > gpurun_out/r02_tests3_all.log 2>&1; echo "all tests rc=$?"
tail -25 gpurun_out/r02_tests3_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke3.log
