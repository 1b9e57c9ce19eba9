#!/bin/bash
# A/B of the internal vertex order: ARAP_REORDER 0 = user order, 2 = forced renumbering; ARAP_PATCH = patch size (1 = pure Morton)
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/reorder.json 2> gpurun_out/reorder.err
  python - <<PY
import json
d=json.load(open("gpurun_out/reorder.json"))
print("$1", "it/s %.1f ms %.3f cg %.1f" % (d["value"], d["ms_per_step"], d["cg"]["iterations_per_arap_iteration"]), " ".join("%s=%.1f" % (k, v["avg_us"]) for k, v in d["kernels"].items() if k in ("cg_spmv","mg_fine_residual","mg_fine_postsmooth","local_step","rhs_residual","mg_restrict_presmooth","mg_prolong_add")))
PY
}
run "ARAP_REORDER=1"
run "ARAP_REORDER=2 ARAP_PATCH=1"
run "ARAP_REORDER=2 ARAP_PATCH=256"
run "ARAP_REORDER=2 ARAP_PATCH=1024"
run "ARAP_REORDER=2 ARAP_PATCH=4096"
