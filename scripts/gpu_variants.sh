#!/bin/bash
# scripts/gpu_variants.sh -- A/B kernel variants (mesh_deform_b200/variants/*.so) on the headline workload.
mkdir -p gpurun_out
for lib in mesh_deform_b200/libarap_b200.so mesh_deform_b200/variants/*.so; do
  name=$(basename $lib .so)
  ARAP_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/variant_$name.json 2> gpurun_out/variant_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/variant_$name.json"))
    print("$name", "it/s %.1f ms %.3f cg %.1f" % (d["value"], d["ms_per_step"], d["cg"]["iterations_per_arap_iteration"]), " ".join("%s=%.1f" % (k, v["avg_us"]) for k, v in d["kernels"].items() if k in ("cg_spmv",)))
except Exception as e:
    print("$name FAILED", e)
PY
done
