#!/bin/bash
mkdir -p gpurun_out
for v in 3 1.5 6 0.75; do
  ARAP_NNZ_PER_LANE=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/lanes.json 2> gpurun_out/lanes.err
  python - <<PY
import json
d=json.load(open("gpurun_out/lanes.json"))
print("NNZ_PER_LANE=$v", "it/s %.1f ms %.3f" % (d["value"], d["ms_per_step"]), " ".join("%s=%.1f" % (k, v["avg_us"]) for k, v in d["kernels"].items() if k.startswith("mg_")))
PY
done
