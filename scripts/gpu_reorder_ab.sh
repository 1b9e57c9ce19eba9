#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
ARAP_REORDER=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
for t in 1 0; do
  ARAP_REORDER=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/reorder_$t.json 2> gpurun_out/reorder_$t.err
  python - <<PY
import json
d=json.load(open("gpurun_out/reorder_$t.json"))
print("ARAP_REORDER=$t", "it/s %.1f ms %.3f cg %.1f prepare %.0f" % (d["value"], d["ms_per_step"], d["cg"]["iterations_per_arap_iteration"], d["prepare_ms"]), " ".join("%s=%.1f" % (k, v["avg_us"]) for k, v in d["kernels"].items() if "avg_us" in v and v["share"] > 0.008))
PY
done
