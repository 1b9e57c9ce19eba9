#!/bin/bash
# A/B: TMA-staged SpMV (ARAP_TMA=1) vs plain, with the parity tests under both.
mkdir -p gpurun_out
ARAP_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for t in 0 1; do
  ARAP_TMA=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tma_$t.json 2> gpurun_out/tma_$t.err
  python - <<PY
import json
d=json.load(open("gpurun_out/tma_$t.json"))
print("ARAP_TMA=$t", "it/s %.1f ms %.3f cg %.1f" % (d["value"], d["ms_per_step"], d["cg"]["iterations_per_arap_iteration"]), " ".join("%s=%.1f" % (k, v["avg_us"]) for k, v in d["kernels"].items() if k in ("cg_spmv","mg_fine_residual","mg_fine_postsmooth")))
PY
done
