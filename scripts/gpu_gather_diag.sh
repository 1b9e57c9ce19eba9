#!/bin/bash
# Diagnostic: kernel times when every gather hits the row itself (self1) or the next rows (self2) -> no inter-CTA re-fetch.
for lib in mesh_deform_b200/libarap_b200.so mesh_deform_b200/variants/libarap_self1.so mesh_deform_b200/variants/libarap_self2.so; do
ARAP_B200_LIB=$PWD/$lib timeout 300 python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
from mesh_deform_b200 import capi, meshgen as G
P, F = G.icosphere(316)
idx, tgt = G.cap_constraints(P)
a = capi.AsRigidAsPossibleDeformation(P.copy(), F, np.float64, max_cg_iterations=6)
a.setConstraints(idx, tgt); a.prepare()
try:
    a.iterate(2)
    a.profile_enable(True); a.profile_reset()
    a.iterate(4)
except Exception as e:
    print("note:", e)
p = a.profile()
print("$lib".split("/")[-1], " ".join("%s=%.1f" % (k, 1e3 * v["ms"] / v["launches"]) for k, v in p.items() if k in ("local_step","rhs_residual","cg_spmv","mg_fine_residual","mg_fine_postsmooth")))
PY
done
