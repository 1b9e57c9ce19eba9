#!/bin/bash
# Diagnostic: local_step with only its gathers (diag1) / only its arithmetic (diag2): which half bounds it?
for lib in mesh_deform_b200/libarap_b200.so mesh_deform_b200/variants/libarap_diag1.so mesh_deform_b200/variants/libarap_diag2.so; do
ARAP_B200_LIB=$PWD/$lib timeout 300 python - <<PY
import sys, numpy as np
sys.path.insert(0, ".")
from mesh_deform_b200 import capi, meshgen as G
P, F = G.icosphere(316)
idx, tgt = G.cap_constraints(P)
a = capi.AsRigidAsPossibleDeformation(P.copy(), F, np.float64, max_cg_iterations=3)
a.setConstraints(idx, tgt); a.prepare()
a.iterate(2)
a.profile_enable(True); a.profile_reset()
try:
    a.iterate(5)
except Exception as e:
    print("note:", e)
p = a.profile()
print("$lib".split("/")[-1], "local_step avg us", 1e3 * p["local_step"]["ms"] / p["local_step"]["launches"], "redo", 1e3 * p["local_step_redo"]["ms"] / p["local_step_redo"]["launches"])
PY
done
