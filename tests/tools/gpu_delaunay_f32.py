import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mesh_deform_b200 import capi
from oracle import oracle as O
from test_gpu_parity import delaunay_patch
P, F = delaunay_patch(6000, 21)
rng = np.random.default_rng(4)
idx = rng.choice(len(P), 60, replace=False).astype(np.int32)
tgt = P[idx] + 0.05 * rng.standard_normal((60, 3))
diag = float(np.linalg.norm(P.max(0) - P.min(0)))
res = {}
for prec in (np.float32, np.float64):
    m = P.astype(prec); o = O.ArapOracle(m, F, prec)
    for i, t in zip(idx, tgt): o.setConstraint(int(i), t)
    o.deform(4); res["oracle", prec] = m.astype(np.float64)
    for solver in (1, 2):
        m2 = P.astype(prec); a = capi.AsRigidAsPossibleDeformation(m2, F, prec, solver=solver)
        a.setConstraints(idx, tgt); a.deform(4); res["engine%d" % solver, prec] = m2.astype(np.float64)
keys = list(res)
for i in range(len(keys)):
    for j in range(i + 1, len(keys)):
        print(keys[i][0], np.dtype(keys[i][1]).name, "vs", keys[j][0], np.dtype(keys[j][1]).name, "%.3e" % (np.abs(res[keys[i]] - res[keys[j]]).max() / diag))
