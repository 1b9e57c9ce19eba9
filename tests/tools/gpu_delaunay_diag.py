"""Diagnostic: parity of the engine on an irregular Delaunay patch for several solver settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from mesh_deform_b200 import capi
from oracle import oracle as O
from test_gpu_parity import delaunay_patch

P, F = delaunay_patch(6000, 21)
rng = np.random.default_rng(4)
idx = rng.choice(len(P), 60, replace=False).astype(np.int32)
tgt = P[idx] + 0.05 * rng.standard_normal((60, 3))
diag = float(np.linalg.norm(P.max(0) - P.min(0)))
for iters in (1, 4):
    omesh = P.copy()
    o = O.ArapOracle(omesh, F, np.float64)
    for i, t in zip(idx, tgt):
        o.setConstraint(int(i), t)
    o.deform(iters)
    for name, kw in (("mg 1e-6", dict(solver=2)), ("mg 1e-8", dict(solver=2, cg_tolerance=1e-8)), ("mg 1e-10", dict(solver=2, cg_tolerance=1e-10)),
                     ("jacobi 1e-9", dict(solver=1)), ("jacobi 1e-12", dict(solver=1, cg_tolerance=1e-12))):
        mesh = P.copy()
        a = capi.AsRigidAsPossibleDeformation(mesh, F, np.float64, **kw)
        a.setConstraints(idx, tgt)
        a.deform(iters)
        d = np.abs(mesh - omesh).max(1)
        st = a.solver_stats()
        print(iters, name, "err/diag %.3e" % (d.max() / diag), "argmax", int(d.argmax()), "n>1e-6", int((d > 1e-6 * diag).sum()),
              "cg its", st["cg_iterations_total"], "relres %.2e" % st["last_relative_residual"], "dE %.2e" % (abs(a.energy() - o.energy()) / o.energy()))
rp, ci, w = a.cotanWeights()
print("weights min %.3e max %.3e  clamped(<=1e-9): %d of %d" % (w.min(), w.max(), int((w <= 1e-9).sum()), w.size))
