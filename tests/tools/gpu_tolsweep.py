"""CG tolerance sweep: parity of the GPU engine vs the oracle as a function of cg_tolerance (run on the GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from mesh_deform_b200 import meshgen as G, capi
from oracle import oracle as O

nu, iters = int(sys.argv[1]), int(sys.argv[2])
P, F = G.icosphere(nu)
idx, tgt = G.cap_constraints(P)
diag = float(np.linalg.norm(P.max(0) - P.min(0)))
omesh = P.copy()
o = O.ArapOracle(omesh, F, np.float64)
for i, t in zip(idx, tgt):
    o.setConstraint(int(i), t)
o.deform(iters)
Eo = o.energy()
for tol in (1e-10, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4):
    mesh = P.copy()
    a = capi.AsRigidAsPossibleDeformation(mesh, F, np.float64, cg_tolerance=tol)
    a.setConstraints(idx, tgt)
    a.prepare()
    a.timer_start(); a.iterate(iters); ms = a.timer_stop()
    pos = a.positions()
    st = a.solver_stats()
    print(json.dumps({"nu": nu, "V": len(P), "iters": iters, "tol": tol, "max_dp_over_diag": float(np.abs(pos - omesh).max() / diag),
                      "rel_dE": abs(a.energy() - Eo) / Eo, "cg_its_per_step": st["cg_iterations_total"] / st["global_steps"], "ms_per_step": ms / iters}))
