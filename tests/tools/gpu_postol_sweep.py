"""Stopping-rule sweep: parity of the engine vs the oracle after N ARAP iterations as a function of position_tolerance
(and the old residual rule for comparison), on meshes of very different conditioning. Run on the GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mesh_deform_b200 import meshgen as G, capi
from oracle import oracle as O
from test_gpu_parity import delaunay_patch


def cases():
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "arap_golden.npz"))
    yield "bar", z["bar_V"], z["bar_F"], g["bar_idx"], g["bar_tgt"], 10
    P, F = G.icosphere(128)
    idx, tgt = G.cap_constraints(P)
    yield "icosphere163k", P, F, idx, tgt, 20
    P, F = G.grid_plane(300, 300)
    idx, tgt = G.grid_constraints(300, 300, P)
    yield "grid90k", P, F, idx, tgt, 20
    for n, seed in ((6000, 21), (60000, 5)):
        P, F = delaunay_patch(n, seed)
        rng = np.random.default_rng(4)
        idx = rng.choice(len(P), max(60, n // 100), replace=False).astype(np.int32)
        tgt = P[idx] + 0.05 * rng.standard_normal((idx.size, 3))
        yield "delaunay%d" % n, P, F, idx, tgt, 20


for name, P, F, idx, tgt, iters in cases():
    P = np.ascontiguousarray(P, np.float64)
    diag = float(np.linalg.norm(P.max(0) - P.min(0)))
    omesh = P.copy()
    o = O.ArapOracle(omesh, F, np.float64)
    for i, t in zip(idx, tgt):
        o.setConstraint(int(i), t)
    o.deform(iters)
    Eo = o.energy()
    rules = [("res 1e-6", dict(cg_tolerance=1e-6)), ("res 1e-8", dict(cg_tolerance=1e-8))]
    for pt in (1e-7, 3e-8):
        for et in (-1.0, 1e-7, 1e-8, 1e-9):
            rules.append(("pos %.0e en %s" % (pt, "off" if et < 0 else "%.0e" % et), dict(position_tolerance=pt, energy_tolerance=et)))
    rules.append(("default", dict()))
    for label, kw in rules:
        mesh = P.copy()
        a = capi.AsRigidAsPossibleDeformation(mesh, F, np.float64, **kw)
        a.setConstraints(idx, tgt)
        a.prepare()
        a.timer_start(); a.iterate(iters); ms = a.timer_stop()
        pos = a.positions()
        st = a.solver_stats()
        print(json.dumps({"mesh": name, "V": len(P), "iters": iters, "rule": label, "max_dp_over_diag": float("%.3e" % (np.abs(pos - omesh).max() / diag)),
                          "rel_dE": float("%.3e" % (abs(a.energy() - Eo) / Eo)), "cg_its_per_step": round(st["cg_iterations_total"] / st["global_steps"], 2),
                          "ms_per_step": round(ms / iters, 3), "last_pos_err": float("%.2e" % st["last_position_error"])}), flush=True)
