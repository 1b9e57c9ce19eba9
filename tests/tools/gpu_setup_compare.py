"""Device-built vs host-built multigrid hierarchy on grids and icospheres: levels, operator complexity, CG iterations per ARAP
iteration, ms per iteration, setup time. usage: python tests/tools/gpu_setup_compare.py [grid:NX ...] [ico:NU ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, json, time, numpy as np
sys.path.insert(0, %r)
from mesh_deform_b200 import meshgen as G, capi
kind, n = sys.argv[1].split(":"); n = int(n)
if kind == "grid":
    P, F = G.grid_plane(n, n); idx, tgt = G.grid_constraints(n, n, P)
else:
    P, F = G.icosphere(n); idx, tgt = G.cap_constraints(P)
a = capi.AsRigidAsPossibleDeformation(P.copy(), F, np.float64)
a.setConstraints(idx, tgt)
t0 = time.perf_counter(); a.prepare(); a.synchronize(); prep = time.perf_counter() - t0
a.iterate(5); a.synchronize()
s0 = a.solver_stats()
a.timer_start(); a.iterate(20); ms = a.timer_stop()
s = a.solver_stats()
print(json.dumps({"mesh": sys.argv[1], "V": int(P.shape[0]), "levels": s["mg_levels"], "complexity": round(s["mg_operator_complexity"], 3),
                  "cg_its_per_step": (s["cg_iterations_total"] - s0["cg_iterations_total"]) / 20.0, "ms_per_step": ms / 20, "prepare_s": round(prep, 3),
                  "setup_device_ms": round(s["setup_device_ms"], 1), "setup_host_ms": round(s["setup_host_ms"], 1)}))
''' % ROOT

for mesh in sys.argv[1:] or ["grid:1000", "ico:316"]:
    for dev in (("1",) if os.environ.get("COMPARE_DEVICE_ONLY") else ("1", "0")):
        env = dict(os.environ, ARAP_MG_DEVICE_SETUP=dev)
        out = subprocess.run([sys.executable, "-c", CHILD, mesh], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        print("device" if dev == "1" else "host  ", line[0] if line else out.stderr[-500:], flush=True)
        if os.environ.get("ARAP_MG_TIMING"):
            print("".join(l + "\n" for l in out.stderr.splitlines() if "] level" in l), end="", flush=True)
