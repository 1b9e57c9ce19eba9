"""CG iterations per ARAP iteration of a grid partitioned in-process over `world` strips on ONE GPU, against the unpartitioned solver,
for the environment given by the caller. usage: python tests/tools/gpu_partition_iterations.py NX [WORLD]"""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mesh_deform_b200 import meshgen as G, capi, partition as PT

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P, F = G.grid_plane(nx, nx)
idx, tgt = G.grid_constraints(nx, nx, P)
owner = PT.strip_owner(P, world)
parts = [capi.PartitionedDeformation(P, F, owner, r, world, capi.TRANSPORT_IN_PROCESS, 4242, np.float64) for r in range(world)]
out = {}


def work(p):
    def run():
        p.setConstraints(idx, tgt)
        assert p.prepare() == capi.ARAP_OK
        p.iterate(5)
        s0 = p.solver_stats()["cg_iterations_total"]
        p.iterate(20)
        out[p.part.rank if hasattr(p.part, "rank") else id(p)] = (p.solver_stats()["cg_iterations_total"] - s0) / 20.0
    return run


t0 = time.perf_counter()
capi.run_partitions_in_process([work(p) for p in parts])
st = parts[0].solver_stats()
print("partitioned x%d" % world, "nx", nx, "cg its/step", sorted(set(out.values())), "levels", st["mg_levels"], "global", st["mg_global"],
      "exchanges", st["comm_exchanges_per_cg_iteration"], "allreduces", st["comm_allreduces_per_cg_iteration"], "%.1f s" % (time.perf_counter() - t0), flush=True)
if not os.environ.get("PARTITION_ONLY"):
    a = capi.AsRigidAsPossibleDeformation(P.copy(), F, np.float64)
    a.setConstraints(idx, tgt)
    a.prepare()
    a.iterate(5)
    s0 = a.solver_stats()["cg_iterations_total"]
    a.iterate(20)
    a.synchronize()
    print("single", "cg its/step", (a.solver_stats()["cg_iterations_total"] - s0) / 20.0, flush=True)
