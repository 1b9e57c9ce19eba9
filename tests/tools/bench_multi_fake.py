"""TEST TOOL (CPU, gloo): drives bench.py's multi-GPU arms -- their collectives, agreement logic, efficiency arithmetic and JSON
assembly -- with stand-ins for the engine classes, so that the host-side plumbing of `bench.py --gpus N` is exercised without
GPUs. Launched by tests/test_bench_multi_gloo.py under torch.distributed.run with 2 processes. The stand-ins 'deform' by moving
every free vertex of the grid by a fixed amount, identically in the partitioned and the single-GPU form, so parity is exact."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import bench                                                    # noqa: E402
from mesh_deform_b200 import capi, partition as PT              # noqa: E402


class _Timer:
    def timer_start(self):
        self._t = time.perf_counter()

    def timer_stop(self):
        return 1e3 * (time.perf_counter() - self._t) + 1.0

    def synchronize(self):
        pass

    def profile_enable(self, on=True):
        pass

    def profile_reset(self):
        pass

    def profile(self):
        return {"halo_exchange": {"launches": 8, "ms": 0.1}, "cg_spmv": {"launches": 1, "ms": 0.03}}


class FakeSingle(_Timer):
    def __init__(self, mesh, faces, precision=None, **kw):
        self.mesh, self.its = mesh, 0
        self.rest = mesh.copy()

    def setConstraints(self, idx, tgt):
        self.idx, self.tgt = np.asarray(idx), np.asarray(tgt)

    def prepare(self):
        return 0

    def iterate(self, n):
        self.its += n

    def positions(self, dtype=np.float64):
        p = self.rest + 0.001 * self.its
        p[self.idx] = self.tgt
        return p.astype(dtype)

    def energy(self):
        return 1.0 + self.its

    def solver_stats(self):
        return {"cg_iterations_total": 9 * self.its, "global_steps": max(1, self.its)}

    def close(self):
        pass


class FakePartitioned:
    def __init__(self, P, F, owner, rank, world, kind, ident, precision=np.float64, **kw):
        assert np.asarray(ident).size == 128
        self.part = PT.build_local_part(F, owner, rank, world)
        self.P, self.world = np.asarray(P, np.float64), world
        self.arap = _Timer()
        self.its = 0

    def setConstraints(self, idx, tgt):
        self.idx, self.tgt = np.asarray(idx), np.asarray(tgt)

    def prepare(self):
        return 0

    def iterate(self, n):
        self.its += n

    def solver_stats(self):
        return {"cg_iterations_total": 9 * self.its, "global_steps": max(1, self.its), "comm_exchanges_per_cg_iteration": 8,
                "comm_allreduces_per_cg_iteration": 2, "comm_halo_bytes_per_cg_iteration": 1234, "mg_levels": 4, "mg_global": 1, "cg_graph": 1,
                "setup_host_ms": 0.0, "setup_device_ms": 1.0}

    def comm_benchmark(self, rounds=200):
        return 12.0, 15.0

    def profile_enable(self, on=True):
        pass

    def profile_reset(self):
        pass

    def profile(self):
        return {"halo_exchange": {"launches": 8, "ms": 0.1}, "cg_spmv": {"launches": 1, "ms": 0.03}}

    def local_energy(self):
        return (1.0 + self.its) / self.world

    def owned_positions(self, dtype=np.float64):
        p = self.P + 0.001 * self.its
        p[self.idx] = self.tgt
        g = self.part.owned_global
        return g, p[g]


class FakeBatch(_Timer):
    def __init__(self, P, F, K, precision=np.float64, **kw):
        self.K, self.its = K, 0

    def setConstraints(self, idx, tgt):
        assert np.asarray(tgt).shape[0] == self.K

    def setRigidConstraints(self, idx, pts, T):
        assert np.asarray(T).shape == (self.K, 4, 4)

    def prepare(self):
        return 0

    def iterate(self, n):
        self.its += n

    def solver_stats(self):
        return {"cg_iterations_total": 2 * self.its, "global_steps": max(1, self.its), "mg_levels": 1}

    def close(self):
        pass


def main():
    rank, world, local_rank, dist = bench.dist_setup(2)
    capi.AsRigidAsPossibleDeformation = FakeSingle
    capi.PartitionedDeformation = FakePartitioned
    capi.BatchDeformation = FakeBatch
    capi.comm_unique_id = lambda: np.arange(128, dtype=np.uint8)
    args = argparse.Namespace(warmup=3, steps=4, transport="nccl", part_nx=int(sys.argv[1]), weak_verts_per_gpu=int(sys.argv[2]), oracle_nx=int(sys.argv[3]),
                              batch=8, multi_budget_s=int(sys.argv[4]))
    line = {"metric": "fake"} if rank == 0 else None
    multi = {}
    if line is not None:
        line["multi_gpu"] = multi
    wd = bench.Watchdog(rank, line, 120)
    wd.start()
    bench.multi_gpu_arms(args, rank, world, local_rank, dist, multi)
    wd.cancel()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("LINE " + json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
