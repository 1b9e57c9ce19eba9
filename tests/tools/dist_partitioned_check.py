"""Partitioned mode over NCCL, one process per GPU (launch with torchrun). Rank 0 checks the gathered result against
the CPU oracle (small mesh) and prints timing for a larger one.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/dist_partitioned_check.py [nx nz iters]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

from mesh_deform_b200 import capi, meshgen as G, partition as PT


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    nz = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    check = nx * nz <= 200000
    P, F = G.grid_plane(nx, nz)
    idx, tgt = G.grid_constraints(nx, nz, P)
    owner = PT.strip_owner(P, world)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.from_numpy(capi.comm_unique_id()))
    dist.broadcast(uid, 0)
    block_jacobi = os.environ.get("ARAP_DIST_BLOCK_JACOBI", "0") == "1"      # the per-rank preconditioner, for comparison
    kind = capi.TRANSPORT_PEER if os.environ.get("ARAP_DIST_TRANSPORT", "nccl") == "peer" else capi.TRANSPORT_NCCL
    p = capi.PartitionedDeformation(P, F, owner, rank, world, kind, uid.cpu().numpy(), np.float64,
                                    global_multigrid=not block_jacobi, device=local)
    p.setConstraints(idx, tgt)
    t0 = time.perf_counter()
    assert p.prepare() == capi.ARAP_OK
    prep = time.perf_counter() - t0
    p.iterate(1)                      # warm-up (NCCL connections)
    dist.barrier()
    torch.cuda.synchronize()
    p.arap.timer_start()
    p.iterate(iters)
    ms = p.arap.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gid, xyz = p.owned_positions()
    pieces = [None] * world
    dist.all_gather_object(pieces, (gid, xyz, p.local_energy(), p.solver_stats()))
    if rank == 0:
        pos = np.zeros_like(P)
        for g, x, _, _ in pieces:
            pos[g] = x
        out = {"world": world, "vertices": int(P.shape[0]), "iterations": iters, "ms_per_iteration": float(t.item()) / iters,
               "transport": "peer memory (direct stores + flags)" if kind == capi.TRANSPORT_PEER else "nccl",
               "preconditioner": "block-Jacobi multigrid per rank" if block_jacobi else "global multigrid, rows partitioned",
               "prepare_s": prep, "cg_graph": pieces[0][3]["cg_graph"], "mg_levels": pieces[0][3]["mg_levels"], "setup_host_ms": pieces[0][3]["setup_host_ms"], "cg_iterations_per_step": pieces[0][3]["cg_iterations_total"] / max(1, pieces[0][3]["global_steps"]),
               "halo_vertices_rank0": int(p.part.n_local - p.part.n_owned)}
        if check:
            from oracle import oracle as O
            omesh = P.copy()
            o = O.ArapOracle(omesh, F, np.float64)
            for i, tg in zip(idx, tgt):
                o.setConstraint(int(i), tg)
            o.deform(iters + 1)
            diag = float(np.linalg.norm(P.max(0) - P.min(0)))
            out["max_dp_over_diag"] = float(np.abs(pos - omesh).max() / diag)
            out["rel_dE"] = abs(sum(e for _, _, e, _ in pieces) - o.energy()) / o.energy()
            out["ok"] = bool(out["max_dp_over_diag"] <= 1e-5 and out["rel_dE"] <= 1e-6)
        print("PARTITIONED " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
