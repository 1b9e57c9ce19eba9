#!/usr/bin/env python
"""bench.py -- the ARAP hot path on the headline workload (BASELINE.json configs[2]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nu 316] [--precision f64|f32]

A "step" is ONE ARAP iteration (local step + global step, reference inc/deform/arap.h:122-129) on the
class-I geodesic icosphere nu=316 (V = 998,562; 5 % lowest-z vertices anchored, 1 % highest-z vertices
dragged by Rz(30 deg) + (0,0,0.3)), PrecisionType double. Prints ONE JSON line (see DESIGN.md "Measurement").

  value     ARAP iterations/s, state resident in HBM, K steps timed with CUDA events on the engine's stream
  e2e       the same through the public C-ABI call arap_deform(host_mesh, 1): one iteration plus the
            write-back of p' into a pinned HOST mesh buffer every step (the reference's deform(1), arap.h:101-138)
  roofline  the dominant kernel of the step (largest share of step time), algorithmic bytes / CUDA-event time
  cpu_baseline  the oracle (C restatement of arap.h, 1 thread) timed on this box's host cores

N > 1 (torchrun): the batched-independent-deformations sharding -- every rank deforms its own 1M-vertex mesh,
no data-path collective; value = N x K iterations / max-over-ranks time.

--impl reference: the CPU oracle alone (the reference itself cannot be built here: no Eigen), same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "arap_iterations_per_sec_1M_verts"
UNIT = "iterations/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", p
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def build_workload(nu):
    from mesh_deform_b200 import meshgen as G
    P, F = G.icosphere(nu)
    idx, tgt = G.cap_constraints(P)
    return P, F, idx, tgt


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms from before the warm-up until after the end-to-end arm."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_load, smax, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            try:
                if len(parts) > 8 and float(parts[8]) >= 10.0:
                    sm_load.append(float(parts[0]))
            except ValueError:
                pass
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        under = sm_load if sm_load else sm
        return {"sm_mhz": float(np.median(under)) if under else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(sm_load),
                "period_ms": 50, "span": "whole run: prepare, warm-up, timed window, per-kernel pass, end-to-end arm, frames"}


def algorithmic_bytes(V, nF, nnz, s):
    """Per-launch algorithmic bytes (SURVEY.md section 8d / BASELINE.md section 3), s = sizeof(scalar).
    Every array a kernel must touch is counted once; neighbour gathers are assumed to be served by cache."""
    d = nnz / V
    return {
        "local_step": V * (15 * s + 4 + (4 + s) * d),              # p, p' read, R written (as 9 scalars), CSR
        "rhs_residual": V * (18 * s + 8 + (4 + s) * d),
        # matrix-free SpMV on the one-ring CSR: w = A z, z the fp32 V-cycle output (float4), w fp64 (3 x 8 B per row), 1-byte row mask
        "cg_spmv": nF * ((s + 4) * d + 4 + 1 + 16 + 24),
        "cg_update": nF * (6 * 24 + 8),
        "cg_direction": nF * (3 * 24 + 8),
        "apply_update": nF * (24 + 2 * 3 * s),
        # multigrid V-cycle, fine level: fp32 weights + float4 vectors, fp64 CG residual as right-hand side
        "mg_fine_residual": nF * ((4 + 4) * d + 4 + 1 + 24 + 16 + 16),
        "mg_fine_postsmooth": nF * ((4 + 4) * d + 4 + 1 + 8 + 24 + 16 + 16),
        # fused CG update: reads z (16), w, d, s, x, r (5 x 24) and 1/diag (8); writes d, s, x, r (4 x 24) and x0 (16)
        "cg_update_mg": nF * (16 + 5 * 24 + 8 + 4 * 24 + 16),
    }


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
            dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        else:                       # CPU tests of the host logic (tests/test_bench_multi_gloo.py); bench.py itself needs GPUs
            dist_mod.init_process_group(backend="gloo")
        dist = dist_mod
    return rank, world, local_rank, dist


def _dev(local_rank):
    """Tensor device of the collectives: the rank's GPU, or the CPU in the gloo tests of the host logic."""
    import torch
    return torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")


def barrier_and_sync(dist):
    import torch
    if dist is not None:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(dist, local_rank, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=_dev(local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_cpu_oracle(P, F, idx, tgt, iters, real):
    """prepare + `iters` iterations of the oracle; returns (seconds per phase dict, mesh, energy, oracle)."""
    from oracle import oracle as O
    mesh = P.astype(real)
    o = O.ArapOracle(mesh, F, real)
    for i, t in zip(idx, tgt):
        o.setConstraint(int(i), t)
    t0 = time.perf_counter()
    ok = o.deform(0)
    t_prepare = time.perf_counter() - t0
    assert ok
    return o, mesh, t_prepare


def reference_arm(args):
    """--impl reference: the reference's CPU path (oracle restatement; the reference itself needs Eigen,
    which is not in this image) on the same config/metric. Single thread: the reference has no threading."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    real = np.float64 if args.precision == "f64" else np.float32
    P, F, idx, tgt = build_workload(args.nu)
    o, mesh, t_prepare = run_cpu_oracle(P, F, idx, tgt, 0, real)
    for _ in range(args.warmup):
        o.deform(1)
    o.reset_timers()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.deform(1)
    dt = time.perf_counter() - t0
    tm = o.timers()
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"icosphere nu={args.nu} V={P.shape[0]} 5% anchors + 1% handles (BASELINE.json configs[2])",
                   "vertices": int(P.shape[0]), "faces": int(F.shape[0]), "n_free": int(o.nFree)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"full workload, {args.steps} ARAP iterations after prepare (prepare {t_prepare:.1f} s incl. "
                                   f"LDL^T factor, excluded like the GPU arm's prepare); per-iteration seconds: "
                                   f"local {tm['local'] / args.steps:.3f} rhs {tm['rhs'] / args.steps:.3f} solve {tm['solve'] / args.steps:.3f}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "prepare_s": t_prepare, "factor_nnz": o.factor_nnz(),
    }
    print(json.dumps(line), flush=True)
    return 0


def batch_arm(args):
    """--workload batch_spheres: BASELINE.json configs[3] -- K independent deformations of the reference's sphere mesh
    (642 vertices), one handle pose per trajectory key frame, 10 iterations each, sharded contiguously over the ranks
    with no data-path collective. Informational: the headline line is the icosphere workload."""
    rank, world, local_rank, dist = dist_setup(args.gpus)
    import torch
    torch.cuda.set_device(local_rank)
    from mesh_deform_b200 import capi, meshgen as G
    from mesh_deform_b200.sharding import shard_range
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    P, F = z["sphere_V"], z["sphere_F"]
    K = args.batch
    begin, end = shard_range(K, rank, world)
    handles = np.sort(np.unique(F[(F == G.SPHERE_HANDLE).any(1)]))

    def T(t=(0, 0, 0), R=np.eye(3)):
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = t
        return M
    traj = capi.TrajectorySE3()                # C ABI arap_trajectory_*: the product's own host-side trajectory
    prev = np.eye(4)
    for step in (T(), T((0.25, 0, 0)), T((0.5, 0, 0)), T(R=G.rot_x(np.pi / 2))):
        prev = prev @ step
        traj.addKeyPose(prev)
    util = capi.DeformationUtil(P, handles, origin=traj(0.0))
    poses = traj.sample(np.arange(begin, end) / max(1, K - 1))
    bdef = capi.BatchDeformation(P, F, end - begin, np.float64, device=local_rank)
    anchor = np.array([G.SPHERE_ANCHOR], np.int32)
    bdef.setConstraints(anchor, np.repeat(P[anchor][None], end - begin, 0))
    util.updateConstraints(poses, bdef)        # one call: every member's handle targets are computed on the device
    t0 = time.perf_counter()
    bdef.prepare()
    prepare_ms = 1e3 * (time.perf_counter() - t0)
    bdef.iterate(args.warmup)
    barrier_and_sync(dist)
    bdef.timer_start()
    bdef.iterate(args.steps)
    ms = bdef.timer_stop()
    barrier_and_sync(dist)
    ms = max_over_ranks(dist, local_rank, ms)
    stats = bdef.solver_stats()
    # second pass with every launch timed by CUDA events: the preconditioner GEMM (packing of the residuals + tcgen05 kernel)
    bdef.profile_enable(True)
    bdef.profile_reset()
    bdef.iterate(args.steps)
    prof = bdef.profile()
    bdef.profile_enable(False)
    if rank != 0:
        return 0
    _, _, peaks = load_peaks()
    gemm = prof.get("mg_dense_solve")
    roofline = None
    if gemm and stats["mg_levels"] == 1:
        members, V = end - begin, P.shape[0]
        useful = 2.0 * V * V * 3.0 * members                       # flops of Z = Inv . R per application
        avg_s = gemm["ms"] * 1e-3 / gemm["launches"]
        peak = float(peaks.get("bf16_tflops", 2250.0)) / 2.0         # TF32 runs at half the dense bf16 rate
        roofline = {"kernel": "batch_pack_b + mg_batch_dense_tc (tcgen05.mma kind::tf32, accumulator in TMEM)", "bound": "tensor",
                    "achieved": 3.0 * useful / avg_s * 1e-12, "peak": peak, "unit": "TFLOP/s", "frac": 3.0 * useful / avg_s * 1e-12 / peak,
                    "useful_tflops": useful / avg_s * 1e-12, "traffic": None, "avg_launch_us": avg_s * 1e6, "launches_per_step": gemm["launches"] / args.steps,
                    "note": "achieved counts the three TF32 products issued per fp32-grade product (hi.hi + lo.hi + hi.lo); useful_tflops counts one; "
                            "peak = half the measured dense bf16 rate (MEASURED_PEAKS.json); the time includes packing the fp64 residuals"}
    line = {"metric": "batch_sphere_deformation_iterations_per_sec", "value": K * args.steps / (ms * 1e-3), "unit": "member-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{K} independent deformations of sphere.obj (642 vertices), trajectory key-frame handle poses (BASELINE.json configs[3])",
                       "members_per_rank": end - begin, "vertices_total": int(K * P.shape[0]),
                       "preconditioner": "one member's dense inverse applied to all members (tcgen05 3xTF32 GEMM)" if stats["mg_levels"] == 1 else
                       "multigrid over the block-diagonal batch, %d levels" % stats["mg_levels"]},
            "deformations_of_10_iterations_per_sec": K / (10 * ms / args.steps * 1e-3),
            "cg_iterations_per_step": stats["cg_iterations_total"] / max(1, stats["global_steps"]), "prepare_ms": prepare_ms,
            "roofline": roofline}
    print(json.dumps(line), flush=True)
    return 0


def partitioned_arm(args):
    """--workload partitioned_grid: BASELINE.json configs[4] -- ONE nx x nx grid mesh (plane.obj topology, 2+2 constraint
    columns) partitioned into strips over the ranks: halo exchange of p', R and the CG direction, all-reduced CG scalars, one
    global multigrid hierarchy with its rows spread over the ranks. Strong scaling: the mesh is fixed,
    value = ARAP iterations/s of the whole mesh. Informational: the headline line is the icosphere workload."""
    rank, world, local_rank, dist = dist_setup(args.gpus)
    import torch
    torch.cuda.set_device(local_rank)
    from mesh_deform_b200 import capi, meshgen as G, partition as PT
    nx = args.nx
    P, F = G.grid_plane(nx, nx)
    idx, tgt = G.grid_constraints(nx, nx, P)
    owner = PT.strip_owner(P, world)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if world > 1:
        if rank == 0:
            uid.copy_(torch.from_numpy(capi.comm_unique_id()))
        dist.broadcast(uid, 0)
        kind = capi.TRANSPORT_PEER if args.transport == "peer" else capi.TRANSPORT_NCCL
        ident = uid.cpu().numpy()
    else:
        kind, ident = capi.TRANSPORT_IN_PROCESS, 7      # a single partition: same code path, no peer
    part = capi.PartitionedDeformation(P, F, owner, rank, world, kind, ident, np.float64, device=local_rank)
    part.setConstraints(idx, tgt)
    t0 = time.perf_counter()
    part.prepare()
    prepare_ms = 1e3 * (time.perf_counter() - t0)
    part.iterate(args.warmup)
    barrier_and_sync(dist)
    part.arap.timer_start()
    part.iterate(args.steps)
    ms = part.arap.timer_stop()
    barrier_and_sync(dist)
    ms = max_over_ranks(dist, local_rank, ms)
    stats = part.solver_stats()
    if rank != 0:
        return 0
    line = {"metric": "arap_iterations_per_sec_partitioned_grid", "value": args.steps / (ms * 1e-3), "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"one {nx} x {nx} grid mesh ({nx * nx} vertices) in {world} strips (BASELINE.json configs[4])",
                       "transport": args.transport if world > 1 else "none", "halo_vertices_rank0": int(part.part.n_local - part.part.n_owned),
                       "preconditioner": "global multigrid hierarchy, rows partitioned" if stats["mg_global"] else "per-rank multigrid",
                       "mg_levels": stats["mg_levels"], "cg_graph": bool(stats["cg_graph"])},
            "cg_iterations_per_step": stats["cg_iterations_total"] / max(1, stats["global_steps"]),
            "prepare_ms": prepare_ms, "hierarchy_setup_host_ms": stats["setup_host_ms"]}
    print(json.dumps(line), flush=True)
    return 0


# =====================================================================================================================
# Multi-GPU workloads (WORLD_SIZE > 1): BASELINE.json configs[3] and configs[4] in the same invocation as the headline
# =====================================================================================================================
def _sync_ok(dist, local_rank, ok):
    """All ranks agree on whether an arm worked (an exception on one rank must not leave the others in a collective)."""
    import torch
    t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=_dev(local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item() > 0.5)


def _nccl_id(dist, rank, capi, local_rank=0):
    import torch
    uid = torch.zeros(128, dtype=torch.uint8, device=_dev(local_rank))
    if rank == 0:
        uid.copy_(torch.from_numpy(capi.comm_unique_id()))
    dist.broadcast(uid, 0)
    return uid.cpu().numpy()


def run_partitioned(dist, rank, world, local_rank, P, F, idx, tgt, owner, warmup, steps, transport, gather=True, comm_rounds=200, profile_steps=3):
    """One nx x nz grid over `world` partitions: W warm-up + K timed ARAP iterations. Returns a dict (every rank; the gathered
    positions only on rank 0)."""
    import torch
    from mesh_deform_b200 import capi
    kind = capi.TRANSPORT_PEER if transport == "peer" else capi.TRANSPORT_NCCL
    part = capi.PartitionedDeformation(P, F, owner, rank, world, kind, _nccl_id(dist, rank, capi, local_rank), np.float64, device=local_rank)
    part.setConstraints(idx, tgt)
    t0 = time.perf_counter()
    part.prepare()
    prepare_s = time.perf_counter() - t0
    part.iterate(warmup)
    barrier_and_sync(dist)
    part.arap.timer_start()
    part.iterate(steps)
    ms = part.arap.timer_stop()
    barrier_and_sync(dist)
    ms = max_over_ranks(dist, local_rank, ms)
    stats = part.solver_stats()
    us_ex, us_ar = part.comm_benchmark(comm_rounds)
    us_ex, us_ar = max_over_ranks(dist, local_rank, us_ex), max_over_ranks(dist, local_rank, us_ar)
    dev = _dev(local_rank)
    e = torch.tensor([part.local_energy()], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    halo = torch.tensor([float(part.part.n_local - part.part.n_owned)], dtype=torch.float64, device=dev)
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    positions = None
    if gather:
        gid, xyz = part.owned_positions()
        full = torch.zeros((P.shape[0], 3), dtype=torch.float64, device=dev)
        full[torch.from_numpy(np.ascontiguousarray(gid)).to(dev)] = torch.from_numpy(xyz).to(dev)
        dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            positions = full.cpu().numpy()
        del full
    # where one CG iteration goes: a few more ARAP iterations launch by launch, every kernel, halo exchange and all-reduce bracketed by
    # CUDA events (rank 0's view; a rank that waits for a neighbour inside an exchange books the wait there -- that IS the cost)
    phases = None
    if profile_steps > 0:
        phases = phase_profile(part, part.solver_stats, profile_steps)
        barrier_and_sync(dist)
    res = {"iterations_per_s": steps / (ms * 1e-3), "ms_per_step": ms / steps, "energy": float(e.item()),
           "cg_iterations_per_step": stats["cg_iterations_total"] / max(1, stats["global_steps"]),
           "exchanges_per_cg_iteration": stats["comm_exchanges_per_cg_iteration"],
           "allreduces_per_cg_iteration": stats["comm_allreduces_per_cg_iteration"],
           "halo_bytes_sent_per_cg_iteration_rank0": stats["comm_halo_bytes_per_cg_iteration"],
           "halo_vertices_max": int(halo.item()), "us_per_exchange": us_ex, "us_per_allreduce": us_ar,
           "mg_levels": stats["mg_levels"], "mg_global": bool(stats["mg_global"]), "cg_graph": stats["cg_graph"],
           "prepare_s": prepare_s, "hierarchy_setup_host_ms": stats["setup_host_ms"], "hierarchy_setup_device_ms": stats["setup_device_ms"],
           "transport": transport, "us_per_cg_iteration_by_phase_rank0": phases, "positions": positions}
    del part
    return res


PHASE_GROUPS = {"halo_exchange": ("halo_exchange",), "allreduce_cg_scalars": ("allreduce_scalars", "cg_finalize"), "allreduce_replicated_level_rhs": ("allreduce_level",),
                "local_step + rhs_residual + apply_update": ("local_step", "local_step_redo", "rhs_residual", "apply_update"),
                "cg_spmv + cg_update": ("cg_spmv", "cg_update_mg"), "multigrid_fine_level": ("mg_fine_residual", "mg_fine_postsmooth"),
                "multigrid_coarse_levels": ("mg_csr_residual", "mg_restrict_presmooth", "mg_prolong_add", "mg_csr_postsmooth", "mg_dense_solve", "mg_tail")}


def phase_profile(handle, stats_fn, steps):
    """`steps` more ARAP iterations launch by launch with every kernel / exchange / all-reduce bracketed by CUDA events -> us per CG
    iteration by phase (the event brackets add ~5 us per launch: the SHARES are the information)."""
    its0 = stats_fn()["cg_iterations_total"]
    handle.profile_enable(True)
    handle.profile_reset()
    handle.iterate(steps)
    prof = handle.profile()
    handle.profile_enable(False)
    cg_its = max(1, stats_fn()["cg_iterations_total"] - its0)
    phases = {g: round(1e3 * sum(prof.get(n, {}).get("ms", 0.0) for n in names) / cg_its, 1) for g, names in PHASE_GROUPS.items()}
    phases["sum_of_phases_us"] = round(sum(phases.values()), 1)
    phases["launches_per_cg_iteration"] = round(sum(v["launches"] for v in prof.values()) / cg_its, 1)
    return phases


def run_single_gpu_grid(P, F, idx, tgt, local_rank, warmup, steps):
    """The same grid problem on ONE GPU with the unpartitioned solver (rank 0 only): the N = 1 point of the scaling curves."""
    from mesh_deform_b200 import capi
    mesh = np.ascontiguousarray(P, dtype=np.float64).copy()
    a = capi.AsRigidAsPossibleDeformation(mesh, F, np.float64, device=local_rank)
    a.setConstraints(idx, tgt)
    t0 = time.perf_counter()
    a.prepare()
    a.synchronize()
    prepare_s = time.perf_counter() - t0
    a.iterate(warmup)
    a.synchronize()
    a.timer_start()
    a.iterate(steps)
    ms = a.timer_stop()
    st = a.solver_stats()
    res = {"iterations_per_s": steps / (ms * 1e-3), "ms_per_step": ms / steps, "energy": a.energy(),
           "cg_iterations_per_step": st["cg_iterations_total"] / max(1, st["global_steps"]), "prepare_s": prepare_s,
           "positions": a.positions(np.float64)}
    res["us_per_cg_iteration_by_phase"] = phase_profile(a, a.solver_stats, 3)
    a.close()
    return res


def _parity(a_pos, a_energy, b_pos, b_energy, P):
    diag = float(np.linalg.norm(P.max(0) - P.min(0)))
    return {"max_dp_over_bbox_diag": float(np.abs(a_pos - b_pos).max() / diag),
            "rel_energy_diff": abs(a_energy - b_energy) / abs(b_energy) if b_energy else None, "tolerance": {"dp": 1e-5, "dE": 1e-6}}


def _strip(res):
    return {k: v for k, v in res.items() if k != "positions"}


class Watchdog:
    """Prints rank 0's JSON line and ends the process if the guarded section does not finish in `seconds` (a rank stuck in a
    collective cannot be interrupted from Python; a timer thread can still print and leave)."""

    def __init__(self, rank, line, seconds):
        self.rank, self.line, self.seconds = rank, line, seconds
        self.timer = None

    def _fire(self):
        if self.rank == 0 and self.line is not None:
            try:
                self.line.setdefault("multi_gpu", {})["watchdog"] = "multi-GPU arms did not finish within %d s; partial results" % self.seconds
                print(json.dumps(self.line, default=str), flush=True)
            except Exception:      # noqa: BLE001
                pass
        os._exit(0)

    def start(self):
        self.timer = threading.Timer(self.seconds, self._fire)
        self.timer.daemon = True
        self.timer.start()

    def cancel(self):
        if self.timer is not None:
            self.timer.cancel()


def multi_gpu_arms(args, rank, world, local_rank, dist, out):
    """Partitioned strong scaling (one grid over N strips and over N blocks), weak scaling (2M vertices per GPU), the sharded
    batch of sphere deformations, and a partitioned run checked against the CPU oracle. Sizes come from --part-nx,
    --weak-verts-per-gpu, --batch, --oracle-nx (defaults = BASELINE.json configs[3], configs[4])."""
    from mesh_deform_b200 import meshgen as G, partition as PT
    t_begin = time.perf_counter()
    out.update({"world": world, "note": "one process per GPU; partitioned arms exchange halos of p', R, the multigrid vectors and the CG's gathered "
                                        "vector and all-reduce the CG scalars every CG iteration (transport: %s)" % args.transport})
    W, K = args.warmup, args.steps

    def over_budget():      # collective: every rank must take the same decision
        return max_over_ranks(dist, local_rank, time.perf_counter() - t_begin) > args.multi_budget_s

    def arm(name, fn):
        """Run fn on every rank; record its dict (rank 0) or the error; never let one rank's exception hang the others."""
        if over_budget():
            out[name] = {"skipped": "time budget of %d s for the multi-GPU arms spent" % args.multi_budget_s}
            return None
        res, err = None, None
        try:
            res = fn()
        except Exception as exc:      # noqa: BLE001
            err = "%s: %s" % (type(exc).__name__, exc)
        if not _sync_ok(dist, local_rank, err is None):
            out[name] = {"error": err or "failed on another rank"}
            return None
        return res

    # ---- strong scaling: ONE grid (configs[4]: 4000 x 4000 = 16M vertices) over N partitions ----------------------------
    nx = args.part_nx
    strong_res = None
    P, F = G.grid_plane(nx, nx)
    idx, tgt = G.grid_constraints(nx, nx, P)
    single = {}

    def single_ref():
        if rank == 0:
            single.update(run_single_gpu_grid(P, F, idx, tgt, local_rank, W, K))
        dist.barrier()
        return True
    arm("single_gpu_reference_%dx%d" % (nx, nx), single_ref)
    for kind, owner_fn in (("strips", PT.strip_owner), ("blocks", PT.block_owner)):
        name = "partitioned_strong_%s" % kind
        res = arm(name, lambda: run_partitioned(dist, rank, world, local_rank, P, F, idx, tgt, owner_fn(P, world), W, K, args.transport))
        if res is not None and kind == "strips":
            strong_res = res
        if res is None or rank != 0:
            continue
        entry = _strip(res)
        entry["workload"] = "one %d x %d grid plane (%d vertices, BASELINE.json configs[4]) in %d %s, %d + %d ARAP iterations" % (nx, nx, nx * nx, world, kind, W, K)
        if single:
            entry["single_gpu_iterations_per_s"] = single["iterations_per_s"]
            entry["speedup_vs_single_gpu"] = res["iterations_per_s"] / single["iterations_per_s"]
            entry["strong_scaling_efficiency"] = res["iterations_per_s"] / (world * single["iterations_per_s"])
            entry["parity_vs_single_gpu"] = _parity(res["positions"], res["energy"], single["positions"], single["energy"], P)
        entry["limiter"] = ("%.0f cross-GPU operations per CG iteration at %.1f us per exchange / %.1f us per all-reduce = %.0f us of a %.0f us CG iteration"
                            % (res["exchanges_per_cg_iteration"] + res["allreduces_per_cg_iteration"], res["us_per_exchange"], res["us_per_allreduce"],
                               res["exchanges_per_cg_iteration"] * res["us_per_exchange"] + res["allreduces_per_cg_iteration"] * res["us_per_allreduce"],
                               1e3 * res["ms_per_step"] / max(1.0, res["cg_iterations_per_step"])))
        if single:
            # the two factors of the speed-up: what the partition costs the preconditioner, and how one CG iteration scales
            t_single = 1e3 * single["ms_per_step"] / max(1.0, single["cg_iterations_per_step"])
            t_part = 1e3 * res["ms_per_step"] / max(1.0, res["cg_iterations_per_step"])
            entry["limiter"] += ("; CG iterations per ARAP iteration %.2f vs %.2f unpartitioned (aggregates confined to the partition blocks); "
                                 "one CG iteration %.0f us vs %.0f us on one GPU = %.2fx on %d GPUs"
                                 % (res["cg_iterations_per_step"], single["cg_iterations_per_step"], t_part, t_single, t_single / t_part, world))
            ph, ph1 = res.get("us_per_cg_iteration_by_phase_rank0"), single.get("us_per_cg_iteration_by_phase")
            if ph and ph1:
                comm = ph["halo_exchange"] + ph["allreduce_cg_scalars"] + ph["allreduce_replicated_level_rhs"]
                entry["limiter"] += ("; event-timed phases: communication %.0f us, coarse multigrid levels %.0f us (one GPU: %.0f), fine-level kernels %.0f us (one GPU: %.0f)"
                                     % (comm, ph["multigrid_coarse_levels"], ph1["multigrid_coarse_levels"],
                                        ph["sum_of_phases_us"] - comm - ph["multigrid_coarse_levels"], ph1["sum_of_phases_us"] - ph1["multigrid_coarse_levels"]))
        out[name] = entry
    if rank == 0 and single:
        out["single_gpu_reference_%dx%d" % (nx, nx)] = _strip(single)
    del P, F
    single.clear()

    # ---- weak scaling: 2M vertices per GPU (16M at N = 8) ---------------------------------------------------------------
    per_gpu = args.weak_verts_per_gpu
    n1 = int(round(np.sqrt(per_gpu)))
    nw = int(round(np.sqrt(per_gpu * world)))
    Pw1, Fw1 = G.grid_plane(n1, n1)
    iw1, tw1 = G.grid_constraints(n1, n1, Pw1)

    def weak_ref():
        if rank == 0:
            single.update(run_single_gpu_grid(Pw1, Fw1, iw1, tw1, local_rank, W, K))
        dist.barrier()
        return True
    arm("weak_single_gpu_reference", weak_ref)
    if nw == nx and strong_res is not None:
        res_w = strong_res                               # the N-GPU point of the weak curve IS the strong-scaling grid
        reused = True
    else:
        Pw, Fw = G.grid_plane(nw, nw)
        iw, tw = G.grid_constraints(nw, nw, Pw)
        res_w = arm("partitioned_weak", lambda: run_partitioned(dist, rank, world, local_rank, Pw, Fw, iw, tw, PT.strip_owner(Pw, world), W, K,
                                                                 args.transport, gather=False))
        reused = False
        del Pw, Fw
    if rank == 0 and res_w is not None:
        entry = _strip(res_w)
        entry["workload"] = "%d x %d grid (%d vertices = %d per GPU) in %d strips%s" % (nw, nw, nw * nw, nw * nw // world, world,
                                                                                       " (same run as partitioned_strong_strips)" if reused else "")
        if single:
            entry["single_gpu_workload"] = "%d x %d grid (%d vertices) on one GPU, unpartitioned solver" % (n1, n1, n1 * n1)
            entry["single_gpu_iterations_per_s"] = single["iterations_per_s"]
            entry["single_gpu_cg_iterations_per_step"] = single["cg_iterations_per_step"]
            entry["weak_scaling_efficiency"] = entry["iterations_per_s"] / single["iterations_per_s"]
            entry["vertex_iterations_per_s"] = entry["iterations_per_s"] * nw * nw
        out["partitioned_weak"] = entry
    del Pw1, Fw1
    single.clear()

    # ---- a partitioned run against the CPU oracle (<= 1M vertices) -------------------------------------------------------
    no = args.oracle_nx
    Po, Fo = G.grid_plane(no, no)
    io, to = G.grid_constraints(no, no, Po)
    its = 4
    res_o = arm("partitioned_vs_oracle", lambda: run_partitioned(dist, rank, world, local_rank, Po, Fo, io, to, PT.block_owner(Po, world), 0, its,
                                                                   args.transport, comm_rounds=20))
    late = over_budget()
    if rank == 0 and res_o is not None and not late:
        from oracle import oracle as O
        omesh = Po.copy()
        o = O.ArapOracle(omesh, Fo, np.float64)
        for i, t in zip(io, to):
            o.setConstraint(int(i), t)
        o.deform(its)
        out["partitioned_vs_oracle"] = {"workload": "%d x %d grid (%d vertices) in %d blocks, %d ARAP iterations, against the CPU oracle (direct LDL^T solve)" % (no, no, no * no, world, its),
                                        "parity": _parity(res_o["positions"], res_o["energy"], omesh, o.energy(), Po),
                                        "cg_iterations_per_step": res_o["cg_iterations_per_step"]}
    dist.barrier()
    del Po, Fo

    # ---- configs[3]: the batch of sphere deformations, sharded contiguously, no data-path collective -----------------------
    def batch_run():
        b = batch_measure(args, rank, world, local_rank, dist)
        ref = {}
        if rank == 0:
            ref = batch_measure(args, 0, 1, local_rank, None)          # all members on one GPU: the N = 1 point
        dist.barrier()
        return b, ref
    res_b = arm("batch_spheres", batch_run)
    if rank == 0 and res_b is not None:
        b, ref = res_b
        b["single_gpu_member_iterations_per_s"] = ref["member_iterations_per_s"]
        b["strong_scaling_efficiency"] = b["member_iterations_per_s"] / (world * ref["member_iterations_per_s"])
        out["batch_spheres"] = b
    out["seconds"] = time.perf_counter() - t_begin
    return out


def batch_measure(args, rank, world, local_rank, dist):
    """K sphere deformations (configs[3]) sharded contiguously over `world` ranks; returns member-iterations/s of the whole batch."""
    from mesh_deform_b200 import capi, meshgen as G
    from mesh_deform_b200.sharding import shard_range
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    P, F = z["sphere_V"], z["sphere_F"]
    K = args.batch
    begin, end = shard_range(K, rank, world)
    handles = np.sort(np.unique(F[(F == G.SPHERE_HANDLE).any(1)]))

    def T(t=(0, 0, 0), R=np.eye(3)):
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = t
        return M
    traj = capi.TrajectorySE3()
    prev = np.eye(4)
    for step in (T(), T((0.25, 0, 0)), T((0.5, 0, 0)), T(R=G.rot_x(np.pi / 2))):
        prev = prev @ step
        traj.addKeyPose(prev)
    util = capi.DeformationUtil(P, handles, origin=traj(0.0))
    poses = traj.sample(np.arange(begin, end) / max(1, K - 1))
    bdef = capi.BatchDeformation(P, F, end - begin, np.float64, device=local_rank)
    anchor = np.array([G.SPHERE_ANCHOR], np.int32)
    bdef.setConstraints(anchor, np.repeat(P[anchor][None], end - begin, 0))
    util.updateConstraints(poses, bdef)
    t0 = time.perf_counter()
    bdef.prepare()
    prepare_ms = 1e3 * (time.perf_counter() - t0)
    bdef.iterate(args.warmup)
    barrier_and_sync(dist)
    bdef.timer_start()
    bdef.iterate(args.steps)
    ms = bdef.timer_stop()
    barrier_and_sync(dist)
    ms = max_over_ranks(dist, local_rank, ms)
    stats = bdef.solver_stats()
    res = {"workload": "%d independent deformations of sphere.obj (642 vertices) with trajectory key-frame handle poses (BASELINE.json configs[3]), "
                       "%d members per rank" % (K, end - begin),
           "member_iterations_per_s": K * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
           "deformations_of_10_iterations_per_s": K / (10 * ms / args.steps * 1e-3),
           "cg_iterations_per_step": stats["cg_iterations_total"] / max(1, stats["global_steps"]), "prepare_ms": prepare_ms,
           "preconditioner": "one member's dense inverse applied to all members (tcgen05 3xTF32 GEMM)" if stats["mg_levels"] == 1 else
                             "multigrid over the block-diagonal batch, %d levels" % stats["mg_levels"]}
    bdef.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nu", type=int, default=316, help="icosphere frequency (316 -> 998,562 vertices)")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=4, help="oracle iterations in the cpu_baseline sample")
    ap.add_argument("--cg-tol", type=float, default=0.0)
    ap.add_argument("--pos-tol", type=float, default=0.0, help="position_tolerance of the multigrid solver (0 = engine default)")
    ap.add_argument("--solver", default="auto", choices=["auto", "jacobi", "mg"])
    ap.add_argument("--workload", default="icosphere", choices=["icosphere", "batch_spheres", "partitioned_grid"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--nx", type=int, default=4000, help="partitioned_grid: the grid is nx x nx vertices (4000 -> 16M, BASELINE.json configs[4])")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "peer"], help="partitioned arms: halo exchange transport")
    ap.add_argument("--part-nx", type=int, default=4000, help="multi-GPU strong scaling: the grid is part-nx x part-nx vertices (4000 -> 16M, configs[4])")
    ap.add_argument("--weak-verts-per-gpu", type=int, default=2000000, help="multi-GPU weak scaling: vertices per GPU")
    ap.add_argument("--oracle-nx", type=int, default=700, help="multi-GPU: side of the partitioned grid checked against the CPU oracle")
    ap.add_argument("--multi-budget-s", type=int, default=540, help="multi-GPU arms: stop starting new arms after this many seconds")
    ap.add_argument("--no-multi", action="store_true", help="N > 1: only the headline replicas (round-1 behaviour)")
    ap.add_argument("--no-f32", action="store_true", help="skip the PrecisionType-float line")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    if args.workload == "batch_spheres":
        return batch_arm(args)
    if args.workload == "partitioned_grid":
        return partitioned_arm(args)

    rank, world, local_rank, dist = dist_setup(args.gpus)
    import torch  # plumbing only: device selection, barrier, max-over-ranks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use gpurun")
    torch.cuda.set_device(local_rank)
    from mesh_deform_b200 import capi

    real = np.float64 if args.precision == "f64" else np.float32
    s = np.dtype(real).itemsize
    P, F, idx, tgt = build_workload(args.nu)
    V = P.shape[0]
    peak_gbs, peak_src, peaks = load_peaks()

    opts = {"device": local_rank, "solver": {"auto": 0, "jacobi": 1, "mg": 2}[args.solver]}
    if args.cg_tol > 0:
        opts["cg_tolerance"] = args.cg_tol
    if args.pos_tol > 0:
        opts["position_tolerance"] = args.pos_tol

    # ---- device-resident arm ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)          # clocks + throttle reasons for the whole measured part of the run
    sampler.start()
    pinned = capi.PinnedArray((V, 3), real)
    mesh = pinned.array
    mesh[...] = P.astype(real)
    arap = capi.AsRigidAsPossibleDeformation(mesh, F, real, **opts)
    arap.setConstraints(idx, tgt)
    t0 = time.perf_counter()
    rc = arap.prepare()
    arap.synchronize()
    prepare_ms = 1e3 * (time.perf_counter() - t0)
    assert rc == capi.ARAP_OK
    prepare_host_setup_ms = arap.solver_stats()["setup_host_ms"]
    prepare_device_setup_ms = arap.solver_stats()["setup_device_ms"]
    rp, ci, _ = arap.cotanWeights()
    nnz = int(ci.size)
    _, n_free = arap.freeIdxMap()

    # warm-up = the COLD START: ARAP iterations 1..W right after the demo-sized handle move (Rz(30 deg) + 0.3 lift), one call
    # each so that the CG work per iteration is on record (the timed window below is the nearly converged steady state)
    cold = {"cg_iterations": [], "ms": []}
    for _ in range(args.warmup):
        before = arap.solver_stats()["cg_iterations_total"]
        arap.timer_start()
        arap.iterate(1)
        cold["ms"].append(arap.timer_stop())
        cold["cg_iterations"].append(int(arap.solver_stats()["cg_iterations_total"] - before))
    arap.synchronize()

    arap.profile_reset()
    st0 = arap.solver_stats()
    barrier_and_sync(dist)
    arap.timer_start()
    arap.iterate(args.steps)
    ms = arap.timer_stop()
    barrier_and_sync(dist)
    ms = max_over_ranks(dist, local_rank, ms)
    prof_plain = arap.profile()
    stats = arap.solver_stats()
    launches = sum(v["launches"] for v in prof_plain.values())
    window_cg = (stats["cg_iterations_total"] - st0["cg_iterations_total"]) / max(1, stats["global_steps"] - st0["global_steps"])

    # parity snapshot at the END of the timed window: W + K iterations after the cold start, the state the number was measured on
    parity = None
    parity_iters = args.warmup + args.steps
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        parity = {"iterations": parity_iters, "gpu_positions": arap.positions(np.float64), "gpu_energy": arap.energy()}

    # ---- per-kernel pass: the same K steps with every launch bracketed by CUDA events -----------
    arap.profile_enable(True)
    arap.profile_reset()
    arap.timer_start()
    arap.iterate(args.steps)
    ms_profiled = arap.timer_stop()
    prof = arap.profile()
    arap.profile_enable(False)

    # ---- end-to-end arm: public call with a HOST mesh, write-back inside the timed region ---------
    arap.deform(1)
    barrier_and_sync(dist)
    t0 = time.perf_counter()
    arap.timer_start()
    for _ in range(args.steps):
        assert arap.deform(1)                      # arap_deform(h, host_mesh, 1): iterate + D2H write-back
    e2e_ms_dev = arap.timer_stop()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier_and_sync(dist)
    e2e_ms = max_over_ranks(dist, local_rank, max(e2e_ms, e2e_ms_dev))

    # ---- the same arm pipelined: arap_deform_async / arap_deform_wait with two page-locked buffers, so the write-back of frame k
    # overlaps the iteration of frame k + 1. Every step still launches one frame and receives one complete frame on the host.
    pinned2 = capi.PinnedArray((V, 3), real)
    bufs = [mesh, pinned2.array]
    arap.deform_async(bufs[0], 1)                  # prime the pipeline (untimed)
    barrier_and_sync(dist)
    t0 = time.perf_counter()
    touched = 0.0
    for k in range(args.steps):
        arap.deform_async(bufs[(k + 1) % 2], 1)
        assert arap.deform_wait()                  # frame k is in bufs[k % 2]
        touched += float(bufs[k % 2][0, 0])
    e2e_pipe_ms = 1e3 * (time.perf_counter() - t0)
    assert arap.deform_wait()                      # drain the extra frame
    barrier_and_sync(dist)
    e2e_pipe_ms = max_over_ranks(dist, local_rank, e2e_pipe_ms)

    # ---- frame protocol of the reference demos: setConstraint + deform(5), dirty every frame (copies on BOTH sides)
    frame_ms = []
    for f in range(3):
        t0 = time.perf_counter()
        arap.setConstraints(idx[-10:], tgt[-10:] + 1e-3 * (f + 1))         # a handle moves -> full dirty rebuild (arap.h:84,102-120)
        assert arap.deform(5)
        frame_ms.append(1e3 * (time.perf_counter() - t0))
    frame_ms = float(np.median(frame_ms))
    clocks = sampler.stop()

    # ---- PrecisionType float on the same workload (informational; single GPU only) -----------------------------------------
    f32 = None
    if world == 1 and args.precision == "f64" and not args.no_f32:
        m32 = P.astype(np.float32)
        a32 = capi.AsRigidAsPossibleDeformation(m32, F, np.float32, **opts)
        a32.setConstraints(idx, tgt)
        a32.prepare()
        a32.iterate(args.warmup)
        a32.synchronize()
        a32.timer_start()
        a32.iterate(args.steps)
        ms32 = a32.timer_stop()
        st32 = a32.solver_stats()
        f32 = {"value": args.steps / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / args.steps,
               "cg_iterations_per_arap_iteration": st32["cg_iterations_total"] / max(1, st32["global_steps"]),
               "positions": a32.positions(np.float64), "energy": a32.energy(), "iterations": parity_iters}
        a32.close()

    if dist is not None and not args.no_multi:
        arap.close()
        del pinned, mesh, bufs, pinned2
    line = None
    if rank == 0:
        value = world * args.steps / (ms * 1e-3)
        e2e_value = world * args.steps / (e2e_ms * 1e-3)
        e2e_pipe_value = world * args.steps / (e2e_pipe_ms * 1e-3)
        ab = algorithmic_bytes(V, n_free, nnz, s)
        total_kernel_ms = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {}
        for name, v in prof.items():
            if v["launches"] == 0:
                continue
            avg_ms = v["ms"] / v["launches"]
            entry = {"launches_per_step": v["launches"] / args.steps, "avg_us": 1e3 * avg_ms, "share": v["ms"] / total_kernel_ms}
            if name in ab and avg_ms > 0:
                entry["algorithmic_bytes"] = ab[name]
                entry["achieved_gbs"] = ab[name] / (avg_ms * 1e-3) / 1e9
                entry["frac_of_peak"] = entry["achieved_gbs"] / peak_gbs
            kernels[name] = entry
        # DRAM traffic per launch of each kernel from the committed `ncu --set full` capture of this same command
        ncu_file = next((f for f in ("r02_ncu_summary.json", "r01_h_ncu_summary.json") if os.path.exists(os.path.join(ROOT, "profiles", f))), None)
        ncu = json.load(open(os.path.join(ROOT, "profiles", ncu_file))) if ncu_file and args.nu == 316 and args.precision == "f64" else {}
        ncu_alias = {"cg_spmv": "cg_spmv_z", "cg_update_mg": "cg_fused_update"}         # profile id -> kernel name in the capture
        for name, entry in kernels.items():
            rec = ncu.get(name) or ncu.get(ncu_alias.get(name, ""))
            if rec and "traffic_bytes" in rec:
                entry["ncu_dram_traffic_bytes"] = rec["traffic_bytes"]
        # The roofline object is pinned to the kernel BASELINE.json's metric names (the local step); the whole step's aggregate and
        # every other kernel sit beside it. (Round 1 picked "the kernel with the largest share", which flipped between three ~11 % kernels.)
        pinned_kernel = "local_step" if "local_step" in kernels and "achieved_gbs" in kernels["local_step"] else \
            max((k for k in kernels if "achieved_gbs" in kernels[k]), key=lambda k: kernels[k]["share"])
        pk = kernels[pinned_kernel]
        step_bytes = sum(kernels[k]["algorithmic_bytes"] * kernels[k]["launches_per_step"] for k in kernels if "algorithmic_bytes" in kernels[k])
        roofline = {"kernel": pinned_kernel, "bound": "hbm", "achieved": pk["achieved_gbs"], "peak": peak_gbs,
                    "unit": "GB/s", "frac": pk["frac_of_peak"],
                    "traffic": pk.get("ncu_dram_traffic_bytes"), "algorithmic_bytes": pk["algorithmic_bytes"],
                    "traffic_source": ("profiles/%s (ncu --set full, dram__bytes_read+write per launch)" % ncu_file) if ncu else None,
                    "peak_source": peak_src, "share_of_step": pk["share"], "avg_launch_us": pk["avg_us"],
                    "step_aggregate": {"algorithmic_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                                       "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak_gbs,
                                       "note": "algorithmic bytes of every fine-level kernel of one ARAP iteration / un-profiled time per iteration "
                                               "(the coarse multigrid levels count as time but not as bytes)"}}
        local = kernels.get("local_step", {})

        cpu_baseline = None
        parity_out = None
        if parity is not None:
            o, omesh, t_prep = run_cpu_oracle(P, F, idx, tgt, 0, real)
            o.reset_timers()
            t0 = time.perf_counter()
            o.deform(parity_iters)
            dt = time.perf_counter() - t0
            tm = o.timers()
            diag = float(np.linalg.norm(P.max(0) - P.min(0)))
            cpu_e = o.energy()
            parity_out = {"iterations": parity_iters,
                          "max_dp_over_bbox_diag": float(np.abs(parity["gpu_positions"] - omesh.astype(np.float64)).max() / diag),
                          "rel_energy_diff": abs(parity["gpu_energy"] - cpu_e) / cpu_e, "tolerance": {"dp": 1e-5, "dE": 1e-6},
                          "note": "engine vs the CPU oracle after the cold start + %d warm-up + %d timed iterations: the state at the end of the timed window" % (args.warmup, args.steps)}
            if f32 is not None:
                f32["max_dp_over_bbox_diag_vs_fp64_oracle"] = float(np.abs(f32.pop("positions") - omesh.astype(np.float64)).max() / diag)
                f32["rel_energy_diff_vs_fp64_oracle"] = abs(f32.pop("energy") - cpu_e) / cpu_e
            cpu_baseline = {"value": parity_iters / dt, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"full workload (V={V}), {parity_iters} ARAP iterations after prepare; prepare {t_prep:.1f} s "
                                      f"(LDL^T factor, {o.factor_nnz()} nnz) excluded; per-iteration s: local {tm['local'] / parity_iters:.3f} "
                                      f"rhs {tm['rhs'] / parity_iters:.3f} solve {tm['solve'] / parity_iters:.3f}; host cores available: {os.cpu_count()}"}
        if f32 is not None:
            f32.pop("positions", None)
            f32.pop("energy", None)

        working_set_mb = (V * (3 * 4 * s + 8) + nnz * (4 + s) + V * 4 + V * 4 * 24) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"icosphere nu={args.nu} V={V} 5% anchors + 1% handles (BASELINE.json configs[2])",
                       "vertices": int(V), "faces": int(F.shape[0]), "nnz": nnz, "n_free": int(n_free),
                       "sharding": "one independent deformation per GPU, no collective (the partitioned and batched workloads are in `multi_gpu`)" if world > 1 else "single GPU",
                       "solver": ("warm-started single-reduction CG, smoothed-aggregation multigrid V(1,1) preconditioner, %d levels, operator complexity %.2f, CG loop %s"
                                  % (stats["mg_levels"], stats["mg_operator_complexity"],
                                     "on the device (one CUDA graph per ARAP iteration)" if stats["cg_graph"] == 2 else "driven by the host (one CUDA graph per CG iteration)"))
                       if stats["mg_levels"] else "warm-started Jacobi-PCG (matrix-free CSR SpMV, 3 RHS)", "stopping_rule": stopping_rule(args),
                       "one_ring_kernels": ("neighbourhood staged through shared memory in tiles of 256 rows (largest tile halo %d)" % stats["tile_max_halo"])
                       if stats["tile_max_halo"] > 0 else "gathers straight from global memory",
                       "vertex_order": "renumbered internally in Morton patches" if stats["renumbered"] else "the caller's order",
                       "l2": f"inputs larger than L2: ~{working_set_mb:.0f} MB touched per step vs 126 MB L2, no explicit flush"},
            "e2e": {"value": e2e_pipe_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(V * 3 * s),
                    "call": "arap_deform_async(h, pinned_host_mesh[k % 2], 1) + arap_deform_wait(h) on a prepared handle: every step enqueues 1 iteration "
                            "+ the write-back of p' and receives one complete frame in host memory; the write-back of frame k overlaps the iteration of "
                            "frame k + 1 (wall clock over the K steps, max over ranks)",
                    "ms_per_step_pipelined": e2e_pipe_ms / args.steps,
                    "synchronous": {"value": e2e_value, "call": "arap_deform(h, pinned_host_mesh, 1): 1 iteration, then the write-back, then return",
                                    "ms_per_step": e2e_ms / args.steps},
                    "h2d_note": "deform() reads the mesh only when a constraint changed (reference arap.h:102-107), so a steady-state step has no "
                                "host input; the per-frame protocol WITH the upload is `frame`",
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "local_step": {"achieved_gbs": local.get("achieved_gbs"), "frac_of_measured_peak": local.get("frac_of_peak"),
                           "frac_of_nominal_8TBs": (local.get("achieved_gbs") or 0) / 8000.0, "avg_us": local.get("avg_us")},
            "kernels": kernels,
            "kernels_note": "avg_us from a second pass of the same K steps with every launch bracketed by CUDA events (graphs off): ~1.5x slower than "
                            "the timed pass overall and ~5 us too long per small kernel; shares agree with the ncu launch list in profiles/",
            "cg": {"iterations_per_arap_iteration": window_cg, "kernel_launches_per_cg_iteration": stats["launches_per_cg_iteration"],
                   "last_relative_residual": stats["last_relative_residual"], "converged": bool(stats["last_converged"])},
            "cold_start": {"what": "ARAP iterations 1..%d right after the handle move (the warm-up), one arap_iterate(1) each" % args.warmup,
                           "cg_iterations": cold["cg_iterations"], "ms": cold["ms"]},
            "prepare_ms": prepare_ms, "prepare_host_setup_ms": prepare_host_setup_ms, "prepare_device_setup_ms": prepare_device_setup_ms,
            "frame": {"protocol": "setConstraint(handles) + deform(5) incl. the dirty rebuild: H2D rest pose, weights/CSR, 5 iterations, D2H (reference demo loop)",
                      "ms": frame_ms, "h2d_bytes": int(V * 3 * s + 10 * (4 + 3 * 8)), "d2h_bytes": int(V * 3 * s),
                      "iterations_per_s": 5.0 / (frame_ms * 1e-3)},
            "profiled_pass_ms_per_step": ms_profiled / args.steps,
            "f32": f32,
            "cpu_baseline": cpu_baseline,
            "parity": parity_out,
        }
    if dist is not None and not args.no_multi:
        # The multi-GPU workloads run AFTER the headline line is assembled, under a watchdog: should a rank ever hang in a
        # collective, rank 0 still prints the line (with whatever arms finished) and every rank leaves before the driver's limit.
        multi = {}
        if line is not None:
            line["multi_gpu"] = multi
        watchdog = Watchdog(rank, line, args.multi_budget_s + 150)
        watchdog.start()
        multi_gpu_arms(args, rank, world, local_rank, dist, multi)
        watchdog.cancel()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def stopping_rule(args):
    """The engine's rule for ending a global solve (include/arap_b200.h: cg_tolerance, position_tolerance)."""
    if args.solver == "jacobi":
        return {"relative_residual": args.cg_tol if args.cg_tol > 0 else 1e-9}
    rule = {}
    if args.cg_tol > 0:
        rule["relative_residual"] = args.cg_tol
    if args.pos_tol > 0 or not args.cg_tol > 0:
        rule["estimated_position_error_over_bbox_diagonal"] = args.pos_tol if args.pos_tol > 0 else 1e-8
    return rule


if __name__ == "__main__":
    sys.exit(main())
