"""GPU tests (pytest -m gpu) added in round 2: the right-hand side against the oracle's `_b`, the overwrite semantics of
setConstraint, the dirty guard of arap_iterate, the not-converged status, rows of very high valence in the CSR build, the
device-side CG loop (step graph) against the host-driven loop, and parity deep into a run (25 iterations) in both precisions."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, bbox_diag
from oracle import oracle as O
from mesh_deform_b200 import capi, meshgen as G
from mesh_deform_b200.capi import AsRigidAsPossibleDeformation as ARAP

pytestmark = pytest.mark.gpu

POS_TOL = 1e-5      # x bbox diagonal (north_star)
E_TOL = 1e-6        # relative (north_star)


def oracle_for(P, F, idx, tgt, prec=np.float64):
    mesh = P.astype(prec)
    o = O.ArapOracle(mesh, F, prec)
    for i, t in zip(idx, tgt):
        o.setConstraint(int(i), t)
    return o, mesh


@pytest.mark.parametrize("solver", [capi.SOLVER_PCG_MG, capi.SOLVER_PCG_JACOBI])
@pytest.mark.parametrize("name", ["bar", "sphere", "ico"])
def test_right_hand_side_matches_oracle(name, solver, meshes, golden):
    """b = bFixed + sum_j w_ij/2 (R_i + R_j)(p_i - p_j) (arap.h:393-414), row by row, after 1 and after 3 iterations.
    `ico` is large enough (5,762 vertices) for the multigrid hierarchy; bar and sphere take the dense-inverse path."""
    if name == "ico":
        P, F = G.icosphere(24)
        idx, tgt = G.cap_constraints(P)
    else:
        P, F = meshes[name]
        idx, tgt = golden[name + "_idx"], golden[name + "_tgt"]
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64, solver=solver)
    a.setConstraints(idx, tgt)
    o, _ = oracle_for(P, F, idx, tgt)
    for its in (1, 2):
        assert a.deform(its) and o.deform(its)
        b = a.rhs()
        ob = np.asarray(o.b(), np.float64)
        ob = ob.reshape(3, -1).T if ob.shape[0] == 3 and ob.ndim == 2 else ob.reshape(-1, 3)
        assert b.shape == ob.shape
        scale = np.abs(ob).max()
        print(name, "rhs max diff / max|b|", np.abs(b - ob).max() / scale)
        assert np.abs(b - ob).max() <= 1e-7 * scale      # R_i carries ~1e-8 (Newton acceptance), everything else is fp64


def test_set_constraint_twice_last_value_wins(meshes):
    """reference arap.h:83 overwrites the map entry; one batched call naming a vertex twice must do the same."""
    P, F = meshes["sphere"]
    first, last = P[32] + [0, 0, 0.9], P[32] + [0, 0, 0.5]
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64)
    a.setConstraints(np.array([37, 32, 32], np.int32), np.stack([P[37], first, last]))
    assert a.deform(3)
    o, omesh = oracle_for(P, F, [37, 32, 32], [P[37], first, last])
    assert o.deform(3)
    assert np.allclose(mesh[32], last, atol=1e-12)
    assert np.abs(mesh - omesh).max() <= POS_TOL * bbox_diag(P)
    # many duplicates in one call (a handle set dragged twice before the next deform): always the last occurrence
    idx = np.tile(np.array([32, 33, 34], np.int32), 50)
    tgt = np.repeat(np.arange(50.0)[:, None, None], 3, 1) * 1e-3 + P[[32, 33, 34]][None]
    a.setConstraints(idx, tgt.reshape(-1, 3))
    assert a.deform(0)
    assert np.allclose(mesh[[32, 33, 34]], tgt[-1], atol=1e-12)
    # the rigid front end de-duplicates handles the same way
    a.setRigidConstraints(np.array([32, 32], np.int32), np.stack([P[32], P[32] + [1, 0, 0]]), np.eye(4))
    assert a.deform(0)
    assert np.allclose(mesh[32], P[32] + [1, 0, 0], atol=1e-12)


def test_iterate_on_a_dirty_handle_is_refused(meshes):
    P, F = meshes["sphere"]
    a = ARAP(P.copy(), F, np.float64)
    a.setConstraints(np.array([37, 32], np.int32), np.stack([P[37], P[32] + [0, 0, 0.5]]))
    assert a.prepare() == capi.ARAP_OK
    a.iterate(1)
    a.setConstraint(32, P[32] + [0, 0, 0.6])
    with pytest.raises(capi.ArapError) as e:
        a.iterate(1)
    assert e.value.code == capi.ARAP_ERR_INVALID and "dirty" in str(e.value)
    assert a.prepare() == capi.ARAP_OK
    a.iterate(1)


@pytest.mark.parametrize("nu", [6, 24])
def test_unconverged_global_solve_is_reported(nu):
    """max_cg_iterations too small for the stopping rule: arap_iterate returns ARAP_NOT_CONVERGED, deform() false
    (the reference's `false` for an unusable system, arap.h:116-117), the iterations still ran."""
    P, F = G.icosphere(nu)
    idx, tgt = G.cap_constraints(P)
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64, solver=capi.SOLVER_PCG_JACOBI, max_cg_iterations=2)
    a.setConstraints(idx, tgt)
    assert a.prepare() == capi.ARAP_OK
    assert a.iterate(2) == capi.ARAP_NOT_CONVERGED
    st = a.solver_stats()
    assert st["last_converged"] == 0 and st["global_steps"] == 2 and st["cg_iterations_total"] == 4
    assert a.deform(1) is False
    b = ARAP(P.copy(), F, np.float64, max_cg_iterations=1)           # multigrid solver, device-side loop
    b.setConstraints(idx, tgt)
    assert b.prepare() == capi.ARAP_OK
    rc = b.iterate(1)
    if b.solver_stats()["mg_levels"] > 1:
        assert rc == capi.ARAP_NOT_CONVERGED


def test_high_valence_fan_csr_bit_exact():
    """A triangle fan whose centre has 3,000 neighbours (6,000 raw triplets in one row): the row goes through the CTA-wide
    sort (row_sort_long_kernel) and must come out exactly as the reference's setFromTriplets orders and sums it."""
    n = 3000
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    P = np.concatenate([[[0, 0, 0]], np.stack([np.cos(ang) * (1 + 0.3 * np.sin(5 * ang)), np.sin(ang), 0.1 * np.cos(3 * ang)], 1)])
    F = np.stack([np.zeros(n, np.int32), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n], 1).astype(np.int32)
    for prec in (np.float64, np.float32):
        a = ARAP(P.astype(prec), F, prec)
        assert a.deform(0)
        rp, ci, w = a.cotanWeights()
        o = O.ArapOracle(P.astype(prec), F, prec)
        o.deform(0)
        orp, oci, ow = o.cotanWeights()
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(w, ow)
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64)
    idx = np.array([1, 2, 3, n // 2], np.int32)
    tgt = P[idx] + [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0.2]]
    a.setConstraints(idx, tgt)
    assert a.deform(3)
    o, omesh = oracle_for(P, F, idx, tgt)
    assert o.deform(3)
    assert np.abs(mesh - omesh).max() <= POS_TOL * bbox_diag(P)


def test_device_side_loop_equals_host_driven_loop():
    """The step graph (whole ARAP iteration in one CUDA graph, CG loop on the device) and the host-driven loop run the same
    kernels in the same order: identical iteration counts and positions. The host-driven handle lives in a subprocess
    because the switch is read when the library is loaded."""
    P, F = G.icosphere(24)
    idx, tgt = G.cap_constraints(P)
    a = ARAP(P.copy(), F, np.float64)
    a.setConstraints(idx, tgt)
    assert a.deform(6)
    st = a.solver_stats()
    assert st["cg_graph"] == 2, "the step graph was not built"
    assert st["global_steps"] == 6 and st["last_converged"] == 1
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from mesh_deform_b200 import meshgen as G\n"
            "from mesh_deform_b200.capi import AsRigidAsPossibleDeformation as ARAP\n"
            "P, F = G.icosphere(24); idx, tgt = G.cap_constraints(P)\n"
            "m = P.copy(); a = ARAP(m, F, np.float64); a.setConstraints(idx, tgt); assert a.deform(6)\n"
            "st = a.solver_stats(); assert st['cg_graph'] == 1, st\n"
            "np.save(sys.argv[1], m); print(st['cg_iterations_total'])\n" % ROOT)
    out = os.path.join(ROOT, "gpurun_out", "host_loop_positions.npy")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ, ARAP_STEP_GRAPH="0")
    res = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    assert int(res.stdout.strip().splitlines()[-1]) == st["cg_iterations_total"]
    assert np.array_equal(np.load(out), a.mesh)


@pytest.mark.parametrize("prec", [np.float64, np.float32])
def test_parity_deep_into_a_run(prec):
    """25 ARAP iterations from a cold start after a demo-sized handle move (163k vertices): the drift against the oracle at
    the iteration count of bench.py's timed window, in both precisions."""
    P, F = G.icosphere(128)
    idx, tgt = G.cap_constraints(P)
    mesh = P.astype(prec)
    a = ARAP(mesh, F, prec)
    a.setConstraints(idx, tgt)
    assert a.deform(25)
    e = a.energy()
    o, omesh = oracle_for(P, F, idx, tgt, prec)
    assert o.deform(25)
    diag = bbox_diag(P)
    dp = np.abs(mesh.astype(np.float64) - omesh.astype(np.float64)).max() / diag
    de = abs(e - o.energy()) / o.energy()
    print("25 iterations,", np.dtype(prec).name, "max dp / diag", dp, "rel dE", de, a.solver_stats())
    if prec == np.float64:
        assert dp <= POS_TOL and de <= E_TOL
    else:
        # PrecisionType float: the float reference itself is only this close to the fp64 one; compare with its own error
        o64, m64 = oracle_for(P, F, idx, tgt, np.float64)
        assert o64.deform(25)
        ref_err = np.abs(omesh.astype(np.float64) - m64).max() / diag
        gpu_err = np.abs(mesh.astype(np.float64) - m64).max() / diag
        print("float: oracle32 vs oracle64", ref_err, "engine32 vs oracle64", gpu_err)
        assert gpu_err <= max(2 * ref_err, 2e-5)


@pytest.mark.parametrize("mesh", ["ico", "grid", "delaunay"])
def test_device_built_hierarchy_matches_host_built(mesh, monkeypatch):
    """The multigrid hierarchy built by the CUDA setup kernels (parallel MIS-2 aggregation, row-wise sparse products) against
    the one built on the host (greedy aggregation in Morton order): both are only preconditioners of the same exact system, so
    the results agree to the solver tolerance and the CG iteration counts stay close; both match the oracle."""
    if mesh == "ico":
        P, F = G.icosphere(64)
        idx, tgt = G.cap_constraints(P)
    elif mesh == "grid":
        P, F = G.grid_plane(220, 180)
        idx, tgt = G.grid_constraints(220, 180, P)
    else:
        rng = np.random.default_rng(5)
        from scipy.spatial import Delaunay
        pts = rng.random((20000, 2))
        F = Delaunay(pts).simplices.astype(np.int32)
        P = np.stack([pts[:, 0], 0.05 * np.sin(6 * pts[:, 0]) * np.cos(5 * pts[:, 1]), pts[:, 1]], 1)
        order = np.argsort(pts[:, 0])
        idx = np.concatenate([order[:200], order[-100:]]).astype(np.int32)
        tgt = P[idx].copy()
        tgt[200:] += [0.0, 0.15, 0.0]
    its = 4
    results = {}
    for side in ("device", "host"):
        monkeypatch.setenv("ARAP_MG_DEVICE_SETUP", "1" if side == "device" else "0")
        m = P.copy()
        a = ARAP(m, F, np.float64)
        a.setConstraints(idx, tgt)
        assert a.deform(its)
        st = a.solver_stats()
        assert (st["setup_device_ms"] > 0) == (side == "device") and (st["setup_host_ms"] > 0) == (side == "host"), st
        results[side] = (m, a.energy(), st)
    o, omesh = oracle_for(P, F, idx, tgt)
    assert o.deform(its)
    diag = bbox_diag(P)
    for side, (m, e, st) in results.items():
        dp, de = np.abs(m - omesh).max() / diag, abs(e - o.energy()) / o.energy()
        print(mesh, side, "levels", st["mg_levels"], "complexity %.3f" % st["mg_operator_complexity"], "CG iterations", st["cg_iterations_total"],
              "setup ms", st["setup_device_ms"] + st["setup_host_ms"], "dp/diag", dp, "rel dE", de)
        assert dp <= POS_TOL and de <= E_TOL
    dev, host = results["device"][2], results["host"][2]
    assert dev["cg_iterations_total"] <= 1.6 * host["cg_iterations_total"] + 4      # a different aggregation, not a worse preconditioner
    assert dev["mg_operator_complexity"] <= 1.9          # sweeps over spatial cells on the Delaunay patch: 1.73 (host greedy 1.52)


def test_viewer_interop_render_buffers(meshes):
    """reference examples/osg_viewer.cpp:45-72: after deform() the viewer recomputes the normals (OpenMesh update_normals) and
    refills its float vertex / normal arrays. arap_get_render_buffers produces both on the device, into host memory or straight
    into a device buffer of the caller (what a mapped OpenGL VBO is)."""
    import torch
    P, F = meshes["sphere"]
    mesh = P.astype(np.float32)
    a = ARAP(mesh, F, np.float64)
    a.setConstraint(37, P[37])
    a.setConstraint(32, P[32] + [0, 0, 0.5])
    assert a.deform(5)
    pos, nrm = a.render_buffers()
    assert np.array_equal(pos, mesh)                                    # the same float cast as the write-back (arap.h:133-135)
    p = a.positions(np.float64)
    fn = np.cross(p[F[:, 1]] - p[F[:, 0]], p[F[:, 2]] - p[F[:, 0]])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    vn = np.zeros_like(p)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    vn /= np.linalg.norm(vn, axis=1, keepdims=True)
    assert np.abs(nrm - vn).max() < 2e-6
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1).max() < 1e-6
    # the same into caller-owned device memory: no host round trip of the geometry
    d_pos = torch.zeros((P.shape[0], 3), dtype=torch.float32, device="cuda")
    d_nrm = torch.zeros((P.shape[0], 3), dtype=torch.float32, device="cuda")
    a.render_buffers(device_pointers=(d_pos.data_ptr(), d_nrm.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(d_pos.cpu().numpy(), pos) and np.array_equal(d_nrm.cpu().numpy(), nrm)
    pos_only, none = a.render_buffers(normals=False)
    assert none is None and np.array_equal(pos_only, pos)


@pytest.mark.parametrize("prec", [np.float64, np.float32])
def test_pipelined_deform_matches_synchronous(prec):
    """arap_deform_async / arap_deform_wait: the write-back of frame k runs on a second stream while frame k + 1 iterates. Every
    frame that arrives in its page-locked buffer is bit-identical to what the synchronous arap_deform writes for the same frame;
    a third frame in flight is refused; a handle move in between (dirty block) drains the pipeline and stays correct."""
    P, F = G.icosphere(24)
    idx, tgt = G.cap_constraints(P)
    frames = 6
    ref_mesh = P.astype(prec)
    a = ARAP(ref_mesh, F, prec)
    a.setConstraints(idx, tgt)
    ref = []
    for k in range(frames):
        if k == 4:
            a.setConstraints(idx[-5:], tgt[-5:] + 0.01)
        assert a.deform(1)
        ref.append(ref_mesh.copy())

    b = ARAP(P.astype(prec), F, prec)
    b.setConstraints(idx, tgt)
    pinned = [capi.PinnedArray(P.shape, prec), capi.PinnedArray(P.shape, prec)]
    bufs = [pinned[0].array, pinned[1].array]
    bufs[0][:] = P.astype(prec)                       # the first frame runs the dirty block: the buffer holds the rest pose
    got = []
    b.deform_async(bufs[0], 1)
    for k in range(1, frames):
        if k == 4:
            b.setConstraints(idx[-5:], tgt[-5:] + 0.01)
            assert b.deform_wait()                    # the move reads the previous frame's result: drain first, then reuse its buffer
            got.append(bufs[(k - 1) % 2].copy())
            bufs[k % 2][:] = bufs[(k - 1) % 2]
            b.deform_async(bufs[k % 2], 1)
            continue
        b.deform_async(bufs[k % 2], 1)
        if k == 2:
            with pytest.raises(capi.ArapError):
                b.deform_async(bufs[0], 1)            # two frames in flight already
        assert b.deform_wait()
        if len(got) < k:
            got.append(bufs[(k - 1) % 2].copy())
    assert b.deform_wait()
    got.append(bufs[(frames - 1) % 2].copy())
    assert len(got) == frames
    for k in range(frames):
        assert np.array_equal(got[k], ref[k]), (k, np.abs(got[k] - ref[k]).max())


def test_wavefront_round_budget_hands_over_to_rim_growth(monkeypatch):
    """The wavefront aggregation of quad-like meshes needs as many rounds as the longest dependency chain of the mesh; beyond its
    budget (ARAP_MG_WAVEFRONT_ROUNDS, by default a few times sqrt(rows)) the rest of the level is aggregated by rim growth from what
    is decided so far. Forced here after 64 rounds on a 220 x 180 plane: still a device-built hierarchy, same parity."""
    monkeypatch.setenv("ARAP_MG_WAVEFRONT_ROUNDS", "64")
    P, F = G.grid_plane(220, 180)
    idx, tgt = G.grid_constraints(220, 180, P)
    m = P.copy()
    a = ARAP(m, F, np.float64)
    a.setConstraints(idx, tgt)
    assert a.deform(4)
    st = a.solver_stats()
    o, omesh = oracle_for(P, F, idx, tgt)
    assert o.deform(4)
    dp, de = np.abs(m - omesh).max() / bbox_diag(P), abs(a.energy() - o.energy()) / o.energy()
    print("budgeted wavefront: levels", st["mg_levels"], "CG iterations", st["cg_iterations_total"], "dp/diag", dp, "rel dE", de)
    assert st["setup_device_ms"] > 0 and st["mg_levels"] >= 3
    assert dp <= POS_TOL and de <= E_TOL
