"""GPU parity tests (pytest -m gpu): the CUDA engine, called through the C ABI (ctypes), against
the CPU oracle on the same inputs. Tolerances are BASELINE.json's: max vertex displacement
<= 1e-5 x bounding-box diagonal, ARAP energy <= 1e-6 relative; CSR indices bit-exact."""
import numpy as np
import pytest

from conftest import bbox_diag
from oracle import oracle as O
from mesh_deform_b200 import meshgen as G
from mesh_deform_b200.capi import AsRigidAsPossibleDeformation as ARAP
from mesh_deform_b200 import capi

pytestmark = pytest.mark.gpu

POS_TOL = 1e-5      # x bbox diagonal (north_star)
E_TOL = 1e-6        # relative (north_star)


def constrain(a, idx, tgt):
    for i, t in zip(idx, tgt):
        a.setConstraint(int(i), t)


def test_cotan_weights_reference_test():
    """reference tests/test_cotan.cpp:20-54 against the engine (float precision like the reference)."""
    P = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    arap = ARAP(P, F)                       # PrecisionType defaults to the mesh scalar (float)
    assert arap.deform(0) is True
    assert arap.dirty                       # unconstrained: early return, stays dirty (arap.h:113-114)
    rp, ci, w = arap.cotanWeights()
    dense = np.zeros((4, 4), np.float32)
    for r in range(4):
        dense[r, ci[rp[r]:rp[r + 1]]] = w[rp[r]:rp[r + 1]]
    expected = np.array([[0, .5, 0, .5], [.5, 0, .5, 0], [0, .5, 0, .5], [.5, 0, .5, 0]], np.float32)
    assert np.linalg.norm(dense - expected) <= 1e-4 * min(np.linalg.norm(dense), np.linalg.norm(expected))
    assert rp.tolist() == [0, 3, 5, 8, 10] and ci.tolist() == [1, 2, 3, 0, 2, 0, 1, 3, 0, 2]
    assert np.array_equal(P, np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32))   # no write-back


@pytest.mark.parametrize("prec", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["bar", "sphere", "plane"])
def test_csr_bit_exact(name, prec, meshes, golden):
    P, F = meshes[name]
    mesh = P.astype(prec)
    a = ARAP(mesh, F, prec)
    constrain(a, golden[name + "_idx"], golden[name + "_tgt"])
    assert a.deform(0)
    rp, ci, w = a.cotanWeights()
    o = O.ArapOracle(P.astype(prec), F, prec)
    o.deform(0)
    orp, oci, ow = o.cotanWeights()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)       # indices: bit-exact
    assert np.array_equal(w, ow)                                      # weights: bit-exact too (same rounding sequence)
    if prec == np.float64:
        assert np.array_equal(rp, golden[name + "_rowptr"]) and np.array_equal(ci, golden[name + "_colidx"])
        assert np.array_equal(w, golden[name + "_w"])
    fm, nfree = a.freeIdxMap()
    o2 = O.ArapOracle(P.astype(prec), F, prec)
    constrain(o2, golden[name + "_idx"], golden[name + "_tgt"])
    o2.deform(0)
    assert nfree == o2.nFree and np.array_equal(fm, o2.freeIdxMap())


SOLVERS = {"jacobi": capi.SOLVER_PCG_JACOBI, "mg": capi.SOLVER_PCG_MG}


@pytest.mark.parametrize("solver", ["jacobi", "mg"])
@pytest.mark.parametrize("name", ["bar", "sphere", "plane"])
def test_deform_matches_golden_fp64(name, solver, meshes, golden):
    P, F = meshes[name]
    iters = int(golden[name + "_iters"])
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64, solver=SOLVERS[solver])
    a.setConstraints(golden[name + "_idx"], golden[name + "_tgt"])
    energies = []
    for _ in range(iters):
        assert a.deform(1)
        energies.append(a.energy())
    diag = bbox_diag(P)
    assert np.abs(mesh - golden[name + "_positions"]).max() <= POS_TOL * diag
    assert np.allclose(energies, golden[name + "_energies"], rtol=E_TOL)
    # much tighter in practice: report
    print(name, solver, "max dp / diag", np.abs(mesh - golden[name + "_positions"]).max() / diag,
          "rel dE", abs(energies[-1] - golden[name + "_energies"][-1]) / golden[name + "_energies"][-1], a.solver_stats())
    assert np.abs(a.rotations() - golden[name + "_rotations"]).max() < 1e-6


def test_first_local_step_rotations_match_oracle(meshes, golden):
    """R_i after one iteration, vertex by vertex (local step arap.h:354-384)."""
    P, F = meshes["bar"]
    mesh, omesh = P.copy(), P.copy()
    a, o = ARAP(mesh, F, np.float64), O.ArapOracle(omesh, F, np.float64)
    constrain(a, golden["bar_idx"], golden["bar_tgt"])
    constrain(o, golden["bar_idx"], golden["bar_tgt"])
    a.deform(1)
    o.deform(1)
    R, Ro = a.rotations(), o.rotations()
    assert np.abs(R - Ro).max() < 1e-7          # Newton accepts at |omega| < 1e-4: remaining error ~1e-8
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-12
    assert np.abs(mesh - omesh).max() <= 1e-8 * bbox_diag(P)


@pytest.mark.parametrize("name", ["bar", "sphere"])
def test_deform_float_precision(name, meshes, golden):
    """PrecisionType = float (the sphere demo's default, deform_sphere.cpp:55): within the float
    reference's own error of the fp64 oracle."""
    P, F = meshes[name]
    iters = int(golden[name + "_iters"])
    mesh = P.astype(np.float32)
    a = ARAP(mesh, F, np.float32)
    a.setConstraints(golden[name + "_idx"], golden[name + "_tgt"])
    assert a.deform(iters)
    omesh = P.astype(np.float32)
    o = O.ArapOracle(omesh, F, np.float32)
    constrain(o, golden[name + "_idx"], golden[name + "_tgt"])
    o.deform(iters)
    diag = bbox_diag(P)
    err_gpu = np.abs(mesh - golden[name + "_positions"]).max() / diag
    err_ref = np.abs(omesh - golden[name + "_positions"]).max() / diag
    print(name, "float: gpu err", err_gpu, "float-oracle err", err_ref)
    assert err_gpu <= max(2 * err_ref, 2e-5)


def test_dirty_protocol_and_warm_continuation(meshes):
    """SURVEY.md section 8b semantics (1)-(6)."""
    P, F = meshes["sphere"]
    mesh = P.astype(np.float32)                  # mesh scalar float, PrecisionType double (deform_bar.cpp:38)
    a = ARAP(mesh, F, np.float64)
    a.setConstraint(37, P[37])
    a.setConstraint(32, P[32] + [0, 0, 0.5])
    assert a.dirty and a.deform(0) and not a.dirty
    assert np.allclose(mesh[32], (P[32] + [0, 0, 0.5]).astype(np.float32))     # (3) deform(0) snaps handles
    assert np.array_equal(mesh[0], P[0].astype(np.float32))
    m2 = P.astype(np.float32)
    b = ARAP(m2, F, np.float64)
    b.setConstraint(37, P[37])
    b.setConstraint(32, P[32] + [0, 0, 0.5])
    a.deform(2); a.deform(3)                                                  # (4) warm continuation
    b.deform(5)
    assert np.abs(mesh - m2).max() <= 1e-6
    # (1) a constraint change re-reads the rest pose from the deformed mesh
    omesh = mesh.copy()
    o = O.ArapOracle(omesh, F, np.float64)
    o.setConstraint(37, P[37]); o.setConstraint(32, P[32] + [0, 0, 0.6])
    a.setConstraint(32, P[32] + [0, 0, 0.6])
    assert a.dirty
    a.deform(3); o.deform(3)
    assert np.abs(mesh - omesh).max() <= POS_TOL * bbox_diag(P)
    assert abs(a.energy() - o.energy()) <= 1e-5 * o.energy() + 1e-12


def test_rigid_motion_of_all_constraints_gives_rigid_result(meshes):
    P, F = meshes["sphere"]
    R = G.rot_z(0.7) @ G.rot_x(-0.4)
    t = np.array([0.3, -0.2, 0.5])
    idx = np.arange(0, len(P), 7)
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64, cg_tolerance=1e-11)       # a limit property: solve (almost) exactly
    a.setConstraints(idx, P[idx] @ R.T + t)
    assert a.deform(30)
    # the flip-flop iteration converges linearly: the oracle is at 1.016e-4 after 30 iterations
    assert abs(np.abs(mesh - (P @ R.T + t)).max() - 1.0159958e-4) < 1e-8
    assert a.deform(70)
    assert np.abs(mesh - (P @ R.T + t)).max() < 1e-7
    assert a.energy() < 1e-12


def test_energy_monotone_and_csr_properties():
    P, F = G.icosphere(24)                      # 5762 vertices
    idx, tgt = G.cap_constraints(P)
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64)
    a.setConstraints(idx, tgt)
    prev = None
    for _ in range(6):
        assert a.deform(1)
        e = a.energy()
        # energy after each global step, measured with the rotations of that iteration's local step, is non-increasing
        if prev is not None:
            assert e <= prev * (1 + 1e-9)
        prev = e
    rp, ci, w = a.cotanWeights()
    import scipy.sparse as sp
    W = sp.csr_matrix((w, ci, rp), shape=(len(P),) * 2)
    assert (abs(W - W.T)).max() == 0 and W.diagonal().max() == 0
    assert all(np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0) for r in range(0, len(P), 37))


@pytest.mark.parametrize("solver", ["jacobi", "mg"])
@pytest.mark.parametrize("nu,iters", [(64, 4)])
def test_midsize_icosphere_parity(nu, iters, solver):
    """Config 3's construction at a size the oracle finishes in seconds (40,962 vertices)."""
    P, F = G.icosphere(nu)
    idx, tgt = G.cap_constraints(P)
    mesh, omesh = P.copy(), P.copy()
    a, o = ARAP(mesh, F, np.float64, solver=SOLVERS[solver]), O.ArapOracle(omesh, F, np.float64)
    a.setConstraints(idx, tgt)
    constrain(o, idx, tgt)
    assert a.deform(iters) and o.deform(iters)
    diag = bbox_diag(P)
    err = np.abs(mesh - omesh).max() / diag
    de = abs(a.energy() - o.energy()) / o.energy()
    print("ico", nu, solver, "err/diag", err, "rel dE", de, a.solver_stats())
    assert err <= POS_TOL and de <= E_TOL


def test_edge_cases():
    # empty constraint set on a real mesh: true, stays dirty, mesh untouched, weights available
    P, F = G.icosphere(4)
    mesh = P.copy()
    a = ARAP(mesh, F, np.float64)
    assert a.deform(5) and a.dirty and np.array_equal(mesh, P)
    assert a.cotanWeights()[1].size == 6 * (len(P) - 2)
    # every vertex constrained: nothing free, result = targets
    b = ARAP(mesh, F, np.float64)
    b.setConstraints(np.arange(len(P)), P * 1.5)
    assert b.deform(3) and np.allclose(mesh, P * 1.5)
    # degenerate (zero-area) face and an isolated vertex: clamps of arap.h:199-218 keep everything finite
    P2 = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0], [5, 5, 5]], np.float64)
    F2 = np.array([[0, 1, 2], [0, 1, 3]], np.int32)
    m2 = P2.copy()
    c = ARAP(m2, F2, np.float64)
    c.setConstraint(0, P2[0]); c.setConstraint(3, P2[3] + [0, 0, 0.2])
    o2 = P2.copy()
    o = O.ArapOracle(o2, F2, np.float64)
    o.setConstraint(0, P2[0]); o.setConstraint(3, P2[3] + [0, 0, 0.2])
    assert np.array_equal(c.prepare() >= 0, True)
    rp, ci, w = c.cotanWeights()
    o.deform(0)
    orp, oci, ow = o.cotanWeights()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(w, ow)
    assert np.isfinite(w).all()
    # invalid arguments are reported, not UB
    with pytest.raises(capi.ArapError):
        a.setConstraint(10 ** 6, [0, 0, 0])


def delaunay_patch(n, seed):
    """An irregular open surface: Delaunay triangulation of random points in the unit square, lifted to a bumpy height field.
    Valences range from 3 to ~12 and there is a boundary -- unlike the generated spheres and grids."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    uv = rng.random((n, 2))
    tri = Delaunay(uv).simplices.astype(np.int32)
    P = np.column_stack([uv[:, 0], 0.15 * np.sin(5 * uv[:, 0]) * np.cos(4 * uv[:, 1]), uv[:, 1]])
    return np.ascontiguousarray(P), np.ascontiguousarray(tri)


@pytest.mark.parametrize("solver", ["mg", "jacobi"])
def test_irregular_delaunay_mesh_parity(solver):
    """Badly conditioned on purpose (sliver triangles: cotan weights from 5e-11 to 3e3). The multigrid solver's default
    stopping rule is a position-error estimate precisely because a residual tolerance tuned on regular meshes left this
    mesh 4e-5 x bbox diagonal away from the direct solve (profiles/r01_h_stopping_rule.txt)."""
    P, F = delaunay_patch(6000, 21)
    val = np.bincount(F.ravel(), minlength=len(P))
    assert val.min() <= 4 and val.max() >= 10                      # really irregular
    rng = np.random.default_rng(4)
    idx = rng.choice(len(P), 60, replace=False).astype(np.int32)
    tgt = P[idx] + 0.05 * rng.standard_normal((60, 3))
    mesh, omesh = P.copy(), P.copy()
    a, o = ARAP(mesh, F, np.float64, solver=SOLVERS[solver]), O.ArapOracle(omesh, F, np.float64)
    a.setConstraints(idx, tgt)
    constrain(o, idx, tgt)
    assert a.deform(4) and o.deform(4)
    rp, ci, w = a.cotanWeights()
    orp, oci, ow = o.cotanWeights()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(w, ow)
    assert w.max() / w.min() > 1e12
    err = np.abs(mesh - omesh).max() / bbox_diag(P)
    de = abs(a.energy() - o.energy()) / o.energy()
    print("delaunay", solver, "err/diag", err, "rel dE", de)
    assert err <= POS_TOL and de <= E_TOL
    # PrecisionType float on the same mesh: the weights are still bit-exact, but a float LDL^T on this system (the float
    # oracle) is itself ~2e-2 x diag away from the double solution -- the engine, which solves in fp64 whatever the
    # PrecisionType, must be at least as close to the double solution as the float reference is.
    fmesh, fomesh = P.astype(np.float32), P.astype(np.float32)
    fa, fo = ARAP(fmesh, F, np.float32, solver=SOLVERS[solver]), O.ArapOracle(fomesh, F, np.float32)
    fa.setConstraints(idx, tgt)
    constrain(fo, idx, tgt)
    assert fa.deform(4) and fo.deform(4)
    frp, fci, fw = fa.cotanWeights()
    forp, foci, fow = fo.cotanWeights()
    assert np.array_equal(frp, forp) and np.array_equal(fci, foci) and np.array_equal(fw, fow)
    err_engine = np.abs(fmesh.astype(np.float64) - omesh).max() / bbox_diag(P)
    err_reference = np.abs(fomesh.astype(np.float64) - omesh).max() / bbox_diag(P)
    print("delaunay float: engine vs double", err_engine, "float oracle vs double", err_reference)
    assert np.isfinite(fmesh).all() and err_engine <= max(err_reference, 1e-3) * 1.05


def test_nonmanifold_duplicate_faces_and_two_components():
    """setFromTriplets semantics on awkward input (arap.h:220-238): an edge shared by three faces, a face listed twice and
    a face with reversed winding all just add their half-cotans; two components, each with its own constraints."""
    P = np.array([[0, 0, 0], [1, 0, 0], [0.5, 1, 0], [0.5, -1, 0.2], [0.5, 0.3, 1.0], [1.5, 1, 0.1],     # fan around edge 0-1
                  [5, 0, 0], [6, 0, 0], [5, 1, 0], [6, 1, 0.3]], np.float64)                                  # second component
    F = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4], [1, 5, 2], [1, 5, 2], [2, 5, 1], [6, 7, 8], [7, 9, 8]], np.int32)
    idx = np.array([0, 3, 6, 9], np.int32)
    tgt = P[idx] + np.array([[0, 0, 0], [0, 0, 0.3], [0, 0, 0], [0.2, 0, 0.2]])
    for prec in (np.float64, np.float32):
        mesh, omesh = P.astype(prec), P.astype(prec)
        a, o = ARAP(mesh, F, prec), O.ArapOracle(omesh, F, prec)
        a.setConstraints(idx, tgt)
        constrain(o, idx, tgt)
        assert a.deform(6) and o.deform(6)
        rp, ci, w = a.cotanWeights()
        orp, oci, ow = o.cotanWeights()
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(w, ow)
        tol = POS_TOL if prec == np.float64 else 300 * POS_TOL
        assert np.abs(mesh.astype(np.float64) - omesh.astype(np.float64)).max() <= tol * bbox_diag(P)


def test_full_size_equivariance_and_energy_descent():
    """Properties that need no oracle, at the headline size (998,562 vertices): (1) the ARAP energy never increases from
    one iteration to the next; (2) moving the whole problem -- rest pose and targets -- by a rigid motion moves the result by
    the same rigid motion (cotan weights, covariances and the linear system are all invariant)."""
    P, F = G.icosphere(316)
    idx, tgt = G.cap_constraints(P)
    a = ARAP(P.copy(), F, np.float64, cg_tolerance=1e-9)
    a.setConstraints(idx, tgt)
    energies = []
    for _ in range(6):
        assert a.deform(1)
        energies.append(a.energy())
    assert all(e1 <= e0 * (1 + 1e-9) for e0, e1 in zip(energies, energies[1:])), energies
    Rm = G.rot_x(0.7) @ G.rot_z(-1.1)
    t = np.array([0.3, -2.0, 1.5])
    b = ARAP(np.ascontiguousarray(P @ Rm.T + t), F, np.float64, cg_tolerance=1e-9)
    b.setConstraints(idx, tgt @ Rm.T + t)
    assert b.deform(6)
    err = np.abs(b.mesh - (a.mesh @ Rm.T + t)).max() / bbox_diag(P)
    print("full-size equivariance err/diag", err, "energies", energies[0], energies[-1], "dE", abs(b.energy() - a.energy()) / a.energy())
    assert err <= 1e-7 and abs(b.energy() - a.energy()) <= 1e-7 * a.energy()


def test_midsize_grid_parity_mg():
    """Config 5's construction (plane.obj topology up-scaled, 2+2 constraint columns) at 200 x 200:
    every quad diagonal carries the 1e-10 clamp weight and iteration-1 covariances are rank 2."""
    n = 200
    P, F = G.grid_plane(n, n)
    idx, tgt = G.grid_constraints(n, n, P)
    mesh, omesh = P.copy(), P.copy()
    a, o = ARAP(mesh, F, np.float64), O.ArapOracle(omesh, F, np.float64)
    a.setConstraints(idx, tgt)
    constrain(o, idx, tgt)
    assert a.deform(3) and o.deform(3)
    err = np.abs(mesh - omesh).max() / bbox_diag(P)
    de = abs(a.energy() - o.energy()) / o.energy()
    print("grid", n, "err/diag", err, "rel dE", de, a.solver_stats())
    assert err <= POS_TOL and de <= E_TOL


def sphere_trajectory():
    """BASELINE.json configs[3] / SURVEY.md section 8d config 4: poses from the SE(3) trajectory through the key poses of
    reference examples/deform_trajectory.cpp:66-69 with the translations scaled by 0.25."""
    def T(t=(0, 0, 0), R=np.eye(3)):
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = t
        return M
    traj = O.TrajectorySE3Oracle()
    prev = np.eye(4)
    traj.addKeyPose(prev)
    prev = prev @ T((0.25, 0, 0))
    traj.addKeyPose(prev)
    prev = prev @ T((0.5, 0, 0))
    traj.addKeyPose(prev)
    prev = prev @ T(R=G.rot_x(np.pi / 2))
    traj.addKeyPose(prev)
    return traj


def test_batch_of_sphere_deformations_matches_oracle(meshes):
    P, F = meshes["sphere"]
    K = 6
    handles = np.sort(np.unique(F[(F == G.SPHERE_HANDLE).any(1)]))          # v32 and its one-ring
    idx = np.concatenate([[G.SPHERE_ANCHOR], handles]).astype(np.int32)
    traj = sphere_trajectory()
    origin = traj(0.0)
    targets = np.zeros((K, idx.size, 3))
    for k in range(K):
        pose = traj(k / (K - 1))
        targets[k, 0] = P[G.SPHERE_ANCHOR]
        targets[k, 1:] = O.handle_targets(origin, pose, P[handles])
    b = capi.BatchDeformation(P, F, K, np.float64)
    b.setConstraints(idx, targets)
    assert b.prepare() == capi.ARAP_OK
    b.iterate(10)
    pos = b.positions()
    diag = bbox_diag(P)
    for k in range(K):
        mesh = P.copy()
        o = O.ArapOracle(mesh, F, np.float64)
        constrain(o, idx, targets[k])
        assert o.deform(10)
        assert np.abs(pos[k] - mesh).max() <= POS_TOL * diag, k


def test_batch_with_member_specific_constraints_falls_back_to_the_general_solver(meshes):
    """The shared-operator shortcut of batches (one member's dense inverse for all) is only valid when every member has the
    same constrained SET. Constraints added through the raw handle can break that; the engine must notice and still be right."""
    import ctypes as C
    P, F = meshes["sphere"]
    K, V = 3, P.shape[0]
    idx = np.array([G.SPHERE_ANCHOR, G.SPHERE_HANDLE], np.int32)
    tg = np.stack([P[idx] + np.array([[0, 0, 0], [0, 0, 0.1 * (m + 1)]]) for m in range(K)])
    extra_vertex = 100
    extra_target = P[extra_vertex] + np.array([0.05, 0.0, 0.0])
    b = capi.BatchDeformation(P, F, K, np.float64)
    b.setConstraints(idx, tg)
    one = np.array([1 * V + extra_vertex], np.int32)                        # member 1 only, super-mesh numbering
    xyz = np.ascontiguousarray(extra_target[None], np.float64)
    b._check(capi.lib().arap_set_constraints(b._h, 1, one.ctypes.data_as(C.c_void_p), xyz.ctypes.data_as(C.c_void_p), 8))
    assert b.prepare() == capi.ARAP_OK
    b.iterate(6)
    pos = b.positions()
    for m in range(K):
        mesh = P.copy()
        o = O.ArapOracle(mesh, F, np.float64)
        constrain(o, idx, tg[m])
        if m == 1:
            o.setConstraint(extra_vertex, extra_target)
        assert o.deform(6)
        assert np.abs(pos[m] - mesh).max() <= POS_TOL * bbox_diag(P), m


def test_rigid_constraint_front_end_matches_per_vertex_constraints(meshes):
    """DeformationUtil::updateConstraints in one call (arap_set_rigid_constraints / arap_batch_set_rigid_constraints,
    targets computed on the device) against setConstraint with the oracle's targets (deformation_util.h:48-57)."""
    P, F = meshes["sphere"]
    K = 5
    handles = np.sort(np.unique(F[(F == G.SPHERE_HANDLE).any(1)]))
    otraj = sphere_trajectory()
    ptraj = capi.TrajectorySE3()
    for pose in otraj._poses:                              # same key poses into the product's trajectory
        ptraj.addKeyPose(pose)
    origin = ptraj(0.0)
    assert np.abs(origin - otraj(0.0)).max() < 1e-12
    us = np.arange(K) / (K - 1)
    poses = ptraj.sample(us)
    util = capi.DeformationUtil(P, handles, origin)
    anchor = np.array([G.SPHERE_ANCHOR], np.int32)
    # single mesh
    for k in (1, K - 1):
        a = capi.AsRigidAsPossibleDeformation(P.copy(), F, np.float64)
        a.setConstraints(anchor, P[anchor])
        util.updateConstraints(poses[k], a)
        assert a.deform(6)
        mesh = P.copy()
        o = O.ArapOracle(mesh, F, np.float64)
        o.setConstraint(int(anchor[0]), P[anchor[0]])
        constrain(o, handles, O.handle_targets(otraj(0.0), otraj(float(us[k])), P[handles]))
        assert o.deform(6)
        assert np.abs(a.mesh - mesh).max() <= POS_TOL * bbox_diag(P)
    # batch: one pose per member
    b = capi.BatchDeformation(P, F, K, np.float64)
    b.setConstraints(anchor, np.repeat(P[anchor][None], K, 0))
    util.updateConstraints(poses, b)
    assert b.prepare() == capi.ARAP_OK
    b.iterate(6)
    pos = b.positions()
    for k in range(K):
        mesh = P.copy()
        o = O.ArapOracle(mesh, F, np.float64)
        o.setConstraint(int(anchor[0]), P[anchor[0]])
        constrain(o, handles, O.handle_targets(otraj(0.0), otraj(float(us[k])), P[handles]))
        assert o.deform(6)
        assert np.abs(pos[k] - mesh).max() <= POS_TOL * bbox_diag(P), k
    with pytest.raises(capi.ArapError):
        b.setRigidConstraints(np.array([P.shape[0]], np.int32), P[:1], poses)


@pytest.mark.parametrize("transport", ["host", "peer"])
@pytest.mark.parametrize("solver", ["mg", "jacobi"])
@pytest.mark.parametrize("world", [1, 2, 4])
def test_partitioned_mesh_in_process_matches_oracle(world, solver, transport):
    """BASELINE.json configs[4] in small: the plane.obj-topology grid with 2+2 constraint columns, partitioned into
    `world` strips that run concurrently on ONE GPU (in-process transport: same solver code path as NCCL, copies
    instead of NVLink). Result must match the unpartitioned CPU oracle."""
    from mesh_deform_b200 import partition as PT
    nx, nz, iters = 64, 48, 4
    P, F = G.grid_plane(nx, nz)
    idx, tgt = G.grid_constraints(nx, nz, P)
    owner = PT.strip_owner(P, world)
    key = 1000 + 10 * world + (0 if solver == "mg" else 1) + (100 if transport == "peer" else 0)
    kind = capi.TRANSPORT_PEER_IN_PROCESS if transport == "peer" else capi.TRANSPORT_IN_PROCESS     # peer: direct stores + flags
    parts = [capi.PartitionedDeformation(P, F, owner, r, world, kind, key, np.float64, solver=SOLVERS[solver])
             for r in range(world)]

    def work(p):
        def run():
            p.setConstraints(idx, tgt)
            assert p.prepare() == capi.ARAP_OK
            p.iterate(iters)
        return run
    capi.run_partitions_in_process([work(p) for p in parts])
    pos = np.zeros_like(P)
    for p in parts:
        gid, xyz = p.owned_positions()
        pos[gid] = xyz
    energy = sum(p.local_energy() for p in parts)
    omesh = P.copy()
    o = O.ArapOracle(omesh, F, np.float64)
    constrain(o, idx, tgt)
    assert o.deform(iters)
    err = np.abs(pos - omesh).max() / bbox_diag(P)
    de = abs(energy - o.energy()) / o.energy()
    print("partitioned", world, solver, transport, "err/diag", err, "rel dE", de, [p.solver_stats()["last_cg_iterations"] for p in parts])
    assert err <= POS_TOL and de <= E_TOL
    its = {p.solver_stats()["cg_iterations_total"] for p in parts}
    assert len(its) == 1                      # every rank took exactly the same control path


def _run_partitioned(P, F, idx, tgt, world, key, iters, transport=None, **kw):
    from mesh_deform_b200 import partition as PT
    owner = PT.strip_owner(P, world)
    kind = capi.TRANSPORT_IN_PROCESS if transport is None else transport
    parts = [capi.PartitionedDeformation(P, F, owner, r, world, kind, key, np.float64, **kw) for r in range(world)]

    def work(p):
        def run():
            p.setConstraints(idx, tgt)
            assert p.prepare() == capi.ARAP_OK
            p.iterate(iters)
        return run
    capi.run_partitions_in_process([work(p) for p in parts])
    pos = np.zeros_like(P)
    for p in parts:
        gid, xyz = p.owned_positions()
        pos[gid] = xyz
    return pos, parts[0].solver_stats()


@pytest.mark.parametrize("replicate_rows", ["0", "2000", None])
@pytest.mark.parametrize("transport", ["host", "peer"])
def test_partitioned_mesh_in_quadrants_matches_oracle(transport, replicate_rows, monkeypatch):
    """A 2 x 2 block partition instead of strips: two ranks have three neighbours (one of them only across the corner), the halo plans of
    the multigrid levels differ from rank to rank. Same parity bar against the unpartitioned oracle.
    replicate_rows (ARAP_MG_REPLICATE_ROWS): "0" keeps every level but the coarsest partitioned (four exchanges per level), "2000"
    replicates the small ones, None is the default (everything below the fine level is replicated at this size)."""
    if replicate_rows is None:
        monkeypatch.delenv("ARAP_MG_REPLICATE_ROWS", raising=False)
    else:
        monkeypatch.setenv("ARAP_MG_REPLICATE_ROWS", replicate_rows)
    nx, nz, iters = 200, 160, 4            # 32k vertices: three multigrid levels, so a partitioned intermediate level
    P, F = G.grid_plane(nx, nz)
    idx, tgt = G.grid_constraints(nx, nz, P)
    owner = ((P[:, 0] > np.median(P[:, 0])).astype(np.int32) + 2 * (P[:, 2] > np.median(P[:, 2])).astype(np.int32)).astype(np.int32)
    kind = capi.TRANSPORT_PEER_IN_PROCESS if transport == "peer" else capi.TRANSPORT_IN_PROCESS
    key = (3001 if transport == "peer" else 3002) + 10 * (["0", "2000", None].index(replicate_rows))
    # (position_tolerance 1e-8: with the default 3e-8 this gently bent, fine grid lands at 7.8e-7 relative energy -- inside the
    #  1e-6 bar, but this test is about the partition plumbing, not about the margin of the stopping rule)
    parts = [capi.PartitionedDeformation(P, F, owner, r, 4, kind, key, np.float64, position_tolerance=1e-8) for r in range(4)]
    assert sorted(len(p.part.neighbor_rank) for p in parts) == [2, 2, 3, 3]      # the quad diagonals join only one pair of opposite quadrants

    def work(p):
        def run():
            p.setConstraints(idx, tgt)
            assert p.prepare() == capi.ARAP_OK
            p.iterate(iters)
        return run
    capi.run_partitions_in_process([work(p) for p in parts])
    pos = np.zeros_like(P)
    for p in parts:
        gid, xyz = p.owned_positions()
        pos[gid] = xyz
    omesh = P.copy()
    o = O.ArapOracle(omesh, F, np.float64)
    constrain(o, idx, tgt)
    assert o.deform(iters)
    err = np.abs(pos - omesh).max() / bbox_diag(P)
    de = abs(sum(p.local_energy() for p in parts) - o.energy()) / o.energy()
    st = parts[0].solver_stats()
    print("quadrants", transport, replicate_rows, "err/diag", err, "rel dE", de, "levels", st["mg_levels"], "global", st["mg_global"],
          "exchanges / all-reduces per CG iteration", st["comm_exchanges_per_cg_iteration"], st["comm_allreduces_per_cg_iteration"])
    assert st["comm_allreduces_per_cg_iteration"] == 2                      # the replicated levels' right-hand side + the CG scalars
    assert st["comm_exchanges_per_cg_iteration"] == (4 + 4 * (st["mg_levels"] - 2) if replicate_rows == "0" else 4 if replicate_rows is None else st["comm_exchanges_per_cg_iteration"])
    assert err <= POS_TOL and de <= E_TOL and st["mg_global"] == 1 and st["mg_levels"] >= 3


def test_partitioned_global_multigrid_keeps_the_iteration_count():
    """The point of the global hierarchy (arap_partition_set_global_mesh): partitioning must not cost CG iterations.
    Block-Jacobi across ranks (each rank preconditioning only its own block) is the baseline it replaces."""
    nx, nz, iters = 192, 160, 3
    P, F = G.grid_plane(nx, nz)
    idx, tgt = G.grid_constraints(nx, nz, P)
    single = capi.AsRigidAsPossibleDeformation(P.copy(), F, np.float64)
    single.setConstraints(idx, tgt)
    assert single.deform(iters)
    its_single = single.solver_stats()["cg_iterations_total"]
    pos_g, st_g = _run_partitioned(P, F, idx, tgt, 4, 2001, iters)
    pos_p, st_p = _run_partitioned(P, F, idx, tgt, 4, 2003, iters, transport=capi.TRANSPORT_PEER_IN_PROCESS)
    assert np.array_equal(pos_p, pos_g) and st_p["cg_iterations_total"] == st_g["cg_iterations_total"]     # same arithmetic, other wires
    pos_b, st_b = _run_partitioned(P, F, idx, tgt, 4, 2002, iters, global_multigrid=False)
    print("cg iterations: single", its_single, "partitioned global", st_g["cg_iterations_total"], "block-Jacobi", st_b["cg_iterations_total"],
          "levels", st_g["mg_levels"])
    diag = bbox_diag(P)
    assert np.abs(pos_g - single.mesh).max() <= POS_TOL * diag
    assert np.abs(pos_b - single.mesh).max() <= POS_TOL * diag
    assert st_g["cg_iterations_total"] <= its_single + 2 * iters          # at most a couple more per global step
    assert st_b["cg_iterations_total"] > st_g["cg_iterations_total"]


def test_hierarchy_reuse_across_dirty_cycles_keeps_parity():
    """The reference's demos change a constraint every frame (full dirty rebuild, arap.h:84,102-120). The engine keeps the
    multigrid hierarchy while the constrained SET is unchanged; the result must still match the oracle frame by frame."""
    P, F = G.icosphere(20)
    idx, tgt = G.cap_constraints(P)
    mesh, omesh = P.copy(), P.copy()
    a, o = ARAP(mesh, F, np.float64), O.ArapOracle(omesh, F, np.float64)
    a.setConstraints(idx, tgt)
    constrain(o, idx, tgt)
    setups = []
    for frame in range(4):
        move = np.array([0.0, 0.0, 0.02 * (frame + 1)])
        a.setConstraints(idx[-5:], tgt[-5:] + move)
        for i, t in zip(idx[-5:], tgt[-5:] + move):
            o.setConstraint(int(i), t)
        assert a.deform(3) and o.deform(3)
        st = a.solver_stats()
        setups.append(st["setup_host_ms"] + st["setup_device_ms"])       # whichever side built the hierarchy
        assert np.abs(mesh - omesh).max() <= POS_TOL * bbox_diag(P)
    assert setups[0] > 0 and all(s == 0 for s in setups[1:])      # one fresh setup, then reuse
    # a NEW constrained vertex changes the mask -> fresh hierarchy
    new = int(np.setdiff1d(np.arange(len(P)), idx)[0])
    a.setConstraint(new, mesh[new].copy())
    o.setConstraint(new, omesh[new].copy())
    assert a.deform(2) and o.deform(2)
    assert a.solver_stats()["setup_host_ms"] + a.solver_stats()["setup_device_ms"] > 0
    assert np.abs(mesh - omesh).max() <= POS_TOL * bbox_diag(P)


@pytest.mark.parametrize("agg_key", ["rim", "sweep"])
def test_partitioned_strips_with_either_aggregation(agg_key, monkeypatch):
    """The global hierarchy of a partitioned solver with its aggregates confined to the partition blocks, built with each of the
    two root elections of the device setup (ARAP_MG_AGG_KEY: rim growth is the default beyond 4 partitions, the wavefront up to 4):
    same parity bar against the unpartitioned oracle, global hierarchy in use."""
    monkeypatch.setenv("ARAP_MG_AGG_KEY", agg_key)
    nx, nz, iters = 200, 160, 4
    P, F = G.grid_plane(nx, nz)
    idx, tgt = G.grid_constraints(nx, nz, P)
    pos, st = _run_partitioned(P, F, idx, tgt, 2, 3100 + (agg_key == "rim"), iters, position_tolerance=1e-8)
    omesh = P.copy()
    o = O.ArapOracle(omesh, F, np.float64)
    constrain(o, idx, tgt)
    assert o.deform(iters)
    err = np.abs(pos - omesh).max() / bbox_diag(P)
    print("strips", agg_key, "err/diag", err, "levels", st["mg_levels"], "CG iterations", st["cg_iterations_total"])
    assert err <= POS_TOL and st["mg_global"] == 1 and st["mg_levels"] >= 3 and st["setup_device_ms"] > 0
