"""CPU tests: the oracle against every pin the reference's own tests hold for the path, against the
committed golden vectors, and against the independent numpy/scipy restatement."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle.numpy_ref import NumpyArap
from mesh_deform_b200 import meshgen as G
from conftest import bbox_diag


def unit_square():
    # reference tests/test_cotan.cpp:26-32
    P = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    return P, F


@pytest.mark.parametrize("prec", [np.float32, np.float64])
def test_cotan_weights_pin(prec):
    """reference tests/test_cotan.cpp:20-54 restated: deform(0) without constraints, then the
    4x4 weight matrix must be isApprox(expected, 1e-4)."""
    P, F = unit_square()
    arap = O.ArapOracle(P.copy(), F, prec)
    assert arap.deform(0) is True
    assert arap.dirty                      # arap.h:113-114 returns before _dirty = false
    rp, ci, w = arap.cotanWeights()
    dense = sp.csr_matrix((w, ci, rp), shape=(4, 4)).toarray().astype(np.float32)
    expected = np.array([[0, .5, 0, .5], [.5, 0, .5, 0], [0, .5, 0, .5], [.5, 0, .5, 0]], np.float32)
    assert dense.shape == (4, 4)
    # Eigen isApprox: |a-b|_F <= prec * min(|a|_F, |b|_F)
    assert np.linalg.norm(dense - expected) <= 1e-4 * min(np.linalg.norm(dense), np.linalg.norm(expected))
    # structure: every mesh edge present (the diagonal 0-2 carries the 1e-10 clamp), sorted columns
    assert rp.tolist() == [0, 3, 5, 8, 10]
    assert ci.tolist() == [1, 2, 3, 0, 2, 0, 1, 3, 0, 2]


def test_trajectory_endpoints_pin(trajectory_golden):
    """reference tests/test_trajectory.cpp:17-39: path(0) ~ pose0, path(1) ~ pose3 at 1e-3."""
    poses = trajectory_golden["key_poses"]
    path = O.TrajectorySE3Oracle()
    for p in poses:
        path.addKeyPose(p)
    for u, want in ((0.0, poses[0]), (1.0, poses[3])):
        got = path(u)
        assert np.linalg.norm(got - want) <= 1e-3 * min(np.linalg.norm(got), np.linalg.norm(want))
    # the oracle reproduces its own committed samples (guards against silent oracle drift)
    for u, want in zip(trajectory_golden["u"], trajectory_golden["samples"]):
        assert np.abs(path(float(u)) - want).max() < 1e-12


def test_se3_log_exp_roundtrip():
    rng = np.random.default_rng(7)
    for _ in range(200):
        xi = rng.standard_normal(6) * np.array([2, 2, 2, 1, 1, 1])
        if np.linalg.norm(xi[3:]) > 3.0:                 # log returns the minimal angle (< pi)
            xi[3:] *= 3.0 / np.linalg.norm(xi[3:])
        T = O.se3_exp(xi)
        assert np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() < 1e-12
        assert np.abs(O.se3_log(T) - xi).max() < 1e-9
    assert np.abs(O.se3_exp(np.zeros(6)) - np.eye(4)).max() == 0


def test_spline_interpolates_key_poses():
    """Interpolate(): the spline passes through pose k at its chord-length parameter."""
    rng = np.random.default_rng(3)
    poses = [O.se3_exp(rng.standard_normal(6) * 0.5) for _ in range(6)]
    path = O.TrajectorySE3Oracle()
    for p in poses:
        path.addKeyPose(p)
    path(0.0)
    for u, p in zip(path._params, poses):
        assert np.abs(path(float(u)) - p).max() < 1e-9


def _run(cls, P, F, idx, tgt, iters, **kw):
    mesh = P.copy()
    a = cls(mesh, F, **kw)
    for i, t in zip(idx, tgt):
        a.setConstraint(int(i), t)
    energies = []
    for _ in range(iters):
        assert a.deform(1)
        energies.append(a.energy())
    return mesh, energies, a


@pytest.mark.parametrize("name", ["bar", "sphere", "plane"])
def test_oracle_matches_golden_and_numpy(name, meshes, golden):
    P, F = meshes[name]
    idx, tgt, iters = golden[name + "_idx"], golden[name + "_tgt"], int(golden[name + "_iters"])
    mesh, energies, a = _run(O.ArapOracle, P, F, idx, tgt, iters, precision=np.float64)
    rp, ci, w = a.cotanWeights()
    assert np.array_equal(rp, golden[name + "_rowptr"]) and np.array_equal(ci, golden[name + "_colidx"])
    assert np.array_equal(w, golden[name + "_w"])
    assert a.nFree == int(golden[name + "_nfree"])
    assert np.abs(mesh - golden[name + "_positions"]).max() < 1e-12
    assert np.allclose(energies, golden[name + "_energies"], rtol=1e-12)
    # independent restatement
    mesh_n, energies_n, _ = _run(NumpyArap, P, F, idx, tgt, iters)
    assert np.abs(mesh - mesh_n).max() < 1e-10 * bbox_diag(P)
    assert np.allclose(energies, energies_n, rtol=1e-9)


def test_survey_scratch_values(meshes, golden):
    """SURVEY.md section 6/8c scratch numbers: bar E_10 = 1.05263564259, sum|p'| = 1657.181486918721;
    sphere converges (|dE| <= 1e-8 E) at iteration 35 with E = 0.136396706."""
    assert abs(golden["bar_energies"][-1] - 1.05263564259) < 1e-10
    assert abs(np.abs(golden["bar_positions"]).sum() - 1657.181486918721) < 1e-8
    e = golden["sphere_energies"]
    k = next(i for i in range(1, len(e)) if abs(e[i] - e[i - 1]) <= 1e-8 * e[i]) + 1
    assert k == 35 and abs(e[-1] - 0.136396706) < 1e-8
    assert all(b <= a * (1 + 1e-12) for a, b in zip(golden["bar_energies"], golden["bar_energies"][1:]))


def test_oracle_semantics(meshes):
    """The `_dirty` protocol and write-back semantics of arap.h:101-138."""
    P, F = meshes["sphere"]
    mesh = P.astype(np.float32)           # OpenMesh default scalar
    a = O.ArapOracle(mesh, F, np.float64)
    a.setConstraint(37, P[37])
    a.setConstraint(32, P[32] + [0, 0, 0.5])
    assert a.deform(0) and not a.dirty
    # deform(0) with constraints snaps the constrained vertices in the mesh (cast to mesh scalar)
    assert np.allclose(mesh[32], (P[32] + [0, 0, 0.5]).astype(np.float32))
    # deform(a); deform(b) == deform(a+b) while no constraint changes
    m2 = P.astype(np.float32)
    b = O.ArapOracle(m2, F, np.float64)
    b.setConstraint(37, P[37])
    b.setConstraint(32, P[32] + [0, 0, 0.5])
    a.deform(2); a.deform(3)
    b.deform(5)
    assert np.array_equal(mesh, m2)
    # a new constraint re-reads the rest pose from the (deformed) mesh: energy restarts from the new rest
    a.setConstraint(32, P[32] + [0, 0, 0.6])
    assert a.dirty and a.deform(0)
    assert np.array_equal(a.rest().astype(np.float32)[0], mesh[0])


def test_rotation_is_proper_and_optimal():
    rng = np.random.default_rng(11)
    for _ in range(300):
        M = rng.standard_normal((3, 3))
        R = O.rotation_from_covariance(M)
        assert abs(np.linalg.det(R) - 1) < 1e-12 and np.abs(R @ R.T - np.eye(3)).max() < 1e-12
        # arap.h:373-382: R = V diag(1,1,det) U^T of cov = U S V^T  (LAPACK cross-check)
        U, s, Vt = np.linalg.svd(M)
        D = np.diag([1, 1, np.linalg.det(Vt.T @ U.T)])
        assert np.abs(R - Vt.T @ D @ U.T).max() < 1e-9


def test_float_oracle_close_to_double(meshes, golden):
    P, F = meshes["bar"]
    idx, tgt = golden["bar_idx"], golden["bar_tgt"]
    mesh, _, _ = _run(O.ArapOracle, P, F, idx, tgt, 10, precision=np.float32)
    assert np.abs(mesh - golden["bar_positions"]).max() < 2e-4 * bbox_diag(P)


def test_generators():
    P, F = G.icosphere(8)
    assert P.shape == (642, 3) and F.shape == (1280, 3)
    assert np.allclose(np.linalg.norm(P, axis=1), 1)
    e = np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1)
    assert P.shape[0] - len(np.unique(e, axis=0)) + F.shape[0] == 2
    P, F = G.grid_plane(20, 20)
    assert P.shape == (400, 3) and F.shape == (722, 3)
    idx, tgt = G.grid_constraints(20, 20, P)
    assert idx.size == 80
