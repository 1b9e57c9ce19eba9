"""tests/golden/make_golden.py -- regenerates tests/golden/*.npz. Run HERE (needs /root/reference).

Inputs: the reference's mesh fixtures reference etc/{bar,sphere,plane}.obj and etc/{anchors,handles}.txt,
converted to arrays (the GPU box has no /root/reference, so the tests read these fixtures instead).
Outputs: golden results of the three small workloads (SURVEY.md section 8d configs 1, 2 and the plane
variant) computed by the C oracle (oracle/arap_oracle.c) and cross-checked in this script against the
independent numpy/scipy restatement (oracle/numpy_ref.py) to 1e-10 before being written.

The reference itself cannot be run (needs Eigen/OpenMesh, absent from this image), so these are
oracle outputs, not reference outputs: "parity unpinned" beyond reference tests/test_cotan.cpp.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle.numpy_ref import NumpyArap  # noqa: E402
from mesh_deform_b200 import meshgen as G  # noqa: E402

REF_ETC = "/root/reference/etc"
HERE = os.path.dirname(os.path.abspath(__file__))


def workloads(meshes):
    P, F = meshes["bar"]
    bar_idx = np.array(G.BAR_ANCHORS + G.BAR_HANDLES, np.int32)
    bar_tgt = np.concatenate([P[list(G.BAR_ANCHORS)], P[list(G.BAR_HANDLES)] @ G.rot_x(np.pi / 4).T])
    yield "bar", P, F, bar_idx, bar_tgt, 10
    P, F = meshes["sphere"]
    yield "sphere", P, F, np.array([G.SPHERE_ANCHOR, G.SPHERE_HANDLE], np.int32), \
        np.stack([P[G.SPHERE_ANCHOR], P[G.SPHERE_HANDLE] + [0, 0, 0.5]]), 35
    P, F = meshes["plane"]
    anchors = np.loadtxt(os.path.join(REF_ETC, "anchors.txt"), dtype=np.int64).ravel()
    handles = np.loadtxt(os.path.join(REF_ETC, "handles.txt"), dtype=np.int64).ravel()
    yield "plane", P, F, np.concatenate([anchors, handles]).astype(np.int32), \
        np.concatenate([P[anchors], P[handles] + [0, 0.3, 0]]), 8


def main():
    meshes = {}
    for name in ("bar", "sphere", "plane"):
        meshes[name] = G.read_obj(os.path.join(REF_ETC, name + ".obj"))
    np.savez_compressed(os.path.join(HERE, "meshes.npz"),
                        **{f"{n}_{k}": v for n, (P, F) in meshes.items() for k, v in (("V", P), ("F", F))})
    out = {}
    for name, P, F, idx, tgt, iters in workloads(meshes):
        mesh_c, mesh_n = P.copy(), P.copy()
        a, b = O.ArapOracle(mesh_c, F, np.float64), NumpyArap(mesh_n, F)
        for i, t in zip(idx, tgt):
            a.setConstraint(i, t)
            b.setConstraint(i, t)
        energies = []
        for _ in range(iters):
            assert a.deform(1) and b.deform(1)
            energies.append(a.energy())
            assert abs(a.energy() - b.energy()) <= 1e-10 * max(1.0, abs(b.energy())), name
        assert np.abs(mesh_c - mesh_n).max() < 1e-10, name
        rp, ci, w = a.cotanWeights()
        out.update({f"{name}_idx": idx, f"{name}_tgt": tgt, f"{name}_iters": np.int32(iters),
                    f"{name}_rowptr": rp, f"{name}_colidx": ci, f"{name}_w": w,
                    f"{name}_positions": mesh_c, f"{name}_energies": np.array(energies),
                    f"{name}_rotations": a.rotations(), f"{name}_nfree": np.int32(a.nFree)})
        print(name, "V", P.shape[0], "nnz", ci.size, "nFree", a.nFree, "E_last", energies[-1])
    np.savez_compressed(os.path.join(HERE, "arap_golden.npz"), **out)

    # trajectory: reference tests/test_trajectory.cpp:23-33 key poses, sampled densely by the oracle
    def T(t=(0, 0, 0), R=np.eye(3)):
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = t
        return M
    pose0 = np.eye(4)
    pose1 = T((0, 0, 1)) @ pose0
    pose2 = T((0, 0, 1)) @ T(R=G.rot_x(np.pi / 4)) @ pose1
    pose3 = T((0, 0, 1)) @ pose2
    traj = O.TrajectorySE3Oracle()
    for p in (pose0, pose1, pose2, pose3):
        traj.addKeyPose(p)
    us = np.linspace(0, 1, 33)
    samples = np.stack([traj(u) for u in us])
    np.savez_compressed(os.path.join(HERE, "trajectory_golden.npz"), key_poses=np.stack([pose0, pose1, pose2, pose3]),
                        u=us, samples=samples)
    print("trajectory endpoints ok:", np.abs(samples[0] - pose0).max(), np.abs(samples[-1] - pose3).max())


if __name__ == "__main__":
    main()
