"""oracle/numpy_ref.py -- TEST INFRASTRUCTURE ONLY.

A second, independently written fp64 restatement of reference inc/deform/arap.h in vectorised
numpy/scipy (LAPACK SVD via numpy.linalg.svd, SuperLU via scipy.sparse.linalg.splu). Its only job
is to cross-check oracle/arap_oracle.c: the reference's own tests pin nothing past the cotan
weights (tests/test_cotan.cpp), so two independently written restatements that agree to ~1e-12
are the strongest pin available for the local step / RHS / solve (arap.h:354-430) in an image
where the reference (needs Eigen) cannot be compiled. Used by tests/ and tests/golden/make_golden.py.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def cotan_weights(p, faces):
    """arap.h:182-239 -> scipy CSR (V x V), duplicates summed, sorted indices."""
    V = p.shape[0]
    v0, v1, v2 = p[faces[:, 0]], p[faces[:, 1]], p[faces[:, 2]]
    l0 = np.sqrt(np.maximum(1e-8, ((v1 - v0) ** 2).sum(1)))
    l1 = np.sqrt(np.maximum(1e-8, ((v2 - v1) ** 2).sum(1)))
    l2 = np.sqrt(np.maximum(1e-8, ((v0 - v2) ** 2).sum(1)))
    s = 0.5 * (l0 + l1 + l2)
    with np.errstate(invalid="ignore"):
        area = np.sqrt(s * (s - l0) * (s - l1) * (s - l2))
    area = np.where(area > 1e-8, area, 1e-8)            # std::max(1e-8, NaN) == 1e-8
    denom = 1.0 / (4.0 * area)
    cot0 = np.maximum(1e-10, (-l0 * l0 + l1 * l1 + l2 * l2) * denom)
    cot1 = np.maximum(1e-10, (l0 * l0 - l1 * l1 + l2 * l2) * denom)
    cot2 = np.maximum(1e-10, (l0 * l0 + l1 * l1 - l2 * l2) * denom)
    i = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2]])
    j = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0]])
    w = 0.5 * np.concatenate([cot0, cot1, cot2])
    W = sp.coo_matrix((np.concatenate([w, w]), (np.concatenate([i, j]), np.concatenate([j, i]))), shape=(V, V)).tocsr()
    W.sum_duplicates()
    W.sort_indices()
    return W


def rotations(W, p, pp):
    """arap.h:354-384: R_i = V diag(1,1,det(V U^T)) U^T of S_i = sum_j w_ij (p_i-p_j)(p'_i-p'_j)^T."""
    V = p.shape[0]
    rows = np.repeat(np.arange(V), np.diff(W.indptr))
    cols = W.indices
    e = p[rows] - p[cols]
    ep = pp[rows] - pp[cols]
    outer = W.data[:, None, None] * e[:, :, None] * ep[:, None, :]
    S = np.zeros((V, 3, 3))
    np.add.at(S, rows, outer)
    U, sig, Vt = np.linalg.svd(S)
    Vm = np.transpose(Vt, (0, 2, 1))
    Ut = np.transpose(U, (0, 2, 1))
    det = np.linalg.det(Vm @ Ut)
    D = np.zeros((V, 3, 3))
    D[:, 0, 0] = 1
    D[:, 1, 1] = 1
    D[:, 2, 2] = det
    return Vm @ D @ Ut, sig


def energy(W, p, pp, R):
    """Sorkine-Alexa eq. (3): sum over directed CSR entries of w_ij |(p'_i-p'_j) - R_i (p_i-p_j)|^2."""
    V = p.shape[0]
    rows = np.repeat(np.arange(V), np.diff(W.indptr))
    cols = W.indices
    e = p[rows] - p[cols]
    ep = pp[rows] - pp[cols]
    res = ep - np.einsum("nab,nb->na", R[rows], e)
    return float((W.data * (res ** 2).sum(1)).sum())


class NumpyArap:
    """Same call protocol as the reference class (ctor / setConstraint / deform)."""

    def __init__(self, positions, faces):
        self.mesh = positions                       # (V,3) array mutated by deform(), like Mesh&
        self.faces = np.asarray(faces, np.int64)
        self.con = {}
        self.dirty = True

    def setConstraint(self, i, loc):
        self.con[int(i)] = np.asarray(loc, np.float64)
        self.dirty = True

    def deform(self, n):
        V = self.mesh.shape[0]
        if self.dirty:
            self.p = np.array(self.mesh, dtype=np.float64)
            self.pp = self.p.copy()
            self.W = cotan_weights(self.p, self.faces)
            self.R = np.tile(np.eye(3), (V, 1, 1))
            is_con = np.zeros(V, bool)
            if self.con:
                idx = np.fromiter(self.con.keys(), dtype=np.int64)
                is_con[idx] = True
                self.pp[idx] = np.stack([self.con[k] for k in idx])
            self.free = np.flatnonzero(~is_con)
            self.is_con = is_con
            if self.free.size == V:
                return True
            Lfull = (sp.diags(np.asarray(self.W.sum(1)).ravel()) - self.W).tocsr()
            self.L = Lfull[self.free][:, self.free].tocsc()
            Wfc = self.W[self.free][:, np.flatnonzero(is_con)]
            self.bFixed = Wfc @ self.pp[is_con]
            try:
                self.lu = spla.splu(self.L)
            except RuntimeError:
                return False
            self.dirty = False
        rows = np.repeat(np.arange(V), np.diff(self.W.indptr))
        cols = self.W.indices
        e = self.p[rows] - self.p[cols]
        for _ in range(n):
            self.R, _ = rotations(self.W, self.p, self.pp)
            contrib = 0.5 * self.W.data[:, None] * np.einsum("nab,nb->na", self.R[rows] + self.R[cols], e)
            b = np.zeros((V, 3))
            np.add.at(b, rows, contrib)
            b = b[self.free] + self.bFixed
            self.pp[self.free] = self.lu.solve(b)
        self.mesh[...] = self.pp.astype(self.mesh.dtype)
        return True

    def energy(self):
        return energy(self.W, self.p, self.pp, self.R)
