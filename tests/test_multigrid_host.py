"""CPU tests of the host-side multigrid setup (mesh_deform_b200/csrc/mg_setup.cpp, mg_partition.cpp) through the test shim
tests/cpp/libmg_host.so: the hierarchy is a good preconditioner, and -- for partitioned meshes -- every rank's share of the
global hierarchy, driven through exactly the exchange sequence of Engine::vcycle_partitioned, reproduces the global V-cycle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT
from mesh_deform_b200 import meshgen as G, partition as PT

CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def shim():
    subprocess.check_call(["make", "-s", "-C", CPP, "libmg_host.so"])
    L = C.CDLL(os.path.join(CPP, "libmg_host.so"))
    L.mgshim_error.restype = C.c_char_p
    L.mgshim_omega.restype = C.c_double
    L.mgshim_build_blocks.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _csr(L, getter_dims, getter, l, which):
    r, c, z = C.c_int(), C.c_int(), C.c_int()
    getter_dims(l, which, C.byref(r), C.byref(c), C.byref(z))
    rp, ci, v = np.zeros(r.value + 1, np.int32), np.zeros(z.value, np.int32), np.zeros(z.value)
    getter(l, which, _p(rp), _p(ci), _p(v))
    return sp.csr_matrix((v, ci, rp), shape=(r.value, c.value))


def build_global(L, P, F, con_idx, owner, coarse=64):
    V = P.shape[0]
    nnz = C.c_int()
    faces = np.ascontiguousarray(F, np.int32)
    xyz = np.ascontiguousarray(P, np.float64)
    L.mgshim_csr(V, faces.shape[0], _p(faces), _p(xyz), None, None, None, C.byref(nnz))
    rp, ci, w = np.zeros(V + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
    L.mgshim_csr(V, faces.shape[0], _p(faces), _p(xyz), _p(rp), _p(ci), _p(w), C.byref(nnz))
    con = np.zeros(V, np.uint8)
    con[con_idx] = 1
    order = np.zeros(V, np.int32)
    L.mgshim_morton(V, _p(xyz), _p(order))
    assert sorted(order.tolist()) == list(range(V))
    block = np.ascontiguousarray(owner, np.int32)
    nl = L.mgshim_build_blocks(V, _p(rp), _p(ci), _p(w), _p(con), 0.0, coarse, _p(order), _p(block))
    levels = []
    for l in range(nl):
        A = _csr(L, L.mgshim_dims, L.mgshim_get, l, 0)
        lv = {"A": A, "omega": L.mgshim_omega(l)}
        invd = np.zeros(A.shape[0])
        L.mgshim_inv_diag(l, _p(invd))
        lv["invd"] = invd
        blk = np.zeros(A.shape[0], np.int32)
        L.mgshim_block(l, _p(blk))
        lv["block"] = blk
        if l + 1 < nl:
            lv["P"] = _csr(L, L.mgshim_dims, L.mgshim_get, l, 1)
            lv["R"] = _csr(L, L.mgshim_dims, L.mgshim_get, l, 2)
        levels.append(lv)
    n_c = L.mgshim_ncoarse()
    assert L.mgshim_has_inverse()
    inv = np.zeros((n_c, n_c))
    L.mgshim_coarse_inverse(_p(inv))
    return levels, inv, (rp, ci, w)


def vcycle_global(levels, inv, b, l=0):
    lv = levels[l]
    if l == len(levels) - 1:
        return inv @ b
    s = (lv["omega"] * lv["invd"])[:, None]
    x = s * b
    r = b - lv["A"] @ x
    x = x + lv["P"] @ vcycle_global(levels, inv, lv["R"] @ r, l + 1)
    return x + s * (b - lv["A"] @ x)


def test_global_csr_matches_numpy_reference(shim):
    from oracle import numpy_ref as NR
    P, F = G.grid_plane(9, 7)
    levels, inv, (rp, ci, w) = build_global(shim, P, F, [0], np.zeros(P.shape[0], np.int32), coarse=8)
    W = NR.cotan_weights(P, F).tocsr()
    assert np.array_equal(rp, W.indptr) and np.array_equal(ci, W.indices)
    assert np.abs(w - W.data).max() < 1e-14


def test_vcycle_preconditioned_cg_converges_fast(shim):
    nx, nz = 48, 40
    P, F = G.grid_plane(nx, nz)
    idx, _ = G.grid_constraints(nx, nz, P)
    levels, inv, _ = build_global(shim, P, F, idx, np.zeros(P.shape[0], np.int32))
    assert len(levels) >= 3
    A = levels[0]["A"]
    rng = np.random.default_rng(3)
    free = levels[0]["invd"] > 0
    b = rng.standard_normal((A.shape[0], 3)) * free[:, None]
    x = np.zeros_like(b)
    r = b.copy()
    z = vcycle_global(levels, inv, r)
    d = z.copy()
    rho = (r * z).sum(0)
    for it in range(40):
        q = A @ d
        alpha = rho / (d * q).sum(0)
        x += alpha * d
        r -= alpha * q
        if np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b):
            break
        z = vcycle_global(levels, inv, r)
        rho_new = (r * z).sum(0)
        d = z + (rho_new / rho) * d
        rho = rho_new
    assert it < 25, it


def quadrant_owner(P):
    """A 2 x 2 block partition (every rank has three neighbours, one of them only across a corner)."""
    cx, cz = np.median(P[:, 0]), np.median(P[:, 2])
    return ((P[:, 0] > cx).astype(np.int32) + 2 * (P[:, 2] > cz).astype(np.int32)).astype(np.int32)


@pytest.mark.parametrize("replicate_rows", [0, 60, 100000])
@pytest.mark.parametrize("world,kind", [(1, "strips"), (2, "strips"), (3, "strips"), (4, "quadrants")])
def test_partitioned_vcycle_equals_global_vcycle(shim, world, kind, replicate_rows):
    """Every rank's slice + the exchange sequence of Engine::vcycle_partitioned == the V-cycle of the whole hierarchy.
    replicate_rows: levels with at most that many rows are kept whole on every rank (0: only the coarsest; 100000: all but
    the finest)."""
    nx, nz = 40, 36
    P, F = G.grid_plane(nx, nz)
    idx, _ = G.grid_constraints(nx, nz, P)
    owner = PT.strip_owner(P, world) if kind == "strips" else quadrant_owner(P)
    levels, inv, _ = build_global(shim, P, F, idx, owner)
    nl = len(levels)
    assert nl >= 3
    for l in range(nl - 1):                                   # aggregates never straddle two blocks
        Pm = levels[l]["P"].tocsr()
        blk_f, blk_c = levels[l]["block"], levels[l + 1]["block"]
        for i in range(Pm.shape[0]):
            cols, vals = Pm.indices[Pm.indptr[i]:Pm.indptr[i + 1]], Pm.data[Pm.indptr[i]:Pm.indptr[i + 1]]
            if cols.size and vals.max() > 0.5:                # the entry of the row's own aggregate
                assert blk_c[cols[np.argmax(vals)]] == blk_f[i]
    # slice for every rank
    ranks = []
    for r in range(world):
        part = PT.build_local_part(F, owner, r, world)
        l2g = np.ascontiguousarray(part.local_to_global, np.int32)
        ok = shim.mgshim_slice_replicated(r, part.n_owned, part.n_local, _p(l2g), replicate_rows)
        assert ok, shim.mgshim_error().decode()
        Lr = shim.mgshim_first_replicated()
        assert 1 <= Lr <= nl - 1
        if replicate_rows == 0:
            assert Lr == nl - 1
        if replicate_rows >= 100000:
            assert Lr == 1
        lv = []
        for l in range(nl):
            n_own, n_halo, n_nbr, n_send, omega = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
            shim.mgshim_local_level(l, C.byref(n_own), C.byref(n_halo), C.byref(n_nbr), C.byref(n_send), C.byref(omega))
            d = {"n_own": n_own.value, "n_halo": n_halo.value, "omega": omega.value}
            gid = np.zeros(n_own.value + n_halo.value, np.int32)
            shim.mgshim_local_global_id(l, _p(gid))
            d["gid"] = gid
            if l == 0:
                d["plan"] = (part.neighbor_rank, part.send_offset, part.send_index, part.recv_offset)
                A0 = levels[0]["A"].tocsr()[gid[:part.n_owned]]          # the engine applies level 0 matrix-free
                g2l = np.full(P.shape[0], -1)
                g2l[gid] = np.arange(gid.size)
                assert (g2l[A0.indices] >= 0).all()
                d["A"] = sp.csr_matrix((A0.data, g2l[A0.indices], A0.indptr), shape=(part.n_owned, gid.size))
                d["invd"] = levels[0]["invd"][gid[:part.n_owned]]
            else:
                nbr, so, si, ro = (np.zeros(n_nbr.value, np.int32), np.zeros(n_nbr.value + 1, np.int32), np.zeros(n_send.value, np.int32),
                                   np.zeros(n_nbr.value + 1, np.int32))
                if l < nl - 1:
                    shim.mgshim_local_plan(l, _p(nbr), _p(so), _p(si), _p(ro))
                    d["A"] = _csr(shim, shim.mgshim_local_dims, shim.mgshim_local_get, l, 0)
                d["plan"] = (nbr, so, si, ro)
                invd = np.zeros(n_own.value)
                shim.mgshim_local_inv_diag(l, _p(invd))
                d["invd"] = invd
            if l + 1 < nl:
                d["P"] = _csr(shim, shim.mgshim_local_dims, shim.mgshim_local_get, l, 1)
                d["R"] = _csr(shim, shim.mgshim_local_dims, shim.mgshim_local_get, l, 2)
            lv.append(d)
        ranks.append(lv)

    def exchange(l, vecs):
        """halo refresh of one level's vector on every rank (what Transport::exchange does)"""
        for r in range(world):
            nbr, so, si, ro = ranks[r][l]["plan"]
            n_own = ranks[r][l]["n_own"]
            for k, q in enumerate(nbr):
                qn, qso, qsi, _ = ranks[q][l]["plan"]
                slot = list(qn).index(r)
                sent = vecs[q][qsi[qso[slot]:qso[slot + 1]]]
                assert sent.shape[0] == ro[k + 1] - ro[k]
                # the sender's entries are exactly the receiver's halo entries, in the same order
                assert np.array_equal(ranks[q][l]["gid"][qsi[qso[slot]:qso[slot + 1]]], ranks[r][l]["gid"][n_own + ro[k]:n_own + ro[k + 1]])
                vecs[r][n_own + ro[k]:n_own + ro[k + 1]] = sent

    rng = np.random.default_rng(9)
    free = levels[0]["invd"] > 0
    b = rng.standard_normal((P.shape[0], 3)) * free[:, None]
    want = vcycle_global(levels, inv, b)

    def zeros(l):
        return [np.zeros((ranks[r][l]["n_own"] + ranks[r][l]["n_halo"], 3)) for r in range(world)]
    x, res, bb, x2 = [None] * nl, [None] * nl, [None] * nl, [None] * nl
    # level 0 pre-smoothing on owned rows (cg_update_mg_kernel), then the engine's sequence
    x[0], res[0], x2[0] = zeros(0), zeros(0), zeros(0)
    bb[0] = zeros(0)
    for r in range(world):
        lv = ranks[r][0]
        n = lv["n_own"]
        bb[0][r][:n] = b[lv["gid"][:n]]
        x[0][r][:n] = (lv["omega"] * lv["invd"])[:, None] * bb[0][r][:n]
    exchange(0, x[0])
    for r in range(world):
        lv = ranks[r][0]
        n = lv["n_own"]
        res[0][r][:n] = bb[0][r][:n] - lv["A"] @ x[0][r]
    exchange(0, res[0])
    for l in range(nl - 1):
        c = l + 1
        x[c], res[c], bb[c], x2[c] = zeros(c), zeros(c), zeros(c), zeros(c)
        for r in range(world):
            f, cl = ranks[r][l], ranks[r][c]
            rows = f["R"].shape[0]
            bb[c][r][:rows] = f["R"] @ res[l][r]
        if c == Lr:                                             # entering the replicated part: every rank filled its rows; the all-reduce
            total = sum(bb[c][r] for r in range(world))
            for r in range(world):
                bb[c][r] = total.copy()
        for r in range(world):
            cl = ranks[r][c]
            rows = ranks[r][l]["R"].shape[0]
            x[c][r][:rows] = (cl["omega"] * cl["invd"])[:rows, None] * bb[c][r][:rows]
        if c == nl - 1:
            break
        exchange(c, x[c])                                       # no-op on replicated levels (no neighbours)
        for r in range(world):
            cl = ranks[r][c]
            n = cl["n_own"]
            res[c][r][:n] = bb[c][r][:n] - cl["A"] @ x[c][r]
        exchange(c, res[c])
    for r in range(world):
        x2[nl - 1][r] = inv @ bb[nl - 1][r]
    for l in range(nl - 2, -1, -1):
        c = l + 1
        if c < Lr:
            exchange(c, x2[c])
        for r in range(world):
            f = ranks[r][l]
            n = f["n_own"]
            x[l][r][:n] += f["P"] @ x2[c][r]
        exchange(l, x[l])
        for r in range(world):
            f = ranks[r][l]
            n = f["n_own"]
            x2[l][r][:n] = x[l][r][:n] + (f["omega"] * f["invd"])[:, None] * (bb[l][r][:n] - f["A"] @ x[l][r])
    got = np.zeros_like(want)
    for r in range(world):
        lv = ranks[r][0]
        got[lv["gid"][:lv["n_own"]]] = x2[0][r][:lv["n_own"]]
    assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
