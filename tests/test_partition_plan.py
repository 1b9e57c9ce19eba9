"""CPU tests of the partitioner (mesh_deform_b200/partition.py): the halo plans of all ranks are mutually consistent
and a partitioned sparse mat-vec with halo exchange reproduces the global one (numpy simulation of what the engine
does with NCCL send/recv)."""
import numpy as np
import pytest
import scipy.sparse as sp

from mesh_deform_b200 import meshgen as G
from mesh_deform_b200 import partition as PT
from oracle.numpy_ref import cotan_weights


def test_block_grid_shapes():
    assert [PT.block_grid(n) for n in (1, 2, 3, 4, 6, 8)] == [(1, 1), (2, 1), (3, 1), (2, 2), (3, 2), (4, 2)]


@pytest.mark.parametrize("owner_kind", ["strips", "blocks"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("mesh", ["grid", "ico"])
def test_partition_plans_are_consistent_and_spmv_matches(mesh, world, owner_kind):
    P, F = G.grid_plane(23, 17) if mesh == "grid" else G.icosphere(6)
    V = P.shape[0]
    owner = PT.strip_owner(P, world) if owner_kind == "strips" else PT.block_owner(P, world)
    assert np.bincount(owner, minlength=world).min() >= V // world - world
    parts = [PT.build_local_part(F, owner, r, world) for r in range(world)]
    assert sum(p.n_owned for p in parts) == V
    assert np.array_equal(np.sort(np.concatenate([p.owned_global for p in parts])), np.arange(V))
    W = cotan_weights(P, F).tocsr()
    x = np.random.default_rng(0).standard_normal((V, 3))
    y_global = W @ x
    # per-rank state: owned values known, halo filled by "exchange"
    local_x = [np.full((p.n_local, 3), np.nan) for p in parts]
    for p, lx in zip(parts, local_x):
        lx[:p.n_owned] = x[p.owned_global]
    for p, lx in zip(parts, local_x):                          # receive side
        for k, q in enumerate(p.neighbor_rank):
            src = parts[q]
            slot = list(src.neighbor_rank).index(p.rank)         # symmetric neighbour lists
            send = src.send_index[src.send_offset[slot]:src.send_offset[slot + 1]]
            n = p.recv_offset[k + 1] - p.recv_offset[k]
            assert send.size == n
            assert (send < src.n_owned).all()
            # the sender's vertices are exactly the ones the receiver expects, in the same order
            assert np.array_equal(src.local_to_global[send], p.local_to_global[p.n_owned + p.recv_offset[k]:p.n_owned + p.recv_offset[k + 1]])
            lx[p.n_owned + p.recv_offset[k]:p.n_owned + p.recv_offset[k + 1]] = local_x[q][send]
    for p, lx in zip(parts, local_x):
        assert not np.isnan(lx).any()
        # local one-ring of owned rows from the LOCAL faces only (what the engine's weight kernels see)
        Wl = cotan_weights(P[p.local_to_global], p.faces).tocsr()
        y_local = (Wl @ lx)[:p.n_owned]
        assert np.abs(y_local - y_global[p.owned_global]).max() < 1e-12
        # local CSR rows of owned vertices are complete (same neighbours, same weights)
        Wg_rows = W[p.owned_global]
        assert Wl[:p.n_owned].nnz == Wg_rows.nnz


def test_local_constraints_include_halo_copies():
    P, F = G.grid_plane(12, 12)
    owner = PT.strip_owner(P, 3)
    idx, tgt = G.grid_constraints(12, 12, P)
    for r in range(3):
        part = PT.build_local_part(F, owner, r, 3)
        li, lt = PT.local_constraints(part, idx, tgt)
        got = set(part.local_to_global[li].tolist())
        want = set(idx.tolist()) & set(part.local_to_global.tolist())
        assert got == want
        assert np.allclose(lt, tgt[[list(idx).index(g) for g in part.local_to_global[li]]])
