"""Vertex partitioning of one mesh over several GPUs (BASELINE.json configs[4]).

Host-side only: decides who owns which vertex and builds, for one rank, the local mesh the engine works on
(owned vertices first, then the one-ring halo grouped by owner, plus every face touching an owned vertex) and the
halo plan handed to `arap_attach_partition` (include/arap_b200.h). Every rank computes its own part from the
global mesh deterministically, so no communication is needed to agree on the plan: the vertices rank p sends to
rank q and the vertices rank q expects from rank p are the same set in the same (global id) order.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class LocalPart:
    rank: int
    world: int
    n_owned: int
    local_to_global: np.ndarray      # (n_local,) global vertex ids: owned (ascending) then halo (by owner, then id)
    faces: np.ndarray                # (f_local, 3) int32, LOCAL vertex indices
    neighbor_rank: np.ndarray        # (n_neighbors,) int32
    send_offset: np.ndarray          # (n_neighbors + 1,) int32
    send_index: np.ndarray           # owned LOCAL indices, grouped by neighbour
    recv_offset: np.ndarray          # (n_neighbors + 1,) int32

    @property
    def n_local(self):
        return int(self.local_to_global.size)

    @property
    def owned_global(self):
        return self.local_to_global[:self.n_owned]


def strip_owner(positions, n_parts):
    """Owner rank per vertex: n_parts slabs of (almost) equal vertex count along the longest bounding-box axis.
    For the grid plane of configs[4] these are strips of rows; halos are then two boundary rows per rank."""
    positions = np.asarray(positions)
    axis = int(np.argmax(positions.max(0) - positions.min(0)))
    order = np.lexsort((np.arange(positions.shape[0]), positions[:, axis]))
    owner = np.empty(positions.shape[0], np.int32)
    bounds = np.linspace(0, positions.shape[0], n_parts + 1).astype(np.int64)
    for p in range(n_parts):
        owner[order[bounds[p]:bounds[p + 1]]] = p
    return owner


def block_grid(n_parts):
    """(pa, pb) with pa * pb == n_parts and pa >= pb as square as possible: 2 -> (2,1), 4 -> (2,2), 8 -> (4,2)."""
    pb = int(np.floor(np.sqrt(n_parts)))
    while n_parts % pb:
        pb -= 1
    return n_parts // pb, pb


def block_owner(positions, n_parts):
    """Owner rank per vertex for a pa x pb arrangement of blocks: pa slabs of equal vertex count along the longest
    bounding-box axis, each cut into pb pieces of equal count along the second-longest. Blocks have up to 8 neighbours
    (strips: 2) but a quarter of the halo of strips on a square grid."""
    positions = np.asarray(positions)
    pa, pb = block_grid(n_parts)
    ext = positions.max(0) - positions.min(0)
    a0 = int(np.argmax(ext))
    ext2 = ext.copy()
    ext2[a0] = -1.0
    a1 = int(np.argmax(ext2))
    n = positions.shape[0]
    owner = np.empty(n, np.int32)
    order = np.lexsort((np.arange(n), positions[:, a0]))
    bounds = np.linspace(0, n, pa + 1).astype(np.int64)
    for p in range(pa):
        slab = order[bounds[p]:bounds[p + 1]]
        sub = slab[np.lexsort((slab, positions[slab, a1]))]
        b2 = np.linspace(0, sub.size, pb + 1).astype(np.int64)
        for q in range(pb):
            owner[sub[b2[q]:b2[q + 1]]] = p * pb + q
    return owner


def build_local_part(faces, owner, rank, world):
    """The local mesh and halo plan of `rank` (see module docstring)."""
    faces = np.asarray(faces)
    fo = owner[faces]                                            # (F, 3) owner of each corner
    mine = (fo == rank).any(1)
    lf = faces[mine]                                             # every face touching an owned vertex
    lfo = fo[mine]
    owned = np.flatnonzero(owner == rank)
    touched = np.unique(lf)
    halo = touched[owner[touched] != rank]
    halo = halo[np.lexsort((halo, owner[halo]))]
    local_to_global = np.concatenate([owned, halo]).astype(np.int64)
    g2l = np.full(owner.size, -1, np.int32)
    g2l[local_to_global] = np.arange(local_to_global.size, dtype=np.int32)
    nbr, counts = np.unique(owner[halo], return_counts=True)
    recv_offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    send_lists = []
    for q in nbr:
        shared = lf[(lfo == q).any(1)]                           # faces touching both an owned and a q-owned vertex
        cand = np.unique(shared[owner[shared] == rank])          # ... their owned corners, ascending global id
        send_lists.append(g2l[cand])
    send_offset = np.concatenate([[0], np.cumsum([len(s) for s in send_lists])]).astype(np.int32)
    send_index = (np.concatenate(send_lists) if send_lists else np.zeros(0)).astype(np.int32)
    return LocalPart(rank=rank, world=world, n_owned=int(owned.size), local_to_global=local_to_global,
                     faces=np.ascontiguousarray(g2l[lf], dtype=np.int32), neighbor_rank=nbr.astype(np.int32),
                     send_offset=send_offset, send_index=send_index, recv_offset=recv_offset)


def local_constraints(part, global_indices, targets):
    """Restrict a global constraint set to the vertices this rank holds (owned AND halo: halo copies of constrained
    vertices must carry their target and their constrained flag too)."""
    global_indices = np.asarray(global_indices)
    targets = np.asarray(targets)
    pos = np.full(int(max(global_indices.max(initial=0), part.local_to_global.max(initial=0))) + 1, -1, np.int64)
    pos[part.local_to_global] = np.arange(part.n_local)
    loc = pos[global_indices]
    keep = loc >= 0
    return loc[keep].astype(np.int32), targets[keep]
