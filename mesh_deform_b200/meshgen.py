"""Synthetic meshes and constraint sets for the workloads BASELINE.json names.

Nothing here is on the solve path; these are the inputs of the benchmark configurations:
  * config 3: class-I geodesic icosphere, frequency nu  -> V = 10 nu^2 + 2 (nu = 316 -> 998,562)
  * config 5: n x m grid plane with the topology of the reference's etc/plane.obj
  * the hard-coded anchor / handle index sets of the reference demos
    (reference examples/deform_bar.cpp:30,33, examples/deform_sphere.cpp:60,65,70).
No RNG is involved in vertex positions.
"""
import numpy as np

# reference examples/deform_bar.cpp:30 (handles, the x=6 face) and :33 (anchors, the x=0 face)
BAR_HANDLES = (4, 5, 7, 11, 14, 15, 16, 17, 18, 19, 26, 27, 38, 39, 45, 46, 48, 49, 50, 51, 52, 53, 54, 55, 56, 76,
               77, 131, 132, 201, 202, 203, 204, 205, 206, 207, 208, 209, 210, 211, 212, 213, 214, 215, 216, 217,
               218, 219, 220, 221, 222, 223, 224, 225, 226, 227, 228, 229, 230, 231, 232, 233, 234, 235, 240, 241,
               257, 307, 308, 324, 325, 327, 328, 329, 330, 331, 332, 333, 334, 354, 355)
BAR_ANCHORS = (0, 1, 2, 3, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95, 96, 97, 98, 99, 100,
               101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 116, 117, 118, 119, 120,
               121, 122, 123, 124, 125, 126, 357, 358, 359, 360, 361, 362, 363, 364, 365, 366, 367, 368, 369, 370,
               371, 372, 373, 374, 375, 376, 377, 378, 379, 380, 381, 382, 383, 384, 385)
# reference examples/deform_sphere.cpp:60 (anchor = south pole) and :65,85 (handle = north pole)
SPHERE_ANCHOR = 37
SPHERE_HANDLE = 32


def read_obj(path):
    """Minimal triangle-OBJ reader: vertex index = order of `v` lines, face index = order of `f`
    lines (what OpenMesh's reader gives the reference's adapter, openmesh_adapter.h:74-99)."""
    verts, faces = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, np.float64), np.asarray(faces, np.int32)


def _icosahedron():
    phi = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0],
                  [0, -1, phi], [0, 1, phi], [0, -1, -phi], [0, 1, -phi],
                  [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int32)
    # rotate so that two opposite icosahedron vertices sit on the z axis (poles, like etc/sphere.obj)
    z = v[0]
    x = np.cross([0.0, 0.0, 1.0], z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    rot = np.stack([x, y, z])
    return v @ rot.T, f


def icosphere(nu):
    """Class-I geodesic sphere of frequency nu on the unit sphere.

    Returns (V,3) float64 positions and (F,3) int32 faces with V = 10 nu^2 + 2, F = 20 nu^2.
    Vertex order: 12 icosahedron corners, then the 30 edges' interior points, then the 20 faces'
    interior points row by row.
    """
    nu = int(nu)
    cv, cf = _icosahedron()
    if nu == 1:
        return cv, cf
    edge_id = {}
    for tri in cf:
        for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0])):
            key = (min(a, b), max(a, b))
            if key not in edge_id:
                edge_id[key] = len(edge_id)
    n_edge_pts = nu - 1
    n_face_pts = (nu - 1) * (nu - 2) // 2
    V = 12 + 30 * n_edge_pts + 20 * n_face_pts
    pos = np.zeros((V, 3))
    pos[:12] = cv
    t = np.arange(1, nu) / nu
    for (a, b), e in edge_id.items():
        base = 12 + e * n_edge_pts
        pos[base:base + n_edge_pts] = (1 - t)[:, None] * cv[a] + t[:, None] * cv[b]

    def edge_point(a, b, k):
        """global id of the point k/nu of the way from corner a to corner b (k array, 0..nu)"""
        lo, hi = (a, b) if a < b else (b, a)
        kk = k if a < b else nu - k
        ids = 12 + edge_id[(lo, hi)] * n_edge_pts + (kk - 1)
        ids = np.where(kk == 0, lo, ids)
        ids = np.where(kk == nu, hi, ids)
        return ids

    faces = []
    ii, jj = np.meshgrid(np.arange(nu + 1), np.arange(nu + 1), indexing="ij")
    valid = (ii + jj) <= nu
    for fidx, (A, B, Cc) in enumerate(cf):
        # point (i,j): weight i/nu on B, j/nu on C, rest on A
        gid = np.full((nu + 1, nu + 1), -1, np.int64)
        interior = valid & (ii > 0) & (jj > 0) & (ii + jj < nu)
        base = 12 + 30 * n_edge_pts + fidx * n_face_pts
        # row-major over i then j among interior points
        gid[interior] = base + np.arange(n_face_pts)
        wi = ii[interior] / nu
        wj = jj[interior] / nu
        pos[base:base + n_face_pts] = ((1 - wi - wj)[:, None] * cv[A] + wi[:, None] * cv[B] + wj[:, None] * cv[Cc])
        k = np.arange(nu + 1)
        gid[k, 0] = edge_point(A, B, k)            # j = 0 edge: A -> B
        gid[0, k] = edge_point(A, Cc, k)           # i = 0 edge: A -> C
        gid[k, nu - k] = edge_point(Cc, B, k)      # i + j = nu edge: C -> B
        i0, j0 = np.nonzero(valid & (ii + jj < nu))
        up = np.stack([gid[i0, j0], gid[i0 + 1, j0], gid[i0, j0 + 1]], axis=1)
        i1, j1 = np.nonzero(valid & (ii + jj < nu - 1))
        down = np.stack([gid[i1 + 1, j1], gid[i1 + 1, j1 + 1], gid[i1, j1 + 1]], axis=1)
        faces.append(up)
        faces.append(down)
    faces = np.concatenate(faces).astype(np.int32)
    pos /= np.linalg.norm(pos, axis=1, keepdims=True)
    return pos, faces


def grid_plane(nx, nz, extent=1.0):
    """nx x nz vertex grid on [-extent, extent]^2 in the xz-plane (y = 0), row-major vertex ids
    (id = iz * nx + ix), every quad split by the (ix,iz)-(ix+1,iz+1) diagonal like etc/plane.obj."""
    xs = np.linspace(-extent, extent, nx)
    zs = np.linspace(-extent, extent, nz)
    X, Z = np.meshgrid(xs, zs, indexing="xy")
    pos = np.stack([X.ravel(), np.zeros(nx * nz), Z.ravel()], axis=1)
    ix, iz = np.meshgrid(np.arange(nx - 1), np.arange(nz - 1), indexing="xy")
    v00 = (iz * nx + ix).ravel()
    v10 = v00 + 1
    v01 = v00 + nx
    v11 = v01 + 1
    f = np.empty((2 * v00.size, 3), np.int32)
    f[0::2] = np.stack([v00, v11, v10], axis=1)
    f[1::2] = np.stack([v00, v01, v11], axis=1)
    return pos, f


def rot_x(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def rot_z(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def cap_constraints(pos, anchor_frac=0.05, handle_frac=0.01, angle_deg=30.0, lift=0.3):
    """Config 3's deterministic constraint set: anchors = the
    floor(anchor_frac V) vertices of lowest z pinned at rest; handles = the floor(handle_frac V)
    of highest z moved to Rz(angle) p + (0,0,lift). Returns (indices int32, targets float64)."""
    V = pos.shape[0]
    order = np.lexsort((np.arange(V), pos[:, 2]))
    na, nh = int(anchor_frac * V), int(handle_frac * V)
    anchors = np.sort(order[:na])
    handles = np.sort(order[V - nh:])
    tgt_h = pos[handles] @ rot_z(np.deg2rad(angle_deg)).T + np.array([0.0, 0.0, lift])
    idx = np.concatenate([anchors, handles]).astype(np.int32)
    tgt = np.concatenate([pos[anchors], tgt_h])
    return idx, tgt


def grid_constraints(nx, nz, pos, columns=2, lift=0.3):
    """Config 5's constraints: the `columns` left-most vertex columns pinned, the `columns`
    right-most lifted by (0, lift, 0) -- the up-scaling of etc/anchors.txt / etc/handles.txt."""
    iz = np.arange(nz)
    anchors = np.concatenate([iz * nx + c for c in range(columns)])
    handles = np.concatenate([iz * nx + (nx - 1 - c) for c in range(columns)])
    anchors.sort()
    handles.sort()
    idx = np.concatenate([anchors, handles]).astype(np.int32)
    tgt = np.concatenate([pos[anchors], pos[handles] + np.array([0.0, lift, 0.0])])
    return idx, tgt
