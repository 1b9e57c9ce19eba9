"""ctypes binding of the C ABI in include/arap_b200.h (libarap_b200.so).

This is the only way Python reaches the engine: plain pointers and sizes, no torch types. The
library is loaded from inside the package directory (built in-tree by `__graft_entry__.build()` or
`make -C mesh_deform_b200/csrc`). There is NO fallback: if the shared library is missing or cannot
be loaded the import of the symbols fails loudly, and every entry point fails with ARAP_ERR_CUDA
when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARAP_B200_LIB") or os.path.join(_HERE, "libarap_b200.so")   # override: kernel-variant experiments

ARAP_OK = 0
ARAP_UNCONSTRAINED = 1
ARAP_NOT_CONVERGED = 2
ARAP_ERR_INVALID = -1
ARAP_ERR_CUDA = -2
ARAP_ERR_SOLVER = -3
ARAP_ERR_ALLOC = -4

SOLVER_AUTO, SOLVER_PCG_JACOBI, SOLVER_PCG_MG = 0, 1, 2
K_COUNT_MAX = 32

# every symbol include/arap_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "arap_default_options", "arap_create", "arap_destroy", "arap_set_constraints", "arap_is_dirty",
    "arap_prepare", "arap_iterate", "arap_get_positions", "arap_deform", "arap_get_csr_nnz", "arap_get_csr",
    "arap_get_free_map", "arap_get_rotations", "arap_get_rhs", "arap_get_render_buffers", "arap_energy", "arap_get_solver_stats", "arap_profile_enable",
    "arap_deform_async", "arap_deform_wait",
    "arap_profile_reset", "arap_profile_get", "arap_kernel_name", "arap_timer_start", "arap_timer_stop",
    "arap_synchronize", "arap_host_alloc", "arap_host_free", "arap_last_error", "arap_create_error",
    "arap_abi_version", "arap_batch_create", "arap_batch_destroy", "arap_batch_set_constraints", "arap_batch_prepare",
    "arap_batch_iterate", "arap_batch_get_positions", "arap_batch_handle", "arap_comm_unique_id", "arap_attach_partition",
    "arap_trajectory_create", "arap_trajectory_destroy", "arap_trajectory_add_key_pose", "arap_trajectory_evaluate",
    "arap_rigid_conjugate", "arap_set_rigid_constraints", "arap_batch_set_rigid_constraints",
    "arap_partition_set_global_mesh", "arap_partition_comm_benchmark",
)


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("solver", C.c_int32),
                ("max_cg_iterations", C.c_int32), ("cg_tolerance", C.c_double),
                ("cg_check_interval", C.c_int32), ("profile", C.c_int32), ("position_tolerance", C.c_double)]


class SolverStats(C.Structure):
    _fields_ = [("cg_iterations_total", C.c_int64), ("global_steps", C.c_int32), ("last_cg_iterations", C.c_int32),
                ("last_relative_residual", C.c_double), ("last_converged", C.c_int32), ("mg_levels", C.c_int32),
                ("mg_operator_complexity", C.c_double), ("setup_host_ms", C.c_double), ("cg_graph", C.c_int32), ("mg_global", C.c_int32), ("last_position_error", C.c_double),
                ("comm_exchanges_per_cg_iteration", C.c_int32), ("comm_allreduces_per_cg_iteration", C.c_int32),
                ("comm_halo_bytes_per_cg_iteration", C.c_int64), ("tile_max_halo", C.c_int32), ("renumbered", C.c_int32),
                ("setup_device_ms", C.c_double), ("launches_per_cg_iteration", C.c_int32), ("reserved1", C.c_int32)]


class GlobalMesh(C.Structure):
    _fields_ = [("n_vertices", C.c_int32), ("n_faces", C.c_int32), ("faces", C.c_void_p), ("rest_xyz", C.c_void_p),
                ("rest_scalar_bytes", C.c_int32), ("owner", C.c_void_p), ("local_to_global", C.c_void_p),
                ("n_constrained", C.c_int32), ("constrained", C.c_void_p)]


class PartitionPlan(C.Structure):
    _fields_ = [("n_owned", C.c_int32), ("n_neighbors", C.c_int32), ("neighbor_rank", C.c_void_p), ("send_offset", C.c_void_p),
                ("send_index", C.c_void_p), ("recv_offset", C.c_void_p)]


TRANSPORT_NCCL, TRANSPORT_IN_PROCESS, TRANSPORT_PEER, TRANSPORT_PEER_IN_PROCESS = 0, 1, 2, 3


class Profile(C.Structure):
    _fields_ = [("launches", C.c_int64 * K_COUNT_MAX), ("milliseconds", C.c_double * K_COUNT_MAX)]


_lib = None


class EngineMissingError(RuntimeError):
    pass


def lib():
    """Load libarap_b200.so (once). Raises EngineMissingError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineMissingError(
            f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C mesh_deform_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    L.arap_default_options.argtypes = [C.POINTER(Options)]
    L.arap_default_options.restype = None
    L.arap_create.argtypes = [vp, i32, i32, i32, C.POINTER(Options), C.POINTER(vp)]
    L.arap_destroy.argtypes = [vp]
    L.arap_destroy.restype = None
    L.arap_set_constraints.argtypes = [vp, i32, vp, vp, i32]
    L.arap_is_dirty.argtypes = [vp]
    L.arap_prepare.argtypes = [vp, vp, i32]
    L.arap_iterate.argtypes = [vp, i32]
    L.arap_get_positions.argtypes = [vp, vp, i32]
    L.arap_deform.argtypes = [vp, vp, i32, i32]
    L.arap_deform_async.argtypes = [vp, vp, i32, i32]
    L.arap_deform_wait.argtypes = [vp]
    L.arap_get_csr_nnz.argtypes = [vp, C.POINTER(i32)]
    L.arap_get_csr.argtypes = [vp, vp, vp, vp]
    L.arap_get_free_map.argtypes = [vp, vp, C.POINTER(i32)]
    L.arap_get_rotations.argtypes = [vp, vp]
    L.arap_get_rhs.argtypes = [vp, vp]
    L.arap_get_render_buffers.argtypes = [vp, vp, vp, i32]
    L.arap_energy.argtypes = [vp, C.POINTER(C.c_double)]
    L.arap_get_solver_stats.argtypes = [vp, C.POINTER(SolverStats)]
    L.arap_profile_enable.argtypes = [vp, i32]
    L.arap_profile_reset.argtypes = [vp]
    L.arap_profile_get.argtypes = [vp, C.POINTER(Profile)]
    L.arap_kernel_name.argtypes = [i32]
    L.arap_kernel_name.restype = C.c_char_p
    L.arap_timer_start.argtypes = [vp]
    L.arap_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    L.arap_synchronize.argtypes = [vp]
    L.arap_batch_create.argtypes = [vp, i32, i32, i32, i32, C.POINTER(Options), C.POINTER(vp)]
    L.arap_batch_destroy.argtypes = [vp]
    L.arap_batch_destroy.restype = None
    L.arap_batch_set_constraints.argtypes = [vp, i32, vp, vp, i32]
    L.arap_batch_prepare.argtypes = [vp, vp, i32]
    L.arap_batch_iterate.argtypes = [vp, i32]
    L.arap_batch_get_positions.argtypes = [vp, vp, i32]
    L.arap_batch_handle.argtypes = [vp]
    L.arap_batch_handle.restype = vp
    L.arap_trajectory_create.argtypes = [C.POINTER(vp)]
    L.arap_trajectory_destroy.argtypes = [vp]
    L.arap_trajectory_destroy.restype = None
    L.arap_trajectory_add_key_pose.argtypes = [vp, vp]
    L.arap_trajectory_evaluate.argtypes = [vp, i32, vp, vp]
    L.arap_rigid_conjugate.argtypes = [vp, vp, vp]
    L.arap_set_rigid_constraints.argtypes = [vp, i32, vp, vp, i32, vp]
    L.arap_batch_set_rigid_constraints.argtypes = [vp, i32, vp, vp, i32, vp]
    L.arap_comm_unique_id.argtypes = [vp, i32]
    L.arap_attach_partition.argtypes = [vp, C.POINTER(PartitionPlan), i32, i32, i32, vp, i32]
    L.arap_partition_set_global_mesh.argtypes = [vp, C.POINTER(GlobalMesh)]
    L.arap_partition_comm_benchmark.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.arap_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.arap_host_free.argtypes = [vp]
    L.arap_last_error.argtypes = [vp]
    L.arap_last_error.restype = C.c_char_p
    L.arap_create_error.restype = C.c_char_p
    L.arap_abi_version.restype = C.c_int
    _lib = L
    return L


def default_options():
    o = Options()
    lib().arap_default_options(C.byref(o))
    return o


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """A numpy array backed by page-locked host memory (arap_host_alloc); keep the object alive."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        self._ptr = C.c_void_p()
        rc = lib().arap_host_alloc(nbytes, C.byref(self._ptr))
        if rc != ARAP_OK:
            raise ArapError(rc, "arap_host_alloc failed")
        buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        if getattr(self, "_ptr", None) and self._ptr.value:
            self.array = None
            lib().arap_host_free(self._ptr)
            self._ptr = C.c_void_p()


class ArapError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"arap_b200 error {code}: {msg}")
        self.code = code


class AsRigidAsPossibleDeformation:
    """Host-side mirror of the reference class deform::AsRigidAsPossibleDeformation<MeshType,
    PrecisionType> (reference inc/deform/arap.h:49-138) on top of the C ABI.

    The "mesh" is a (V,3) float32/float64 numpy array (the reference's `Mesh&`; float32 is the
    OpenMesh default scalar) plus an (F,3) int32 face array. Same call protocol and semantics:
    `setConstraint(idx, loc)`, `deform(n) -> bool` (re-reads the rest pose from the mesh when a
    constraint changed, writes p' back into the mesh array).
    """

    def __init__(self, positions, faces, precision=None, **options):
        if positions.dtype not in (np.float32, np.float64) or not positions.flags.c_contiguous:
            raise TypeError("positions must be a C-contiguous float32/float64 (V,3) array")
        self.mesh = positions
        self.faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        self.real = np.dtype(precision if precision is not None else positions.dtype)
        self.nV = positions.shape[0]
        opt = default_options()
        for k, v in options.items():
            setattr(opt, k, v)
        self._h = C.c_void_p()
        rc = lib().arap_create(_ptr(self.faces), self.faces.shape[0], self.nV, self.real.itemsize, C.byref(opt),
                               C.byref(self._h))
        if rc != ARAP_OK:
            self._h = None
            raise ArapError(rc, lib().arap_create_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().arap_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise ArapError(rc, lib().arap_last_error(self._h).decode())
        return rc

    # -- reference API ------------------------------------------------------------------------
    def setConstraint(self, vidx, loc):
        self.setConstraints(np.array([vidx], np.int32), np.asarray(loc, np.float64).reshape(1, 3))

    def setConstraints(self, indices, locations):
        """Batched setConstraint: n indices, (n,3) locations (float32 or float64)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        loc = np.ascontiguousarray(locations)
        if loc.dtype not in (np.float32, np.float64):
            loc = loc.astype(np.float64)
        self._check(lib().arap_set_constraints(self._h, idx.size, _ptr(idx), _ptr(loc), loc.dtype.itemsize))


    def setRigidConstraints(self, indices, rest_points, transform):
        """setConstraint(indices[k], transform @ rest_points[k]) in one call (DeformationUtil::updateConstraints,
        reference deformation_util.h:48-57); transform: 4x4."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        pts = np.ascontiguousarray(rest_points, dtype=np.float64).reshape(-1, 3)
        T = np.ascontiguousarray(transform, dtype=np.float64).reshape(4, 4)
        assert pts.shape[0] == idx.size
        self._check(lib().arap_set_rigid_constraints(self._h, idx.size, _ptr(idx), _ptr(pts), 8, _ptr(T)))
    def deform(self, numberOfIterations):
        """-> bool, like the reference (False only when the linear system is unusable)."""
        rc = lib().arap_deform(self._h, _ptr(self.mesh), self.mesh.dtype.itemsize, int(numberOfIterations))
        if rc in (ARAP_ERR_INVALID, ARAP_ERR_CUDA, ARAP_ERR_ALLOC):
            raise ArapError(rc, lib().arap_last_error(self._h).decode())
        return rc in (ARAP_OK, ARAP_UNCONSTRAINED)     # ARAP_NOT_CONVERGED / ARAP_ERR_SOLVER: the reference's `false`

    def deform_async(self, mesh, numberOfIterations):
        """Pipelined deform (arap_deform_async): iterations + write-back into `mesh` (a page-locked array of the handle's shape) are
        enqueued, nothing is waited for; at most two frames in flight, each in its own buffer. deform_wait() completes the oldest."""
        assert mesh.flags["C_CONTIGUOUS"] and mesh.shape == self.mesh.shape
        return self._check(lib().arap_deform_async(self._h, _ptr(mesh), mesh.dtype.itemsize, int(numberOfIterations)))

    def deform_wait(self):
        rc = lib().arap_deform_wait(self._h)
        if rc in (ARAP_ERR_INVALID, ARAP_ERR_CUDA, ARAP_ERR_ALLOC):
            raise ArapError(rc, lib().arap_last_error(self._h).decode())
        return rc in (ARAP_OK, ARAP_UNCONSTRAINED)

    # -- split protocol (what deform() is made of) ----------------------------------------------
    @property
    def dirty(self):
        return bool(lib().arap_is_dirty(self._h))

    def prepare(self, rest=None):
        rest = self.mesh if rest is None else np.ascontiguousarray(rest)
        return self._check(lib().arap_prepare(self._h, _ptr(rest), rest.dtype.itemsize))

    def iterate(self, n):
        return self._check(lib().arap_iterate(self._h, int(n)))

    def positions(self, dtype=None):
        out = np.zeros((self.nV, 3), dtype or self.real)
        self._check(lib().arap_get_positions(self._h, _ptr(out), out.dtype.itemsize))
        return out

    # -- inspection -----------------------------------------------------------------------------
    def cotanWeights(self):
        """(rowptr, colidx, values) of _edgeWeights -- reference tests/accessor.h:16-22."""
        nnz = C.c_int32()
        self._check(lib().arap_get_csr_nnz(self._h, C.byref(nnz)))
        rp = np.zeros(self.nV + 1, np.int32)
        ci = np.zeros(nnz.value, np.int32)
        w = np.zeros(nnz.value, self.real)
        self._check(lib().arap_get_csr(self._h, _ptr(rp), _ptr(ci), _ptr(w)))
        return rp, ci, w

    def freeIdxMap(self):
        out = np.zeros(self.nV, np.int32)
        n = C.c_int32()
        self._check(lib().arap_get_free_map(self._h, _ptr(out), C.byref(n)))
        return out, n.value

    def rotations(self):
        out = np.zeros((self.nV, 3, 3), self.real)
        self._check(lib().arap_get_rotations(self._h, _ptr(out)))
        return out

    def rhs(self):
        """_b of the reference (arap.h:393-414) for the current rotations: (n_free, 3) float64 in free-index order."""
        _, n_free = self.freeIdxMap()
        out = np.zeros((n_free, 3), np.float64)
        self._check(lib().arap_get_rhs(self._h, _ptr(out)))
        return out

    def render_buffers(self, normals=True, device_pointers=None):
        """Viewer interop (reference examples/osg_viewer.cpp:45-72): float32 positions and vertex normals of the current pose.
        device_pointers = (positions_ptr, normals_ptr or 0): write into device memory of the caller (a mapped VBO) instead."""
        if device_pointers is not None:
            self._check(lib().arap_get_render_buffers(self._h, C.c_void_p(device_pointers[0]), C.c_void_p(device_pointers[1] or None), 1))
            return None
        pos = np.zeros((self.nV, 3), np.float32)
        nrm = np.zeros((self.nV, 3), np.float32) if normals else None
        self._check(lib().arap_get_render_buffers(self._h, _ptr(pos), _ptr(nrm) if normals else None, 0))
        return pos, nrm

    def energy(self):
        e = C.c_double()
        self._check(lib().arap_energy(self._h, C.byref(e)))
        return e.value

    def solver_stats(self):
        s = SolverStats()
        self._check(lib().arap_get_solver_stats(self._h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in SolverStats._fields_ if f[0] != "reserved"}

    # -- measurement ----------------------------------------------------------------------------
    def profile_enable(self, on=True):
        self._check(lib().arap_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(lib().arap_profile_reset(self._h))

    def profile(self):
        p = Profile()
        self._check(lib().arap_profile_get(self._h, C.byref(p)))
        out = {}
        for k in range(K_COUNT_MAX):
            name = lib().arap_kernel_name(k).decode()
            if name and p.launches[k]:
                out[name] = {"launches": int(p.launches[k]), "ms": float(p.milliseconds[k])}
        return out

    def timer_start(self):
        self._check(lib().arap_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(lib().arap_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._check(lib().arap_synchronize(self._h))


class TrajectorySE3:
    """deform::TrajectorySE3 (reference inc/deform/trajectory.h:33-83) over the C ABI's arap_trajectory_*: key poses in,
    smooth pose curve out (cubic B-spline through the se(3) logs). Poses are 4x4 numpy arrays."""

    def __init__(self):
        self._t = C.c_void_p()
        rc = lib().arap_trajectory_create(C.byref(self._t))
        if rc != ARAP_OK:
            self._t = None
            raise ArapError(rc, "arap_trajectory_create failed")

    def close(self):
        if getattr(self, "_t", None):
            lib().arap_trajectory_destroy(self._t)
            self._t = None

    __del__ = close

    def addKeyPose(self, transform):
        T = np.ascontiguousarray(transform, dtype=np.float64).reshape(4, 4)
        rc = lib().arap_trajectory_add_key_pose(self._t, _ptr(T))
        if rc != ARAP_OK:
            raise ArapError(rc, "arap_trajectory_add_key_pose failed")
        return T.copy()

    def sample(self, times):
        """(n, 4, 4) poses at `times` (each in [0,1])."""
        u = np.ascontiguousarray(np.atleast_1d(times), dtype=np.float64)
        out = np.zeros((u.size, 4, 4))
        rc = lib().arap_trajectory_evaluate(self._t, u.size, _ptr(u), _ptr(out))
        if rc != ARAP_OK:
            raise ArapError(rc, lib().arap_create_error().decode())
        return out

    def __call__(self, time):
        return self.sample([time])[0]


def rigid_conjugate(origin, pose):
    """origin @ pose @ origin^-1 (isometry inverse), reference deformation_util.h:38,51."""
    a = np.ascontiguousarray(origin, dtype=np.float64).reshape(4, 4)
    b = np.ascontiguousarray(pose, dtype=np.float64).reshape(4, 4)
    out = np.zeros((4, 4))
    rc = lib().arap_rigid_conjugate(_ptr(a), _ptr(b), _ptr(out))
    if rc != ARAP_OK:
        raise ArapError(rc, "arap_rigid_conjugate failed")
    return out


class DeformationUtil:
    """deform::DeformationUtil (reference inc/deform/deformation_util.h:19-64): remembers where the handles are at
    construction; updateConstraints(t, arap) pins every handle at origin @ t @ origin^-1 @ p0 in ONE call. `arap` may
    be an AsRigidAsPossibleDeformation (t: 4x4) or a BatchDeformation (t: (K,4,4), one transform per member)."""

    def __init__(self, mesh_positions, handles, origin=None):
        self.handles = np.ascontiguousarray(handles, dtype=np.int32)
        self.points = np.array(np.asarray(mesh_positions, dtype=np.float64)[self.handles])
        self.origin = np.eye(4) if origin is None else np.array(origin, dtype=np.float64).reshape(4, 4)

    def updateConstraints(self, t, arap):
        t = np.asarray(t, dtype=np.float64)
        if t.ndim == 2:
            arap.setRigidConstraints(self.handles, self.points, rigid_conjugate(self.origin, t))
        else:
            arap.setRigidConstraints(self.handles, self.points, np.stack([rigid_conjugate(self.origin, x) for x in t]))


class BatchDeformation:
    """K independent deformations of one mesh advanced together (C ABI arap_batch_*): same topology, rest pose and
    constrained vertex set, per-member targets -- BASELINE.json configs[3] (one member per trajectory key frame)."""

    def __init__(self, rest_positions, faces, batch_size, precision=np.float64, **options):
        self.rest = np.ascontiguousarray(rest_positions, dtype=np.float64)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        self.real = np.dtype(precision)
        self.nV, self.K = self.rest.shape[0], int(batch_size)
        opt = default_options()
        for k, v in options.items():
            setattr(opt, k, v)
        self._b = C.c_void_p()
        rc = lib().arap_batch_create(_ptr(self.faces), self.faces.shape[0], self.nV, self.K, self.real.itemsize, C.byref(opt),
                                     C.byref(self._b))
        if rc != ARAP_OK:
            self._b = None
            raise ArapError(rc, lib().arap_create_error().decode())
        self._h = C.c_void_p(lib().arap_batch_handle(self._b))

    def close(self):
        if getattr(self, "_b", None):
            lib().arap_batch_destroy(self._b)
            self._b = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise ArapError(rc, lib().arap_last_error(self._h).decode())
        return rc

    def setConstraints(self, indices, targets):
        """indices: (n,), targets: (K, n, 3)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        tgt = np.ascontiguousarray(targets, dtype=np.float64)
        assert tgt.shape == (self.K, idx.size, 3)
        self._check(lib().arap_batch_set_constraints(self._b, idx.size, _ptr(idx), _ptr(tgt), 8))

    def setRigidConstraints(self, indices, rest_points, transforms):
        """Member m: setConstraint(indices[k], transforms[m] @ rest_points[k]); transforms: (K, 4, 4). The targets are
        computed on the device (arap_batch_set_rigid_constraints)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        pts = np.ascontiguousarray(rest_points, dtype=np.float64).reshape(-1, 3)
        T = np.ascontiguousarray(transforms, dtype=np.float64)
        assert pts.shape[0] == idx.size and T.shape == (self.K, 4, 4)
        self._check(lib().arap_batch_set_rigid_constraints(self._b, idx.size, _ptr(idx), _ptr(pts), 8, _ptr(T)))

    def prepare(self):
        return self._check(lib().arap_batch_prepare(self._b, _ptr(self.rest), 8))

    def iterate(self, n):
        return self._check(lib().arap_batch_iterate(self._b, int(n)))

    def positions(self, dtype=np.float64):
        out = np.zeros((self.K, self.nV, 3), dtype)
        self._check(lib().arap_batch_get_positions(self._b, _ptr(out), out.dtype.itemsize))
        return out

    def total_energy(self):
        e = C.c_double()
        self._check(lib().arap_energy(self._h, C.byref(e)))
        return e.value

    def solver_stats(self):
        s = SolverStats()
        self._check(lib().arap_get_solver_stats(self._h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in SolverStats._fields_}

    def profile_enable(self, on=True):
        self._check(lib().arap_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(lib().arap_profile_reset(self._h))

    def profile(self):
        p = Profile()
        self._check(lib().arap_profile_get(self._h, C.byref(p)))
        return {lib().arap_kernel_name(k).decode(): {"launches": int(p.launches[k]), "ms": float(p.milliseconds[k])}
                for k in range(K_COUNT_MAX) if p.launches[k] and lib().arap_kernel_name(k)}

    def timer_start(self):
        self._check(lib().arap_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(lib().arap_timer_stop(self._h, C.byref(ms)))
        return ms.value


def comm_unique_id():
    """128-byte NCCL unique id (call on rank 0, broadcast to the other ranks, e.g. with torch.distributed)."""
    buf = np.zeros(128, np.uint8)
    rc = lib().arap_comm_unique_id(_ptr(buf), 128)
    if rc != ARAP_OK:
        raise ArapError(rc, lib().arap_create_error().decode())
    return buf


class PartitionedDeformation:
    """One rank's share of a mesh partitioned over several GPUs (C ABI arap_attach_partition).

    positions/faces are the GLOBAL mesh (every rank holds it on the host); `owner` maps vertex -> rank
    (mesh_deform_b200.partition.strip_owner). transport: TRANSPORT_NCCL (one process per GPU, comm_id = the broadcast
    128-byte unique id) or TRANSPORT_IN_PROCESS (several partitions on one GPU driven by one host thread each,
    comm_id = an int group key) -- the latter is how the single-GPU test tier runs the very same solver code path.
    """

    def __init__(self, positions, faces, owner, rank, world, transport, comm_id, precision=np.float64, global_multigrid=True,
                 **options):
        from . import partition as part_mod
        self.part = part_mod.build_local_part(faces, owner, rank, world)
        self.global_rest = np.ascontiguousarray(positions, dtype=np.float64)
        self.global_faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        self.owner = np.ascontiguousarray(owner, dtype=np.int32)
        self.global_multigrid = bool(global_multigrid)
        self._constrained = np.zeros(0, np.int32)
        self.local_mesh = np.ascontiguousarray(self.global_rest[self.part.local_to_global], dtype=np.dtype(precision))
        self.arap = AsRigidAsPossibleDeformation(self.local_mesh, self.part.faces, precision, **options)
        pl = self.part
        self._keep = [np.ascontiguousarray(a, dtype=np.int32) for a in (pl.neighbor_rank, pl.send_offset, pl.send_index, pl.recv_offset)]
        plan = PartitionPlan(pl.n_owned, int(pl.neighbor_rank.size), *[a.ctypes.data for a in self._keep])
        if transport in (TRANSPORT_NCCL, TRANSPORT_PEER):
            ident = np.ascontiguousarray(comm_id, dtype=np.uint8)
        else:
            ident = np.array([int(comm_id)], np.int32)
        self.arap._check(lib().arap_attach_partition(self.arap._h, C.byref(plan), rank, world, transport, _ptr(ident), ident.nbytes))

    def setConstraints(self, global_indices, targets):
        """Every rank passes the SAME global constraint set; each keeps the part it holds (owned and halo)."""
        from . import partition as part_mod
        self._constrained = np.union1d(self._constrained, np.asarray(global_indices, dtype=np.int32)).astype(np.int32)
        idx, tgt = part_mod.local_constraints(self.part, global_indices, targets)
        if idx.size:
            self.arap.setConstraints(idx, tgt)

    def prepare(self):
        if self.global_multigrid:
            # the global mesh lets every rank build the same whole-mesh multigrid hierarchy (arap_partition_set_global_mesh)
            l2g = np.ascontiguousarray(self.part.local_to_global, dtype=np.int32)
            con = np.ascontiguousarray(self._constrained, dtype=np.int32)
            g = GlobalMesh(self.global_rest.shape[0], self.global_faces.shape[0], self.global_faces.ctypes.data,
                           self.global_rest.ctypes.data, 8, self.owner.ctypes.data, l2g.ctypes.data, con.size, con.ctypes.data)
            self.arap._check(lib().arap_partition_set_global_mesh(self.arap._h, C.byref(g)))
        return self.arap.prepare()

    def iterate(self, n):
        return self.arap.iterate(n)

    def owned_positions(self, dtype=np.float64):
        """(global vertex ids, positions) of the vertices this rank owns."""
        return self.part.owned_global, self.arap.positions(dtype)[:self.part.n_owned]

    def local_energy(self):
        return self.arap.energy()

    def profile_enable(self, on=True):
        self.arap.profile_enable(on)

    def profile_reset(self):
        self.arap.profile_reset()

    def profile(self):
        return self.arap.profile()

    def comm_benchmark(self, rounds=200):
        """(microseconds per halo exchange, per all-reduce) on this rank's stream; collective."""
        a, b = C.c_double(), C.c_double()
        self.arap._check(lib().arap_partition_comm_benchmark(self.arap._h, int(rounds), C.byref(a), C.byref(b)))
        return a.value, b.value

    def solver_stats(self):
        return self.arap.solver_stats()


def run_partitions_in_process(workers):
    """Run one callable per partition concurrently (one host thread each; ctypes releases the GIL inside the C ABI).
    Exceptions are re-raised. Used with TRANSPORT_IN_PROCESS."""
    import threading
    errors = [None] * len(workers)

    def wrap(k, fn):
        try:
            fn()
        except BaseException as exc:      # noqa: BLE001
            errors[k] = exc
    threads = [threading.Thread(target=wrap, args=(k, fn)) for k, fn in enumerate(workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
