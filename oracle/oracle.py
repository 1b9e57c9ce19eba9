"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front end of the CPU oracle (oracle/arap_oracle.c, oracle/trajectory_oracle.c), the
restatement of reference inc/deform/arap.h and inc/deform/trajectory.h. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product (mesh_deform_b200/, inc/, include/) never does.

`ArapOracle` mirrors the reference class `deform::AsRigidAsPossibleDeformation<MeshType,
PrecisionType>` (arap.h:49-466): ctor(mesh) / setConstraint(idx, loc) / deform(n) -> bool, where the
"mesh" is a (V,3) positions array (float32 = the OpenMesh default scalar, or float64) plus an
(F,3) int32 face array; deform() writes the result back into that array like arap.h:133-135.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

TIMER_NAMES = ("geometry", "weights", "assembly", "factor", "local", "rhs", "solve", "writeback")


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("arap_oracle.c", "sparse_ldlt.c", "sparse_ldlt.h", "trajectory_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs))
    if force or stale:
        try:
            subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
        except (OSError, subprocess.CalledProcessError):
            if not os.path.exists(_LIB_PATH):
                raise
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    for sfx in ("_f64", "_f32"):
        g = lambda n: getattr(L, n + sfx)
        g("oracle_create").restype = vp
        g("oracle_create").argtypes = [i, i, vp]
        g("oracle_destroy").argtypes = [vp]
        g("oracle_set_constraint").argtypes = [vp, i, vp]
        g("oracle_deform").restype = i
        g("oracle_deform").argtypes = [vp, vp, i, i]
        g("oracle_energy").restype = d
        g("oracle_energy").argtypes = [vp]
        for n in ("oracle_nnz", "oracle_nfree", "oracle_dirty", "oracle_L_nnz"):
            g(n).restype = i
            g(n).argtypes = [vp]
        g("oracle_factor_nnz").restype = C.c_longlong
        g("oracle_factor_nnz").argtypes = [vp]
        g("oracle_get_csr").argtypes = [vp, vp, vp, vp]
        g("oracle_get_L").argtypes = [vp, vp, vp, vp]
        for n in ("oracle_get_free_map", "oracle_get_rotations", "oracle_get_rest", "oracle_get_positions",
                  "oracle_get_bfixed", "oracle_get_b", "oracle_get_timers"):
            g(n).argtypes = [vp, vp]
        g("oracle_reset_timers").argtypes = [vp]
        g("oracle_rotation_from_covariance").argtypes = [vp, vp]
        g("oracle_svd3").argtypes = [vp, vp, vp, vp]
    L.traj_se3_exp.argtypes = [vp, vp]
    L.traj_se3_log.argtypes = [vp, vp]
    L.traj_fit.restype = i
    L.traj_fit.argtypes = [i, vp, vp, vp, vp]
    L.traj_eval.argtypes = [i, vp, vp, d, vp]
    L.traj_handle_targets.argtypes = [vp, vp, i, vp, vp]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class ArapOracle:
    """Restatement of deform::AsRigidAsPossibleDeformation (arap.h:49-466)."""

    def __init__(self, positions, faces, precision=np.float64):
        """positions: (V,3) float32/float64 array, MUTATED by deform() (the reference's Mesh&);
        faces: (F,3) int; precision: the reference's PrecisionType (np.float32 / np.float64)."""
        assert positions.dtype in (np.float32, np.float64) and positions.flags.c_contiguous
        self.mesh = positions
        self.faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        self.real = np.dtype(precision)
        self._sfx = "_f64" if self.real == np.float64 else "_f32"
        self.nV, self.nF = positions.shape[0], self.faces.shape[0]
        self._h = self._f("oracle_create")(self.nV, self.nF, _ptr(self.faces))

    def _f(self, name):
        return getattr(lib(), name + self._sfx)

    def __del__(self):
        if getattr(self, "_h", None):
            self._f("oracle_destroy")(self._h)
            self._h = None

    def setConstraint(self, vidx, loc):
        loc = np.ascontiguousarray(loc, dtype=np.float64)
        self._f("oracle_set_constraint")(self._h, int(vidx), _ptr(loc))

    def deform(self, numberOfIterations):
        return bool(self._f("oracle_deform")(self._h, _ptr(self.mesh), int(self.mesh.dtype == np.float32),
                                             int(numberOfIterations)))

    # -- inspection ---------------------------------------------------------------------------
    def energy(self):
        return float(self._f("oracle_energy")(self._h))

    @property
    def nFree(self):
        return self._f("oracle_nfree")(self._h)

    @property
    def dirty(self):
        return bool(self._f("oracle_dirty")(self._h))

    def factor_nnz(self):
        return int(self._f("oracle_factor_nnz")(self._h))

    def cotanWeights(self):
        """(rowptr, colidx, val) of _edgeWeights -- what tests/accessor.h:16-22 exposes."""
        nnz = self._f("oracle_nnz")(self._h)
        rp = np.zeros(self.nV + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, self.real)
        self._f("oracle_get_csr")(self._h, _ptr(rp), _ptr(ci), _ptr(v))
        return rp, ci, v

    def systemMatrix(self):
        nnz = self._f("oracle_L_nnz")(self._h)
        rp = np.zeros(self.nFree + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz, self.real)
        self._f("oracle_get_L")(self._h, _ptr(rp), _ptr(ci), _ptr(v))
        return rp, ci, v

    def _get(self, name, shape, dtype=None):
        out = np.zeros(shape, dtype or self.real)
        self._f(name)(self._h, _ptr(out))
        return out

    def freeIdxMap(self):
        return self._get("oracle_get_free_map", self.nV, np.int32)

    def rotations(self):
        """(V,3,3), R[v][a][b] = R_v(a,b)."""
        return self._get("oracle_get_rotations", (self.nV, 3, 3))

    def rest(self):
        return self._get("oracle_get_rest", (self.nV, 3))

    def positions(self):
        return self._get("oracle_get_positions", (self.nV, 3))

    def bFixed(self):
        return self._get("oracle_get_bfixed", (self.nFree, 3))

    def b(self):
        return self._get("oracle_get_b", (self.nFree, 3))

    def timers(self):
        t = self._get("oracle_get_timers", len(TIMER_NAMES), np.float64)
        return dict(zip(TIMER_NAMES, t.tolist()))

    def reset_timers(self):
        self._f("oracle_reset_timers")(self._h)


def rotation_from_covariance(cov, precision=np.float64):
    """R = V diag(1,1,det(V U^T)) U^T of one 3x3 covariance (arap.h:376-382)."""
    real = np.dtype(precision)
    sfx = "_f64" if real == np.float64 else "_f32"
    cov = np.ascontiguousarray(cov, dtype=real)
    out = np.zeros((3, 3), real)
    getattr(lib(), "oracle_rotation_from_covariance" + sfx)(_ptr(cov), _ptr(out))
    return out


def svd3(m, precision=np.float64):
    real = np.dtype(precision)
    sfx = "_f64" if real == np.float64 else "_f32"
    m = np.ascontiguousarray(m, dtype=real)
    u, s, v = np.zeros((3, 3), real), np.zeros(3, real), np.zeros((3, 3), real)
    getattr(lib(), "oracle_svd3" + sfx)(_ptr(m), _ptr(u), _ptr(s), _ptr(v))
    return u, s, v


# -- trajectory front end (trajectory.h, deformation_util.h) ---------------------------------------

def se3_exp(xi):
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    T = np.zeros((4, 4))
    lib().traj_se3_exp(_ptr(xi), _ptr(T))
    return T


def se3_log(T):
    T = np.ascontiguousarray(T, dtype=np.float64)
    xi = np.zeros(6)
    lib().traj_se3_log(_ptr(T), _ptr(xi))
    return xi


class TrajectorySE3Oracle:
    """Restatement of deform::TrajectorySE3 (trajectory.h:33-83)."""

    def __init__(self):
        self._poses = []
        self._dirty = False

    def addKeyPose(self, T):
        self._poses.append(np.array(T, dtype=np.float64).reshape(4, 4))
        self._dirty = True
        return T

    def __call__(self, time):
        n = len(self._poses)
        if self._dirty:
            poses = np.ascontiguousarray(np.stack(self._poses))
            self._ctrl = np.zeros((n, 6))
            self._knots = np.zeros(n + 4)
            self._params = np.zeros(n)
            if not lib().traj_fit(n, _ptr(poses), _ptr(self._ctrl), _ptr(self._knots), _ptr(self._params)):
                raise ValueError("spline interpolation needs at least 4 key poses")
            self._dirty = False
        T = np.zeros((4, 4))
        lib().traj_eval(n, _ptr(self._ctrl), _ptr(self._knots), float(time), _ptr(T))
        return T


def handle_targets(origin, t, points):
    """DeformationUtil::updateConstraints arithmetic (deformation_util.h:48-57)."""
    origin = np.ascontiguousarray(origin, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    out = np.zeros_like(points)
    lib().traj_handle_targets(_ptr(origin), _ptr(t), points.shape[0], _ptr(points), _ptr(out))
    return out
