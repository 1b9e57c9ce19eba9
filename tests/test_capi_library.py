"""CPU tests: the C-ABI shared library loads and exports every symbol include/arap_b200.h declares
(no compute calls without a GPU), and the host math shim matches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from mesh_deform_b200 import capi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "arap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(arap_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mesh_deform_b200 import capi
    lib = capi.lib()                      # raises EngineMissingError if the .so was not built
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libarap_b200.so does not export {name}"
    assert set(declared) == set(capi.EXPORTED_SYMBOLS)
    assert lib.arap_abi_version() == 1


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.join(ROOT, "mesh_deform_b200", "libarap_b200.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_default_options_and_argument_validation():
    from mesh_deform_b200 import capi
    o = capi.default_options()
    assert o.struct_size == C.sizeof(capi.Options) and o.device == -1 and o.cg_tolerance == 0
    lib = capi.lib()
    h = C.c_void_p()
    faces = np.array([[0, 1, 5]], np.int32)     # vertex 5 out of range for V = 3
    rc = lib.arap_create(faces.ctypes.data_as(C.c_void_p), 1, 3, 8, None, C.byref(h))
    assert rc == capi.ARAP_ERR_INVALID and not h.value
    rc = lib.arap_create(faces.ctypes.data_as(C.c_void_p), 1, 6, 2, None, C.byref(h))   # bad precision
    assert rc == capi.ARAP_ERR_INVALID
    assert lib.arap_is_dirty(None) == 1
    assert lib.arap_iterate(None, 1) == capi.ARAP_ERR_INVALID


def test_no_cpu_fallback_in_product():
    """The product path must never touch the oracle."""
    for base in ("mesh_deform_b200", "inc", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".h", ".cu", ".cuh", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle" not in text.lower() or f in (), f"{base}/{f} mentions the oracle"


# ---- the engine's per-element arithmetic, compiled for the host (tests/cpp/math_host_shim.cpp) -----
@pytest.fixture(scope="module")
def host_math():
    src = os.path.join(ROOT, "tests", "cpp", "math_host_shim.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "libarap_math_host.so")
    hdr = os.path.join(ROOT, "mesh_deform_b200", "csrc", "arap_math.cuh")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("dt,sfx,tol", [(np.float64, "f64", 1e-12), (np.float32, "f32", 5e-6)])
def test_device_rotation_math_matches_oracle(host_math, dt, sfx, tol):
    rng = np.random.default_rng(0)
    for trial in range(3000):
        M = rng.standard_normal((3, 3))
        kind = trial % 5
        u, s, vt = np.linalg.svd(M)
        if kind == 1:
            M = M @ np.diag([1, 1e-2, 1e-3])
        elif kind == 2:
            M = u @ np.diag([s[0], s[1], 0]) @ vt           # rank 2 (planar one-ring)
        elif kind == 3:
            M = u @ np.diag([s[0], s[1], -s[2]]) @ vt       # needs the det flip
        elif kind == 4:
            M = M * 1e-9
        M = np.ascontiguousarray(M, dtype=dt)
        sv = np.linalg.svd(M.astype(np.float64), compute_uv=False)
        gap = (sv[1] + np.sign(np.linalg.det(M.astype(np.float64))) * sv[2]) / sv[0]
        if gap < (1e-6 if dt == np.float64 else 3e-2):
            continue
        want = O.rotation_from_covariance(M.astype(np.float64), np.float64)
        q = np.zeros(4, dt)
        getattr(host_math, "math_quat_" + sfx)(_p(M), _p(q))
        R = np.zeros((3, 3), dt)
        getattr(host_math, "math_quat_to_matrix_" + sfx)(_p(q), _p(R))
        assert abs(np.linalg.norm(q.astype(np.float64)) - 1) < (1e-14 if dt == np.float64 else 1e-6)
        assert np.abs(R - want).max() * gap < tol


def test_device_rotation_degenerate_inputs_stay_rotations(host_math):
    for M in (np.zeros((3, 3)), np.outer([1.0, 2, 3], [0.5, -1, 2]), np.diag([1.0, 0, 0])):
        R = np.zeros((3, 3))
        host_math.math_rotation_f64(_p(np.ascontiguousarray(M)), _p(R))
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12
    R = np.zeros((3, 3))
    host_math.math_rotation_f64(_p(np.zeros((3, 3))), _p(R))
    assert np.array_equal(R, np.eye(3))       # JacobiSVD of 0 gives U = V = I


@pytest.mark.parametrize("dt,sfx", [(np.float64, "f64"), (np.float32, "f32")])
def test_device_cotan_math_is_bit_exact(host_math, meshes, dt, sfx):
    P, F = meshes["sphere"]
    Pp = P.astype(dt)
    a = O.ArapOracle(Pp.copy(), F, dt)
    a.deform(0)
    rp, ci, w = a.cotanWeights()
    lookup = {}
    for r in range(len(P)):
        for k in range(rp[r], rp[r + 1]):
            lookup[(r, ci[k])] = w[k]
    acc = {}
    out = np.zeros(3, dt)
    for f in F:
        v = [np.ascontiguousarray(Pp[i]) for i in f]
        getattr(host_math, "math_cotan_" + sfx)(_p(v[0]), _p(v[1]), _p(v[2]), _p(out))
        for k, (x, y) in enumerate(((f[0], f[1]), (f[1], f[2]), (f[2], f[0]))):
            acc.setdefault((min(x, y), max(x, y)), []).append(dt(out[k]))
    for (x, y), vals in acc.items():
        s = vals[0]
        for t in vals[1:]:
            s = dt(s + t)
        assert lookup[(x, y)] == s and lookup[(y, x)] == s


def test_trajectory_abi_matches_reference_pins_and_oracle(trajectory_golden):
    """arap_trajectory_* (host arithmetic, no device needed): the endpoint pins of reference tests/test_trajectory.cpp:
    17-39, the committed oracle samples in between, and a 7-key-pose curve (interior knots) against the oracle."""
    poses = trajectory_golden["key_poses"]
    path = capi.TrajectorySE3()
    for p_ in poses:
        back = path.addKeyPose(p_)
        assert np.array_equal(back, p_)                                   # addKeyPose returns its argument (trajectory.h:58)
    for u, want in ((0.0, poses[0]), (1.0, poses[3])):
        got = path(u)
        assert np.linalg.norm(got - want) <= 1e-3 * min(np.linalg.norm(got), np.linalg.norm(want))
    got = path.sample(trajectory_golden["u"])
    assert np.abs(got - trajectory_golden["samples"]).max() < 1e-12
    rng = np.random.default_rng(5)
    a, b = capi.TrajectorySE3(), O.TrajectorySE3Oracle()
    for _ in range(7):
        xi = rng.standard_normal(6) * np.array([1, 1, 1, 0.6, 0.6, 0.6])
        T = O.se3_exp(xi)
        a.addKeyPose(T)
        b.addKeyPose(T)
    us = np.linspace(0, 1, 41)
    want = np.stack([b(float(u)) for u in us])
    assert np.abs(a.sample(us) - want).max() < 1e-11


def test_trajectory_abi_rejects_too_few_key_poses():
    path = capi.TrajectorySE3()
    for _ in range(3):
        path.addKeyPose(np.eye(4))
    with pytest.raises(capi.ArapError):
        path(0.5)


def test_rigid_conjugate_matches_oracle():
    """arap_rigid_conjugate = origin * t * origin^-1 of reference deformation_util.h:51."""
    rng = np.random.default_rng(11)
    origin = O.se3_exp(rng.standard_normal(6))
    t = O.se3_exp(rng.standard_normal(6) * 0.5)
    pts = rng.standard_normal((9, 3))
    M = capi.rigid_conjugate(origin, t)
    got = pts @ M[:3, :3].T + M[:3, 3]
    assert np.abs(got - O.handle_targets(origin, t, pts)).max() < 1e-13


def test_header_is_plain_c(tmp_path):
    """include/arap_b200.h is the drop-in boundary: a C header (C99, pedantic), usable from any FFI, no C++ or torch types."""
    src = tmp_path / "abi.c"
    src.write_text('#include "arap_b200.h"\n'
                   'int main(void) { arap_options o; arap_global_mesh g; arap_partition_plan p; arap_solver_stats s; arap_profile q;\n'
                   '  (void)g; (void)p; (void)s; (void)q; arap_default_options(&o); return arap_abi_version() == ARAP_B200_ABI_VERSION ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
    # and it links against the library without any C++ runtime symbols leaking into the interface
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", os.path.join(ROOT, "mesh_deform_b200"),
                           "-larap_b200", "-Wl,-rpath," + os.path.join(ROOT, "mesh_deform_b200")])
    assert subprocess.run([str(exe)]).returncode == 0
