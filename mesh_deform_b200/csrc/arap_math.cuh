// arap_math.cuh -- per-element arithmetic of the ARAP hot path, shared by all kernels.
//
// Everything here is `__host__ __device__` so that the exact same arithmetic can be unit-tested
// on a GPU-less machine (tests/cpp/test_math_host.cpp compiles this header with g++). The library
// itself only ever calls these functions from device code; there is no CPU execution path.
//
// Reference lines restated:
//   cotan_half_weights()      reference inc/deform/arap.h:199-218 (+ the 0.5 of :225-232)
//   rotation_from_covariance() reference inc/deform/arap.h:376-382 (JacobiSVD + det fix), returned
//                              as a unit quaternion instead of a 3x3 matrix
#pragma once

#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define ARAP_HD __host__ __device__ __forceinline__
#define ARAP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ARAP_HD inline
#define ARAP_HD_NOINLINE inline
#endif

namespace arap {

template <typename S> struct Vec4T;
template <> struct alignas(16) Vec4T<float> { float x, y, z, w; };
template <> struct alignas(32) Vec4T<double> { double x, y, z, w; };

struct alignas(8) Vec3d { double x, y, z; };

// ---- rounding-controlled primitives ----------------------------------------------------------
// The weights must match the reference bit for bit where possible, so the cotan code uses
// explicitly rounded add/mul (no FMA contraction), mirroring a plain C++ build of arap.h.
#if defined(__CUDA_ARCH__)
ARAP_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
ARAP_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
ARAP_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
ARAP_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
ARAP_HD float sqrt_rn(float a) { return __fsqrt_rn(a); }
ARAP_HD double sqrt_rn(double a) { return __dsqrt_rn(a); }
ARAP_HD float div_rn(float a, float b) { return __fdiv_rn(a, b); }
ARAP_HD double div_rn(double a, double b) { return __ddiv_rn(a, b); }
ARAP_HD float rsqrt_fast(float a) { return rsqrtf(a); }
ARAP_HD double rsqrt_fast(double a) { return rsqrt(a); }
#else
ARAP_HD float add_rn(float a, float b) { volatile float r = a + b; return r; }
ARAP_HD float mul_rn(float a, float b) { volatile float r = a * b; return r; }
ARAP_HD double add_rn(double a, double b) { volatile double r = a + b; return r; }
ARAP_HD double mul_rn(double a, double b) { volatile double r = a * b; return r; }
ARAP_HD float sqrt_rn(float a) { return sqrtf(a); }
ARAP_HD double sqrt_rn(double a) { return sqrt(a); }
ARAP_HD float div_rn(float a, float b) { volatile float r = a / b; return r; }
ARAP_HD double div_rn(double a, double b) { volatile double r = a / b; return r; }
ARAP_HD float rsqrt_fast(float a) { return 1.0f / sqrtf(a); }
ARAP_HD double rsqrt_fast(double a) { return 1.0 / sqrt(a); }
#endif

// std::max(a, b): returns a unless a < b, so max_std(c, NaN) == c (matters at arap.h:208).
template <typename S> ARAP_HD S max_std(S a, S b) { return (a < b) ? b : a; }

// ---- cotangent weights of one face (arap.h:199-218) -------------------------------------------
// v0,v1,v2: corner positions. Output: half cotans (cot/2) for edges e0=(v0,v1), e1=(v1,v2), e2=(v2,v0),
// i.e. the triplet values of arap.h:225-232.
template <typename S>
ARAP_HD void cotan_half_weights(const S v0[3], const S v1[3], const S v2[3], S out[3]) {
    // squaredNorm() of a fixed-size 3-vector: Eigen's unrolled reduction is x^2 + (y^2 + z^2)
    S a[3], b[3], c[3];
    for (int d = 0; d < 3; ++d) { a[d] = add_rn(v1[d], -v0[d]); b[d] = add_rn(v2[d], -v1[d]); c[d] = add_rn(v0[d], -v2[d]); }
    S l0 = add_rn(mul_rn(a[0], a[0]), add_rn(mul_rn(a[1], a[1]), mul_rn(a[2], a[2])));
    S l1 = add_rn(mul_rn(b[0], b[0]), add_rn(mul_rn(b[1], b[1]), mul_rn(b[2], b[2])));
    S l2 = add_rn(mul_rn(c[0], c[0]), add_rn(mul_rn(c[1], c[1]), mul_rn(c[2], c[2])));
    l0 = sqrt_rn(max_std(S(1e-8), l0));
    l1 = sqrt_rn(max_std(S(1e-8), l1));
    l2 = sqrt_rn(max_std(S(1e-8), l2));
    const S semip = mul_rn(S(0.5), add_rn(add_rn(l0, l1), l2));
    const S heron = mul_rn(mul_rn(mul_rn(semip, add_rn(semip, -l0)), add_rn(semip, -l1)), add_rn(semip, -l2));
    const S area = max_std(S(1e-8), sqrt_rn(heron));
    const S denom = div_rn(S(1.0), mul_rn(S(4.0), area));
    const S q0 = mul_rn(l0, l0), q1 = mul_rn(l1, l1), q2 = mul_rn(l2, l2);
    S cot0 = mul_rn(add_rn(add_rn(-q0, q1), q2), denom);
    S cot1 = mul_rn(add_rn(add_rn(q0, -q1), q2), denom);
    S cot2 = mul_rn(add_rn(add_rn(q0, q1), -q2), denom);
    cot0 = max_std(S(1e-10), cot0);
    cot1 = max_std(S(1e-10), cot1);
    cot2 = max_std(S(1e-10), cot2);
    out[0] = mul_rn(cot0, S(0.5));
    out[1] = mul_rn(cot1, S(0.5));
    out[2] = mul_rn(cot2, S(0.5));
}

// ---- quaternions (w, x, y, z) ----------------------------------------------------------------
template <typename S>
ARAP_HD void quat_to_matrix(S qw, S qx, S qy, S qz, S r[9]) {
    const S xx = qx * qx, yy = qy * qy, zz = qz * qz;
    const S xy = qx * qy, xz = qx * qz, yz = qy * qz;
    const S wx = qw * qx, wy = qw * qy, wz = qw * qz;
    r[0] = S(1) - S(2) * (yy + zz); r[1] = S(2) * (xy - wz);        r[2] = S(2) * (xz + wy);
    r[3] = S(2) * (xy + wz);        r[4] = S(1) - S(2) * (xx + zz); r[5] = S(2) * (yz - wx);
    r[6] = S(2) * (xz - wy);        r[7] = S(2) * (yz + wx);        r[8] = S(1) - S(2) * (xx + yy);
}

// R(q) v for a unit quaternion without forming the matrix: t = 2 q_v x v ; R v = v + w t + q_v x t (18 multiply-adds instead
// of ~20 for the matrix plus 9 for the product, and no nine live matrix entries -- the right-hand-side kernel rotates one edge
// per neighbour and is short of registers).
template <typename S>
ARAP_HD void quat_rotate(S qw, S qx, S qy, S qz, S vx, S vy, S vz, S &ox, S &oy, S &oz) {
    const S tx = S(2) * (qy * vz - qz * vy), ty = S(2) * (qz * vx - qx * vz), tz = S(2) * (qx * vy - qy * vx);
    ox = vx + qw * tx + (qy * tz - qz * ty);
    oy = vy + qw * ty + (qz * tx - qx * tz);
    oz = vz + qw * tz + (qx * ty - qy * tx);
}

// Proper rotation matrix (row-major) -> unit quaternion, largest-component selection, branch-free.
template <typename S>
ARAP_HD void matrix_to_quat(const S r[9], S q[4]) {
    const S t0 = S(1) + r[0] + r[4] + r[8];
    const S t1 = S(1) + r[0] - r[4] - r[8];
    const S t2 = S(1) - r[0] + r[4] - r[8];
    const S t3 = S(1) - r[0] - r[4] + r[8];
    const S a = r[7] - r[5], b = r[2] - r[6], c = r[3] - r[1];     // 4wx, 4wy, 4wz
    const S d = r[1] + r[3], e = r[2] + r[6], f = r[5] + r[7];     // 4xy, 4xz, 4yz
    S qw = t0, qx = a, qy = b, qz = c, tm = t0;
    if (t1 > tm) { tm = t1; qw = a; qx = t1; qy = d; qz = e; }
    if (t2 > tm) { tm = t2; qw = b; qx = d; qy = t2; qz = f; }
    if (t3 > tm) { tm = t3; qw = c; qx = e; qy = f; qz = t3; }
    // |q|^2 = tm * 4 ... normalise exactly instead of trusting tm (protects against drift in r)
    const S n2 = qw * qw + qx * qx + qy * qy + qz * qz;
    S inv = rsqrt_fast(n2);
    inv = inv * (S(1.5) - S(0.5) * n2 * inv * inv);                 // one Newton step: full precision
    q[0] = qw * inv; q[1] = qx * inv; q[2] = qy * inv; q[3] = qz * inv;
}

// ---- local step: covariance -> rotation (arap.h:376-382) --------------------------------------
template <typename S> struct JacobiSweeps;
template <> struct JacobiSweeps<float> { static constexpr int value = 5; };
template <> struct JacobiSweeps<double> { static constexpr int value = 7; };

// One Hestenes rotation: make columns p and q of `a` orthogonal, accumulate into `v`.
template <typename S>
ARAP_HD void hestenes_rotate(S a[9], S v[9], int p, int q) {
    const S ap0 = a[p], ap1 = a[3 + p], ap2 = a[6 + p];
    const S aq0 = a[q], aq1 = a[3 + q], aq2 = a[6 + q];
    const S alpha = ap0 * ap0 + ap1 * ap1 + ap2 * ap2;
    const S beta = aq0 * aq0 + aq1 * aq1 + aq2 * aq2;
    const S gamma = ap0 * aq0 + ap1 * aq1 + ap2 * aq2;
    // tan of the rotation angle: t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (beta - alpha) / (2 gamma),
    // written without the division by gamma so that gamma == 0 gives t == 0.
    const S h = beta - alpha, g = S(2) * gamma;
    const S hyp = sqrt(h * h + g * g);
    const S den = fabs(h) + hyp;
    S t = (den > S(0)) ? fabs(g) / den : S(0);
    t = ((h < S(0)) != (g < S(0))) ? -t : t;
    const S c = S(1) / sqrt(S(1) + t * t);
    const S s = c * t;
    a[p] = c * ap0 - s * aq0; a[3 + p] = c * ap1 - s * aq1; a[6 + p] = c * ap2 - s * aq2;
    a[q] = s * ap0 + c * aq0; a[3 + q] = s * ap1 + c * aq1; a[6 + q] = s * ap2 + c * aq2;
    const S vp0 = v[p], vp1 = v[3 + p], vp2 = v[6 + p];
    const S vq0 = v[q], vq1 = v[3 + q], vq2 = v[6 + q];
    v[p] = c * vp0 - s * vq0; v[3 + p] = c * vp1 - s * vq1; v[6 + p] = c * vp2 - s * vq2;
    v[q] = s * vp0 + c * vq0; v[3 + q] = s * vp1 + c * vq1; v[6 + q] = s * vp2 + c * vq2;
}

// Swap columns p,q of a and v as a proper 90-degree rotation (col p <- col q, col q <- -col p).
template <typename S>
ARAP_HD void swap_columns_proper(S a[9], S v[9], int p, int q) {
    for (int r = 0; r < 3; ++r) {
        S t = a[3 * r + p]; a[3 * r + p] = a[3 * r + q]; a[3 * r + q] = -t;
        t = v[3 * r + p]; v[3 * r + p] = v[3 * r + q]; v[3 * r + q] = -t;
    }
}

// cov (row-major, cov[3a+b] = sum_j w_ij (p_i-p_j)_a (p'_i-p'_j)_b) -> R = V diag(1,1,det(V U^T)) U^T
// as a row-major matrix. One-sided (Hestenes) Jacobi: cov * V = U * Sigma with V a proper rotation;
// U's third column is taken as u1 x u2, which IS the det-fix of arap.h:380-382.
template <typename S>
ARAP_HD void rotation_matrix_from_covariance(const S cov[9], S rot[9]) {
    S a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    S scale = 0;
    for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(cov[i]));
    if (!(scale > S(0))) {                       // zero covariance (or NaN): JacobiSVD gives U = V = I
        for (int i = 0; i < 9; ++i) rot[i] = v[i];
        return;
    }
    const S inv_scale = S(1) / scale;
    for (int i = 0; i < 9; ++i) a[i] = cov[i] * inv_scale;
#pragma unroll 1
    for (int sweep = 0; sweep < JacobiSweeps<S>::value; ++sweep) {
        hestenes_rotate(a, v, 0, 1);
        hestenes_rotate(a, v, 0, 2);
        hestenes_rotate(a, v, 1, 2);
    }
    // sort columns by norm, descending
    S n0 = a[0] * a[0] + a[3] * a[3] + a[6] * a[6];
    S n1 = a[1] * a[1] + a[4] * a[4] + a[7] * a[7];
    S n2 = a[2] * a[2] + a[5] * a[5] + a[8] * a[8];
    if (n0 < n1) { swap_columns_proper(a, v, 0, 1); S t = n0; n0 = n1; n1 = t; }
    if (n0 < n2) { swap_columns_proper(a, v, 0, 2); S t = n0; n0 = n2; n2 = t; }
    if (n1 < n2) { swap_columns_proper(a, v, 1, 2); S t = n1; n1 = n2; n2 = t; }
    // u1 = a_1 / |a_1|
    S inv = S(1) / sqrt(n0);
    S u1[3] = {a[0] * inv, a[3] * inv, a[6] * inv};
    // u2 = a_2 made orthogonal to u1, normalised; rank-1 covariances get an arbitrary orthogonal u2
    S d12 = a[1] * u1[0] + a[4] * u1[1] + a[7] * u1[2];
    S u2[3] = {a[1] - d12 * u1[0], a[4] - d12 * u1[1], a[7] - d12 * u1[2]};
    S m2 = u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2];
    const S eps = sizeof(S) == 4 ? S(1e-12) : S(1e-28);          // (sigma_2 / sigma_1)^2 floor
    if (!(m2 > eps)) {
        // pick the coordinate axis least aligned with u1 and orthogonalise it
        const S ax = fabs(u1[0]), ay = fabs(u1[1]), az = fabs(u1[2]);
        S e[3] = {0, 0, 0};
        if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
        const S de = e[0] * u1[0] + e[1] * u1[1] + e[2] * u1[2];
        u2[0] = e[0] - de * u1[0]; u2[1] = e[1] - de * u1[1]; u2[2] = e[2] - de * u1[2];
        m2 = u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2];
    }
    inv = S(1) / sqrt(m2);
    u2[0] *= inv; u2[1] *= inv; u2[2] *= inv;
    const S u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    // R = v_1 u_1^T + v_2 u_2^T + v_3 u_3^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            rot[3 * r + c] = v[3 * r + 0] * u1[c] + v[3 * r + 1] * u2[c] + v[3 * r + 2] * u3[c];
}

template <typename S>
ARAP_HD void rotation_from_covariance(const S cov[9], S q[4]) {
    S rot[9];
    rotation_matrix_from_covariance(cov, rot);
    matrix_to_quat(rot, q);
}

// ---- warm-started Newton iteration for the same rotation --------------------------------------------
// R_i maximises tr(R cov) over SO(3) (that is what V diag(1,1,det) U^T of arap.h:376-382 is). Starting
// from the previous ARAP iteration's R_i (rotations change little between iterations), Newton's method
// on SO(3), R <- R exp([omega]x) with (tr(B) I - sym(B)) omega = axl(B), B = cov R, converges
// quadratically. The Hessian K = tr(B) I - sym(B) is positive definite only in the basin of the GLOBAL
// maximum (its eigenvalues there are s2+-s3, s1+-s3, s1+s2), so "K stayed positive definite and the step
// went to zero" certifies the same rotation the SVD gives; anything else returns false and the caller
// falls back to the Jacobi SVD above. ~150 flops per step instead of ~2500 for the fp64 Jacobi sweeps.
template <typename S> struct NewtonTol;
template <> struct NewtonTol<float> {
    static constexpr float step2 = 1e-7f;      // stop when |omega|^2 < step2 (error after the step ~ |omega|^2)
    static constexpr float det_min = 1e-4f;
};
template <> struct NewtonTol<double> {
    static constexpr double step2 = 1e-8;       // |omega| < 1e-4 on the accepted step: remaining rotation error ~1e-8 rad
    static constexpr double det_min = 1e-9;
};

template <typename S>
ARAP_HD S refined_rsqrt(S x) {
    S y = (S)rsqrt_fast((float)x);
    y = y * (S(1.5) - S(0.5) * x * y * y);
    y = y * (S(1.5) - S(0.5) * x * y * y);
    if (sizeof(S) == 8) y = y * (S(1.5) - S(0.5) * x * y * y);
    return y;
}
template <typename S>
ARAP_HD S refined_rcp(S x) {
#if defined(__CUDA_ARCH__)
    S y = (S)__frcp_rn((float)x);
#else
    S y = (S)(1.0f / (float)x);
#endif
    y = y * (S(2) - x * y);
    y = y * (S(2) - x * y);
    if (sizeof(S) == 8) y = y * (S(2) - x * y);
    return y;
}

// c: covariance scaled so that max|c| = 1. q: in = warm start (unit quaternion w,x,y,z), out = result.
// step2_accept: stop once |omega|^2 of the step just taken is below it (the remaining error is ~|omega|^2).
template <typename S>
ARAP_HD bool rotation_newton(const S c[9], S q[4], int max_steps, S step2_accept) {
    for (int it = 0; it < max_steps; ++it) {
        S r[9];
        quat_to_matrix<S>(q[0], q[1], q[2], q[3], r);
        S b[9];                                    // B = C R
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) b[3 * i + j] = c[3 * i] * r[j] + c[3 * i + 1] * r[3 + j] + c[3 * i + 2] * r[6 + j];
        const S g0 = b[5] - b[7], g1 = b[6] - b[2], g2 = b[1] - b[3];
        const S tr = b[0] + b[4] + b[8];
        const S k00 = tr - b[0], k11 = tr - b[4], k22 = tr - b[8];
        const S k01 = S(-0.5) * (b[1] + b[3]), k02 = S(-0.5) * (b[2] + b[6]), k12 = S(-0.5) * (b[5] + b[7]);
        const S c00 = k11 * k22 - k12 * k12, c01 = k02 * k12 - k01 * k22, c02 = k01 * k12 - k02 * k11;
        const S c11 = k00 * k22 - k02 * k02, c12 = k01 * k02 - k00 * k12, c22 = k00 * k11 - k01 * k01;
        const S det = k00 * c00 + k01 * c01 + k02 * c02;
        if (!(k00 > S(0) && c22 > S(0) && det > NewtonTol<S>::det_min)) return false;
        const S inv_det = refined_rcp<S>(det);
        const S ox = S(0.5) * inv_det * (c00 * g0 + c01 * g1 + c02 * g2);     // omega / 2
        const S oy = S(0.5) * inv_det * (c01 * g0 + c11 * g1 + c12 * g2);
        const S oz = S(0.5) * inv_det * (c02 * g0 + c12 * g1 + c22 * g2);
        const S w = q[0], x = q[1], y = q[2], z = q[3];
        S nw = w - x * ox - y * oy - z * oz;
        S nx = x + w * ox + y * oz - z * oy;
        S ny = y + w * oy + z * ox - x * oz;
        S nz = z + w * oz + x * oy - y * ox;
        const S inv_n = refined_rsqrt<S>(nw * nw + nx * nx + ny * ny + nz * nz);
        q[0] = nw * inv_n; q[1] = nx * inv_n; q[2] = ny * inv_n; q[3] = nz * inv_n;
        const S step2 = S(4) * (ox * ox + oy * oy + oz * oz);
        if (!(step2 < S(4))) return false;          // |omega| >= 2 rad (or NaN): not in the Newton basin
        if (step2 < step2_accept) return true;
    }
    return false;
}

// float: plain fp32 Newton.
ARAP_HD bool rotation_newton_certified(const float c[9], float q[4]) {
    return rotation_newton<float>(c, q, 6, NewtonTol<float>::step2);
}
// double: plain fp64 Newton, accepted once |omega| < 1e-4 on the step just taken (remaining error ~1e-8 rad, three
// orders of magnitude inside the 1e-5 x bbox-diagonal parity bar; it does not accumulate: every iteration re-converges).
// (A mixed variant -- fp32 approach, fp64 polish -- measured slower: the kernel is bound by the latency of
// the serial chain per thread, and the extra fp32 steps lengthen it.)
ARAP_HD bool rotation_newton_certified(const double c[9], double q[4]) {
    return rotation_newton<double>(c, q, 6, NewtonTol<double>::step2);
}

// Newton only: true if certified (q_out valid), false if the caller has to run the Jacobi SVD for this covariance.
template <typename S>
ARAP_HD bool rotation_from_covariance_newton_only(const S cov[9], const S q_prev[4], S q_out[4]) {
    S scale = 0;
    for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(cov[i]));
    if (!(scale > S(0))) return false;
    const S inv_scale = S(1) / scale;
    S c[9];
    for (int i = 0; i < 9; ++i) c[i] = cov[i] * inv_scale;
    S q[4] = {q_prev[0], q_prev[1], q_prev[2], q_prev[3]};
    if (!rotation_newton_certified(c, q)) return false;
    q_out[0] = q[0]; q_out[1] = q[1]; q_out[2] = q[2]; q_out[3] = q[3];
    return true;
}

// Local step kernel body: warm-started Newton, Jacobi SVD fallback. q_prev/q_out may alias.
template <typename S>
ARAP_HD void rotation_from_covariance_warm(const S cov[9], const S q_prev[4], S q_out[4]) {
    S scale = 0;
    for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(cov[i]));
    if (scale > S(0)) {
        const S inv_scale = S(1) / scale;
        S c[9];
        for (int i = 0; i < 9; ++i) c[i] = cov[i] * inv_scale;
        S q[4] = {q_prev[0], q_prev[1], q_prev[2], q_prev[3]};
        if (rotation_newton_certified(c, q)) {
            q_out[0] = q[0]; q_out[1] = q[1]; q_out[2] = q[2]; q_out[3] = q[3];
            return;
        }
    }
    rotation_from_covariance<S>(cov, q_out);
}

}  // namespace arap
