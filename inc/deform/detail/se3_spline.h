// deform/detail/se3_spline.h -- array-only SE(3) log/exp and the interpolating cubic B-spline through se(3) logs.
//
// The arithmetic behind deform::TrajectorySE3 (reference inc/deform/trajectory.h:61-73) and
// deform::DeformationUtil::updateConstraints (reference inc/deform/deformation_util.h:48-57), free of any matrix
// library so that the header-only C++ API (inc/deform/trajectory.h) and the C ABI (arap_trajectory_*,
// include/arap_b200.h) share one implementation. The reference takes these from two libraries this build does not
// depend on; their published algorithms are implemented here directly:
//   * Sophus::SE3Group log / exp  (reference trajectory.h:65,72): tangent = [translation part; rotation vector],
//     exp: R = exp_SO3(w), t = V u;  log: w = log_SO3(R), u = V^-1 t.
//   * Eigen::SplineFitting<Spline<S,6,3>>::Interpolate(points, 3)  (reference trajectory.h:67): chord-length
//     parameters, knot averaging, collocation solve for the control points; evaluation by the
//     Cox-de Boor recurrence on the knot span containing the parameter (reference trajectory.h:71).
// All transforms are 4x4 row-major arrays of 16 scalars.
#ifndef DEFORM_DETAIL_SE3_SPLINE_H
#define DEFORM_DETAIL_SE3_SPLINE_H

#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

namespace deform {

namespace se3 {

/** 6-vector [upsilon; omega] -> rigid transform (4x4 row-major in `T`). */
template <class S>
void exp(const S xi[6], S T[16]) {
    const S wx = xi[3], wy = xi[4], wz = xi[5];
    const S theta2 = wx * wx + wy * wy + wz * wz, theta = std::sqrt(theta2);
    S a, b, c;   // R = I + a W + b W^2 ; V = I + b W + c W^2
    if (theta < S(1e-4)) { a = S(1) - theta2 / 6; b = S(0.5) - theta2 / 24; c = S(1) / 6 - theta2 / 120; }
    else { a = std::sin(theta) / theta; b = (S(1) - std::cos(theta)) / theta2; c = (theta - std::sin(theta)) / (theta2 * theta); }
    const S W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    S W2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { S s = 0; for (int k = 0; k < 3; ++k) s += W[3 * i + k] * W[3 * k + j]; W2[3 * i + j] = s; }
    for (int i = 0; i < 16; ++i) T[i] = 0;
    T[15] = 1;
    for (int i = 0; i < 3; ++i) {
        S t = 0;
        for (int j = 0; j < 3; ++j) {
            const S I = (i == j) ? S(1) : S(0);
            T[4 * i + j] = I + a * W[3 * i + j] + b * W2[3 * i + j];
            t += (I + b * W[3 * i + j] + c * W2[3 * i + j]) * xi[j];
        }
        T[4 * i + 3] = t;
    }
}

/** rigid transform (4x4 row-major) -> 6-vector [upsilon; omega], rotation angle in [0, pi]. */
template <class S>
void log(const S T[16], S xi[6]) {
    // rotation -> unit quaternion (w >= 0), largest-component branch for accuracy
    const S r00 = T[0], r01 = T[1], r02 = T[2], r10 = T[4], r11 = T[5], r12 = T[6], r20 = T[8], r21 = T[9], r22 = T[10];
    S q[4];
    const S tr = r00 + r11 + r22;
    if (tr > 0) { const S s = std::sqrt(tr + 1) * 2; q[0] = s / 4; q[1] = (r21 - r12) / s; q[2] = (r02 - r20) / s; q[3] = (r10 - r01) / s; }
    else if (r00 > r11 && r00 > r22) { const S s = std::sqrt(1 + r00 - r11 - r22) * 2; q[0] = (r21 - r12) / s; q[1] = s / 4; q[2] = (r01 + r10) / s; q[3] = (r02 + r20) / s; }
    else if (r11 > r22) { const S s = std::sqrt(1 + r11 - r00 - r22) * 2; q[0] = (r02 - r20) / s; q[1] = (r01 + r10) / s; q[2] = s / 4; q[3] = (r12 + r21) / s; }
    else { const S s = std::sqrt(1 + r22 - r00 - r11) * 2; q[0] = (r10 - r01) / s; q[1] = (r02 + r20) / s; q[2] = (r12 + r21) / s; q[3] = s / 4; }
    const S qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const S sg = (q[0] < 0 ? S(-1) : S(1)) / qn;
    for (int i = 0; i < 4; ++i) q[i] *= sg;
    const S n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3], n = std::sqrt(n2);
    const S k = (n < S(1e-6)) ? (S(2) / q[0] - S(2) * n2 / (3 * q[0] * q[0] * q[0])) : (S(2) * std::atan2(n, q[0]) / n);
    const S wx = k * q[1], wy = k * q[2], wz = k * q[3];
    const S theta2 = wx * wx + wy * wy + wz * wz, theta = std::sqrt(theta2);
    // V^-1 = I - W/2 + g W^2
    const S g = (theta < S(1e-4)) ? (S(1) / 12 + theta2 / 720) : ((S(1) - theta * std::cos(theta / 2) / (2 * std::sin(theta / 2))) / theta2);
    const S W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    S W2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { S s = 0; for (int m = 0; m < 3; ++m) s += W[3 * i + m] * W[3 * m + j]; W2[3 * i + j] = s; }
    const S t[3] = {T[3], T[7], T[11]};
    for (int i = 0; i < 3; ++i) {
        S u = 0;
        for (int j = 0; j < 3; ++j) u += (((i == j) ? S(1) : S(0)) - W[3 * i + j] / 2 + g * W2[3 * i + j]) * t[j];
        xi[i] = u;
    }
    xi[3] = wx; xi[4] = wy; xi[5] = wz;
}

}  // namespace se3

namespace detail {

/** out = origin * t * origin^-1 with origin^-1 taken as an isometry inverse (reference deformation_util.h:38,51). */
template <class S>
void conjugate_rigid(const S origin[16], const S t[16], S out[16]) {
    S inv[16], tmp[16];
    for (int i = 0; i < 16; ++i) inv[i] = 0;
    inv[15] = 1;
    for (int i = 0; i < 3; ++i) {
        S ti = 0;
        for (int j = 0; j < 3; ++j) { inv[4 * i + j] = origin[4 * j + i]; ti -= origin[4 * j + i] * origin[4 * j + 3]; }
        inv[4 * i + 3] = ti;
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { S s = 0; for (int k = 0; k < 4; ++k) s += origin[4 * i + k] * t[4 * k + j]; tmp[4 * i + j] = s; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { S s = 0; for (int k = 0; k < 4; ++k) s += tmp[4 * i + k] * inv[4 * k + j]; out[4 * i + j] = s; }
}

/** Key poses in, smooth pose curve out. Needs >= 4 key poses before evaluation (degree 3). */
template <class S>
class SplineSE3 {
public:
    typedef S Scalar;
    SplineSE3() : _dirty(false) {}

    void addKeyPose(const Scalar T[16]) { _poses.insert(_poses.end(), T, T + 16); _dirty = true; }
    int numberOfKeyPoses() const { return (int)_poses.size() / 16; }

    /** Pose at u in [0,1] as a 4x4 row-major matrix. */
    void pose(Scalar u, Scalar T[16]) {
        if (_dirty) { fit(); _dirty = false; }
        Scalar xi[6];
        evaluate(u, xi);
        se3::exp<Scalar>(xi, T);
    }

private:
    enum { Degree = 3 };

    // index of the knot span containing u (Eigen Spline::Span)
    int span(Scalar u) const {
        const int nk = (int)_knots.size();
        if (u <= _knots[0]) return Degree;
        int pos = nk - Degree - 1;
        for (int i = Degree - 1; i < nk - Degree - 1; ++i) if (_knots[(size_t)i] > u) { pos = i; break; }
        return pos - 1;
    }
    // the Degree+1 non-vanishing B-spline basis functions at u
    void basis(Scalar u, int sp, Scalar N[Degree + 1]) const {
        Scalar left[Degree + 1], right[Degree + 1];
        N[0] = 1;
        for (int j = 1; j <= Degree; ++j) {
            left[j] = u - _knots[(size_t)(sp + 1 - j)];
            right[j] = _knots[(size_t)(sp + j)] - u;
            Scalar saved = 0;
            for (int r = 0; r < j; ++r) {
                const Scalar tmp = N[r] / (right[r + 1] + left[j - r]);
                N[r] = saved + right[r + 1] * tmp;
                saved = left[j - r] * tmp;
            }
            N[j] = saved;
        }
    }
    void evaluate(Scalar u, Scalar xi[6]) const {
        const int sp = span(u);
        Scalar N[Degree + 1];
        basis(u, sp, N);
        for (int d = 0; d < 6; ++d) xi[d] = 0;
        for (int k = 0; k <= Degree; ++k)
            for (int d = 0; d < 6; ++d) xi[d] += N[k] * _ctrl[6 * (size_t)(sp - Degree + k) + d];
    }
    // SplineFitting::Interpolate(points, 3) on the se(3) logs of the key poses
    void fit() {
        const int n = numberOfKeyPoses();
        std::vector<Scalar> pts(6 * (size_t)n);
        for (int i = 0; i < n; ++i) se3::log<Scalar>(&_poses[16 * (size_t)i], &pts[6 * (size_t)i]);
        std::vector<Scalar> u((size_t)n, 0);                        // chord-length parameters
        for (int i = 1; i < n; ++i) {
            Scalar s = 0;
            for (int d = 0; d < 6; ++d) { const Scalar e = pts[6 * (size_t)i + d] - pts[6 * (size_t)(i - 1) + d]; s += e * e; }
            u[(size_t)i] = u[(size_t)i - 1] + std::sqrt(s);
        }
        const Scalar total = u[(size_t)n - 1];
        for (int i = 0; i < n; ++i) u[(size_t)i] = total > 0 ? u[(size_t)i] / total : Scalar(i) / Scalar(n - 1);
        u[(size_t)n - 1] = 1;
        _knots.assign((size_t)(n + Degree + 1), 0);                 // knot averaging
        for (int j = 1; j < n - Degree; ++j) {
            Scalar s = 0;
            for (int k = 0; k < Degree; ++k) s += u[(size_t)(j + k)];
            _knots[(size_t)(j + Degree)] = s / Degree;
        }
        for (int k = 0; k <= Degree; ++k) _knots[_knots.size() - 1 - (size_t)k] = 1;
        std::vector<Scalar> A((size_t)n * n, 0);                    // collocation matrix
        for (int i = 1; i < n - 1; ++i) {
            const int sp = span(u[(size_t)i]);
            Scalar N[Degree + 1];
            basis(u[(size_t)i], sp, N);
            for (int k = 0; k <= Degree; ++k) A[(size_t)i * n + (size_t)(sp - Degree + k)] = N[k];
        }
        A[0] = 1;
        A[(size_t)n * n - 1] = 1;
        _ctrl = pts;                                                // solve A ctrl = pts (Gaussian elimination, partial pivoting)
        for (int c = 0; c < n; ++c) {
            int piv = c;
            for (int r = c + 1; r < n; ++r) if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)piv * n + c])) piv = r;
            if (piv != c) {
                for (int k = 0; k < n; ++k) std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]);
                for (int k = 0; k < 6; ++k) std::swap(_ctrl[6 * (size_t)c + k], _ctrl[6 * (size_t)piv + k]);
            }
            for (int r = c + 1; r < n; ++r) {
                const Scalar f = A[(size_t)r * n + c] / A[(size_t)c * n + c];
                if (f == 0) continue;
                for (int k = c; k < n; ++k) A[(size_t)r * n + k] -= f * A[(size_t)c * n + k];
                for (int k = 0; k < 6; ++k) _ctrl[6 * (size_t)r + k] -= f * _ctrl[6 * (size_t)c + k];
            }
        }
        for (int c = n - 1; c >= 0; --c)
            for (int k = 0; k < 6; ++k) {
                Scalar s = _ctrl[6 * (size_t)c + k];
                for (int r = c + 1; r < n; ++r) s -= A[(size_t)c * n + r] * _ctrl[6 * (size_t)r + k];
                _ctrl[6 * (size_t)c + k] = s / A[(size_t)c * n + c];
            }
    }

    bool _dirty;
    std::vector<Scalar> _poses;   // 4x4 row-major per key pose
    std::vector<Scalar> _ctrl;    // 6 per control point
    std::vector<Scalar> _knots;
};

}  // namespace detail
}  // namespace deform

#endif
