/*
 * oracle/trajectory_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (double precision) of the reference's trajectory front end,
 * reference inc/deform/trajectory.h:61-73 (TrajectorySE3::operator()), and of
 * reference inc/deform/deformation_util.h:48-57 (DeformationUtil::updateConstraints).
 *
 * The arithmetic of trajectory.h lives in two un-vendored dependencies that are absent from
 * /root/reference and from this image (versions unpinned: cmake/FindSophus.cmake:4-10,
 * cmake/FindEigen3.cmake:4-10); their published algorithms are restated here:
 *   - Sophus::SE3Group<S>(Matrix4).log() / SE3Group::exp(tangent).affine3()   (trajectory.h:65,72)
 *       tangent = [upsilon (translation part); omega (rotation vector)],
 *       exp: R = exp_SO3(omega), t = V upsilon;  log: omega = log_SO3(R), upsilon = V^-1 t.
 *   - Eigen::SplineFitting<Spline<S,6,3>>::Interpolate(points, 3)              (trajectory.h:67)
 *       chord-length parameters, knot averaging, collocation solve for the control points;
 *       Spline::operator()(u) = sum_j N_j(u) ctrl_j on the knot span containing u (trajectory.h:71).
 *
 * PARITY PIN: reference tests/test_trajectory.cpp:23-38 (endpoints of a 4-key-pose path, 1e-3).
 * Interior values are unpinned by the reference.
 *
 * Matrices are 4x4 row-major double[16].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void cross_matrix(const double w[3], double O[9]) {
    O[0] = 0; O[1] = -w[2]; O[2] = w[1];
    O[3] = w[2]; O[4] = 0; O[5] = -w[0];
    O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}
static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = s;
        }
}

/* rotation matrix -> unit quaternion (w, x, y, z), w >= 0 */
static void quat_from_matrix(const double R[9], double q[4]) {
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double sgn = q[0] < 0 ? -1.0 : 1.0;
    for (int i = 0; i < 4; ++i) q[i] *= sgn / n;
}

/* Sophus SO3 log: rotation vector of minimal angle */
static void so3_log(const double R[9], double w[3]) {
    double q[4];
    quat_from_matrix(R, q);
    double n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3], n = sqrt(n2), k;
    if (n < 1e-10) k = 2.0 / q[0] - 2.0 * n2 / (3.0 * q[0] * q[0] * q[0]);
    else k = 2.0 * atan2(n, q[0]) / n;
    for (int i = 0; i < 3; ++i) w[i] = k * q[1 + i];
}

/* coefficients of R = I + a O + b O^2 and V = I + b O + c O^2, theta = |omega| */
static void so3_coeffs(double theta, double *a, double *b, double *c) {
    if (theta < 1e-6) {
        double t2 = theta * theta;
        *a = 1.0 - t2 / 6.0; *b = 0.5 - t2 / 24.0; *c = 1.0 / 6.0 - t2 / 120.0;
    } else {
        *a = sin(theta) / theta;
        *b = (1.0 - cos(theta)) / (theta * theta);
        *c = (theta - sin(theta)) / (theta * theta * theta);
    }
}

/* Sophus SE3Group::exp (trajectory.h:72): xi = [upsilon; omega] -> 4x4 */
void traj_se3_exp(const double xi[6], double T[16]) {
    const double *ups = xi, *om = xi + 3;
    double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    double a, b, c, O[9], O2[9];
    so3_coeffs(theta, &a, &b, &c);
    cross_matrix(om, O);
    mat3_mul(O, O, O2);
    memset(T, 0, sizeof(double) * 16);
    T[15] = 1;
    for (int i = 0; i < 3; ++i) {
        double ti = 0;
        for (int j = 0; j < 3; ++j) {
            double I = (i == j) ? 1.0 : 0.0;
            T[4 * i + j] = I + a * O[3 * i + j] + b * O2[3 * i + j];
            ti += (I + b * O[3 * i + j] + c * O2[3 * i + j]) * ups[j];
        }
        T[4 * i + 3] = ti;
    }
}

/* Sophus SE3Group(Matrix4).log() (trajectory.h:65): 4x4 -> xi = [upsilon; omega] */
void traj_se3_log(const double T[16], double xi[6]) {
    double R[9], t[3], om[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[3 * i + j] = T[4 * i + j]; t[i] = T[4 * i + 3]; }
    so3_log(R, om);
    double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    double O[9], O2[9], k;
    cross_matrix(om, O);
    mat3_mul(O, O, O2);
    if (theta < 1e-6) k = 1.0 / 12.0 + theta * theta / 720.0;
    else k = (1.0 - theta * cos(0.5 * theta) / (2.0 * sin(0.5 * theta))) / (theta * theta);
    for (int i = 0; i < 3; ++i) {
        double u = 0;
        for (int j = 0; j < 3; ++j) u += (((i == j) ? 1.0 : 0.0) - 0.5 * O[3 * i + j] + k * O2[3 * i + j]) * t[j];
        xi[i] = u;
        xi[3 + i] = om[i];
    }
}

/* B-spline machinery of Eigen's unsupported Splines module ---------------------------------- */

/* Spline::Span: index of the knot span containing u */
static int bs_span(double u, int degree, const double *knots, int nknots) {
    if (u <= knots[0]) return degree;
    int lo = degree - 1, hi = nknots - degree - 1;          /* upper_bound over knots[lo, hi) */
    int pos = hi;
    for (int i = lo; i < hi; ++i) if (knots[i] > u) { pos = i; break; }
    return pos - 1;
}

/* Spline::BasisFunctions: the degree+1 non-vanishing basis functions at u (Cox-de Boor) */
static void bs_basis(double u, int degree, const double *knots, int span, double *N) {
    double left[8], right[8];
    N[0] = 1.0;
    for (int j = 1; j <= degree; ++j) {
        left[j] = u - knots[span + 1 - j];
        right[j] = knots[span + j] - u;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            double tmp = N[r] / (right[r + 1] + left[j - r]);
            N[r] = saved + right[r + 1] * tmp;
            saved = left[j - r] * tmp;
        }
        N[j] = saved;
    }
}

/* dense solve A X = B (n x n, m right-hand sides), partial pivoting; stands in for HouseholderQR */
static int dense_solve(int n, int m, double *A, double *B) {
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
        if (A[piv * n + c] == 0.0) return 0;
        if (piv != c) {
            for (int k = 0; k < n; ++k) { double t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
            for (int k = 0; k < m; ++k) { double t = B[c * m + k]; B[c * m + k] = B[piv * m + k]; B[piv * m + k] = t; }
        }
        for (int r = c + 1; r < n; ++r) {
            double f = A[r * n + c] / A[c * n + c];
            if (f == 0.0) continue;
            for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
            for (int k = 0; k < m; ++k) B[r * m + k] -= f * B[c * m + k];
        }
    }
    for (int c = n - 1; c >= 0; --c)
        for (int k = 0; k < m; ++k) {
            double s = B[c * m + k];
            for (int r = c + 1; r < n; ++r) s -= A[c * n + r] * B[r * m + k];
            B[c * m + k] = s / A[c * n + c];
        }
    return 1;
}

/*
 * SplineFitting::Interpolate(points, 3) on the se(3) logs of the key poses (trajectory.h:62-69).
 * poses: n 4x4 matrices. Outputs: ctrl[n*6] (control point j at ctrl+6j), knots[n+4], params[n].
 * Returns 1 on success (needs n >= 4).
 */
int traj_fit(int n, const double *poses, double *ctrl, double *knots, double *params) {
    const int degree = 3;
    if (n < degree + 1) return 0;
    double *pts = (double *)malloc(sizeof(double) * 6 * (size_t)n);
    for (int i = 0; i < n; ++i) traj_se3_log(poses + 16 * (size_t)i, pts + 6 * (size_t)i);
    /* ChordLengths */
    params[0] = 0;
    for (int i = 1; i < n; ++i) {
        double s = 0;
        for (int d = 0; d < 6; ++d) { double e = pts[6 * i + d] - pts[6 * (i - 1) + d]; s += e * e; }
        params[i] = params[i - 1] + sqrt(s);
    }
    double total = params[n - 1];
    for (int i = 0; i < n; ++i) params[i] /= total;
    params[n - 1] = 1.0;
    /* KnotAveraging */
    int nk = n + degree + 1;
    for (int j = 1; j < n - degree; ++j) {
        double s = 0;
        for (int k = 0; k < degree; ++k) s += params[j + k];
        knots[j + degree] = s / degree;
    }
    for (int k = 0; k <= degree; ++k) { knots[k] = 0.0; knots[nk - 1 - k] = 1.0; }
    /* collocation matrix */
    double *A = (double *)calloc((size_t)n * n, sizeof(double));
    for (int i = 1; i < n - 1; ++i) {
        int span = bs_span(params[i], degree, knots, nk);
        double N[4];
        bs_basis(params[i], degree, knots, span, N);
        for (int k = 0; k <= degree; ++k) A[i * n + span - degree + k] = N[k];
    }
    A[0] = 1.0; A[(size_t)n * n - 1] = 1.0;
    memcpy(ctrl, pts, sizeof(double) * 6 * (size_t)n);
    int ok = dense_solve(n, 6, A, ctrl);
    free(A); free(pts);
    return ok;
}

/* _spline(time) then SE3Group::exp(...).affine3() (trajectory.h:71-72) */
void traj_eval(int n, const double *ctrl, const double *knots, double u, double T[16]) {
    const int degree = 3;
    int nk = n + degree + 1;
    int span = bs_span(u, degree, knots, nk);
    double N[4], xi[6] = {0, 0, 0, 0, 0, 0};
    bs_basis(u, degree, knots, span, N);
    for (int k = 0; k <= degree; ++k)
        for (int d = 0; d < 6; ++d) xi[d] += N[k] * ctrl[6 * (size_t)(span - degree + k) + d];
    traj_se3_exp(xi, T);
}

/*
 * DeformationUtil::updateConstraints (deformation_util.h:48-57): tabs = origin * t * origin^-1
 * (origin^-1 taken as an isometry inverse, :38), targets_i = tabs * p0_i.
 * points: H x 3 handle rest positions; out: H x 3.
 */
void traj_handle_targets(const double origin[16], const double t[16], int H, const double *points, double *out) {
    double inv[16], tmp[16], tabs[16];
    memset(inv, 0, sizeof(inv));
    inv[15] = 1;
    for (int i = 0; i < 3; ++i) {
        double ti = 0;
        for (int j = 0; j < 3; ++j) { inv[4 * i + j] = origin[4 * j + i]; ti -= origin[4 * j + i] * origin[4 * j + 3]; }
        inv[4 * i + 3] = ti;
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += origin[4 * i + k] * t[4 * k + j];
            tmp[4 * i + j] = s;
        }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += tmp[4 * i + k] * inv[4 * k + j];
            tabs[4 * i + j] = s;
        }
    for (int h = 0; h < H; ++h)
        for (int i = 0; i < 3; ++i)
            out[3 * h + i] = tabs[4 * i] * points[3 * h] + tabs[4 * i + 1] * points[3 * h + 1] + tabs[4 * i + 2] * points[3 * h + 2] + tabs[4 * i + 3];
}
