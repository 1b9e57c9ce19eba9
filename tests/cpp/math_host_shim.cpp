// tests/cpp/math_host_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the engine's per-element arithmetic (mesh_deform_b200/csrc/arap_math.cuh, which is
// __host__ __device__) for the host so that it can be checked against the oracle on a machine
// without a GPU. The product never runs this code on the CPU.
#include "../../mesh_deform_b200/csrc/arap_math.cuh"

extern "C" {
void math_rotation_f64(const double *cov, double *rot) { arap::rotation_matrix_from_covariance<double>(cov, rot); }
void math_rotation_f32(const float *cov, float *rot) { arap::rotation_matrix_from_covariance<float>(cov, rot); }
void math_quat_f64(const double *cov, double *q) { arap::rotation_from_covariance<double>(cov, q); }
void math_quat_f32(const float *cov, float *q) { arap::rotation_from_covariance<float>(cov, q); }
void math_quat_warm_f64(const double *cov, const double *qprev, double *q) { arap::rotation_from_covariance_warm<double>(cov, qprev, q); }
void math_quat_warm_f32(const float *cov, const float *qprev, float *q) { arap::rotation_from_covariance_warm<float>(cov, qprev, q); }
int math_newton_f64(const double *c, double *q, int steps) { return arap::rotation_newton<double>(c, q, steps, 1e-10) ? 1 : 0; }
int math_newton_f32(const float *c, float *q, int steps) { return arap::rotation_newton<float>(c, q, steps, 1e-7f) ? 1 : 0; }
int math_newton_certified_f64(const double *c, double *q) { return arap::rotation_newton_certified(c, q) ? 1 : 0; }
void math_quat_to_matrix_f64(const double *q, double *r) { arap::quat_to_matrix<double>(q[0], q[1], q[2], q[3], r); }
void math_quat_to_matrix_f32(const float *q, float *r) { arap::quat_to_matrix<float>(q[0], q[1], q[2], q[3], r); }
void math_cotan_f64(const double *v0, const double *v1, const double *v2, double *out) { arap::cotan_half_weights<double>(v0, v1, v2, out); }
void math_cotan_f32(const float *v0, const float *v1, const float *v2, float *out) { arap::cotan_half_weights<float>(v0, v1, v2, out); }
}
