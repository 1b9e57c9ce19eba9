// tests/cuda/batch_gemm_check.cu -- TEST PROGRAM (run on the GPU box): the tcgen05 batch GEMM against the SIMT one and a CPU
// double-precision reference on random data, plus timings.   usage: batch_gemm_check [V] [K]
#include "../../mesh_deform_b200/csrc/batch_gemm_tc.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace arap;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char **argv) {
    const int V = argc > 1 ? std::atoi(argv[1]) : 642, K = argc > 2 ? std::atoi(argv[2]) : 4096;
    const int ld = (V + 3) & ~3;
    std::vector<float> inv((size_t)V * ld, 0.f);
    std::vector<Vec3d> r((size_t)K * V);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / 16777216.0 - 0.5; };
    for (int i = 0; i < V; ++i) for (int k = 0; k < V; ++k) inv[(size_t)i * ld + k] = (float)(rnd() * (i == k ? 4.0 : 0.3));
    for (auto &v : r) v = Vec3d{rnd(), rnd() * 1e-3, rnd() * 10.0};
    float *d_inv, *d_hi, *d_lo; Vec3d *d_r; MgVec *d_z1, *d_z2; CgScalars *d_cg;
    CK(cudaMalloc(&d_inv, inv.size() * 4)); CK(cudaMalloc(&d_hi, inv.size() * 4)); CK(cudaMalloc(&d_lo, inv.size() * 4));
    CK(cudaMalloc(&d_r, r.size() * sizeof(Vec3d))); CK(cudaMalloc(&d_z1, r.size() * sizeof(MgVec))); CK(cudaMalloc(&d_z2, r.size() * sizeof(MgVec)));
    CK(cudaMalloc(&d_cg, sizeof(CgScalars))); CK(cudaMemset(d_cg, 0, sizeof(CgScalars)));
    CK(cudaMemcpy(d_inv, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_r, r.data(), r.size() * sizeof(Vec3d), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_z1, 0xff, r.size() * sizeof(MgVec))); CK(cudaMemset(d_z2, 0xff, r.size() * sizeof(MgVec)));
    tf32_split_kernel<<<(unsigned)((inv.size() + kBlock - 1) / kBlock), kBlock>>>(inv.size(), d_inv, d_hi, d_lo);
    CK(cudaFuncSetAttribute(mg_batch_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    const dim3 g1((V + kBgM - 1) / kBgM, (K + kBgMembers - 1) / kBgMembers), g2((V + kTcM - 1) / kTcM, (K + kTcMembers - 1) / kTcMembers);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms1 = 0, ms2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mg_batch_dense_kernel<<<g1, 256>>>(V, ld, K, d_inv, d_r, d_z1, d_cg);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventRecord(e0);
        mg_batch_dense_tc_kernel<<<g2, kTcThreads, kTcSmemBytes>>>(V, ld, K, d_hi, d_lo, d_r, d_z2, d_cg);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms2, e0, e1);
    }
    CK(cudaGetLastError());
    std::vector<MgVec> z1(r.size()), z2(r.size());
    CK(cudaMemcpy(z1.data(), d_z1, r.size() * sizeof(MgVec), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(z2.data(), d_z2, r.size() * sizeof(MgVec), cudaMemcpyDeviceToHost));
    double e_simt = 0, e_tc = 0, scale = 0;
    const int members[] = {0, 1, 63, 64, K / 2 + 7, K - 1};
    for (int m : members) {
        if (m < 0 || m >= K) continue;
        for (int i = 0; i < V; i += (V > 200 ? 5 : 1)) {
            double ref[3] = {0, 0, 0}, mag[3] = {0, 0, 0};
            for (int k = 0; k < V; ++k) {
                const double a = inv[(size_t)i * ld + k]; const Vec3d &b = r[(size_t)m * V + k];
                ref[0] += a * b.x; ref[1] += a * b.y; ref[2] += a * b.z;
                mag[0] += std::fabs(a * b.x); mag[1] += std::fabs(a * b.y); mag[2] += std::fabs(a * b.z);
            }
            const MgVec a = z1[(size_t)m * V + i], b = z2[(size_t)m * V + i];
            const double g1v[3] = {a.x, a.y, a.z}, g2v[3] = {b.x, b.y, b.z};
            for (int c = 0; c < 3; ++c) {       // error relative to sum |a_k b_k|: what rounding in the products can produce
                e_simt = std::fmax(e_simt, std::fabs(g1v[c] - ref[c]) / mag[c]); e_tc = std::fmax(e_tc, std::fabs(g2v[c] - ref[c]) / mag[c]); scale = std::fmax(scale, std::fabs(ref[c]));
            }
        }
    }
    const double gflop = 2.0 * V * (double)V * 3.0 * K * 1e-9;
    std::printf("V=%d K=%d  SIMT %.3f ms (%.1f TFLOP/s) max rel err %.2e | tcgen05 3xTF32 %.3f ms (%.1f TFLOP/s useful) max rel err %.2e | max|ref| %.2e\n",
                V, K, ms1, gflop / ms1, e_simt, ms2, gflop / ms2, e_tc, scale);
    std::printf("%s\n", (e_tc < 2e-6 && e_simt < 2e-6) ? "GEMM CHECK OK" : "GEMM CHECK FAILED");
    return 0;
}
