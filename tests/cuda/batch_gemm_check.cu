// tests/cuda/batch_gemm_check.cu -- TEST PROGRAM (run on the GPU box): the tcgen05 batch GEMM (operands pre-packed, staged by TMA
// bulk copies, 3xTF32) against the SIMT fp32 one and a CPU double-precision reference on random data, plus timings.
//   usage: batch_gemm_check [V] [K]
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
    const int n_mt = (V + kTcM - 1) / kTcM, n_nt = (K + kTcMembers - 1) / kTcMembers, n_ks = (V + kTcK - 1) / kTcK;
    float *d_inv, *d_apack, *d_bpack; Vec3d *d_r; MgVec *d_z1, *d_z2; CgScalars *d_cg;
    CK(cudaMalloc(&d_inv, inv.size() * 4));
    CK(cudaMalloc(&d_r, r.size() * sizeof(Vec3d))); CK(cudaMalloc(&d_z1, r.size() * sizeof(MgVec))); CK(cudaMalloc(&d_z2, r.size() * sizeof(MgVec)));
    CK(cudaMalloc(&d_apack, sizeof(float) * (size_t)n_mt * n_ks * 2 * kTcAFloats)); CK(cudaMalloc(&d_bpack, sizeof(float) * (size_t)n_nt * n_ks * 2 * kTcBFloats));
    CK(cudaMalloc(&d_cg, sizeof(CgScalars))); CK(cudaMemset(d_cg, 0, sizeof(CgScalars)));
    CK(cudaMemcpy(d_inv, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_r, r.data(), r.size() * sizeof(Vec3d), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_z1, 0xff, r.size() * sizeof(MgVec))); CK(cudaMemset(d_z2, 0xff, r.size() * sizeof(MgVec)));
    CK(cudaFuncSetAttribute(mg_batch_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    {
        const size_t items = (size_t)n_mt * n_ks * kTcKcores * kTcM;
        batch_pack_a_kernel<<<(unsigned)((items + kBlock - 1) / kBlock), kBlock>>>(V, ld, n_mt, n_ks, d_inv, d_apack);
    }
    const dim3 g1((V + kBgM - 1) / kBgM, (K + kBgMembers - 1) / kBgMembers);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_simt = 0, ms_pack = 0, ms_tc = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mg_batch_dense_kernel<<<g1, 256>>>(V, ld, K, d_inv, d_r, d_z1, d_cg);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_simt, e0, e1);
        const size_t items = (size_t)n_nt * n_ks * kTcKcores * kTcN;
        cudaEventRecord(e0);
        batch_pack_b_kernel<<<(unsigned)((items + kBlock - 1) / kBlock), kBlock>>>(V, K, n_nt, n_ks, d_r, d_bpack, d_cg);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_pack, e0, e1);
        cudaEventRecord(e0);
        mg_batch_dense_tc_kernel<<<dim3(n_mt, n_nt), 128, kTcSmemBytes>>>(V, K, n_ks, d_apack, d_bpack, d_z2, d_cg);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_tc, e0, e1);
    }
    CK(cudaGetLastError());
    std::vector<MgVec> z1(r.size()), z2(r.size());
    CK(cudaMemcpy(z1.data(), d_z1, r.size() * sizeof(MgVec), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(z2.data(), d_z2, r.size() * sizeof(MgVec), cudaMemcpyDeviceToHost));
    double e_simt = 0, e_tc = 0;
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
                e_simt = std::fmax(e_simt, std::fabs(g1v[c] - ref[c]) / mag[c]);
                e_tc = std::fmax(e_tc, std::fabs(g2v[c] - ref[c]) / mag[c]);
            }
        }
    }
    const double gflop = 2.0 * V * (double)V * 3.0 * K * 1e-9;
    std::printf("V=%d K=%d  SIMT fp32 %.3f ms (%.1f TFLOP/s) max rel err %.2e | tcgen05 3xTF32: pack %.3f ms + gemm %.3f ms (%.1f TFLOP/s useful, x3 issued) "
                "max rel err %.2e\n", V, K, ms_simt, gflop / ms_simt, e_simt, ms_pack, ms_tc, gflop / ms_tc, e_tc);
    std::printf("%s\n", (e_tc < 2e-6 && e_simt < 2e-6) ? "GEMM CHECK OK" : "GEMM CHECK FAILED");
    return 0;
}
