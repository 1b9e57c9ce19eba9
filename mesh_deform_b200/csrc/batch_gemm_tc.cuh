// batch_gemm_tc.cuh -- the batch preconditioner GEMM  Z (V x 3K) = Inv (V x V) . R (V x 3K)  on the 5th-generation tensor
// cores (tcgen05.mma, accumulator in tensor memory), the fast path next to the SIMT mg_batch_dense_kernel (mg_kernels.cuh).
//
// One CTA (128 threads) computes a 128-row x 64-member tile: D (128 x 192, fp32) lives in 192 TMEM columns. The operands of a
// k-stage sit in shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices; SBO = distance
// between core matrices along M/N, LBO = along K) and are consumed by tcgen05.mma.kind::tf32 (M128 N192 K8) issued by one thread.
// Precision: TF32 has a 10-bit mantissa, far too little for a preconditioner that should let CG finish in one or two
// iterations, so both operands are split hi + lo (hi = the nearest TF32 value, lo = the rest, rounded to TF32 as well) and
// three products are accumulated, hi.hi + lo.hi + hi.lo: ~1e-6 relative to sum |a b| (measured), close to fp32 (1e-7).
#pragma once

#include "mg_kernels.cuh"
#include "tma_stage.cuh"

namespace arap {

constexpr int kTcM = 128, kTcMembers = 64, kTcN = kTcMembers * 3;                // tile: rows x members (x 3 components)
constexpr int kTcTmemCols = 256;                                                // >= kTcN, power of two
constexpr uint32_t kTcInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

// nearest TF32 value of a (10-bit mantissa), as an fp32
__device__ __forceinline__ float tf32_round(float a) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a));
    return __uint_as_float(u);
}

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_byte_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start address, LBO, SBO in 16-byte units; version 1; no swizzle
    return (uint64_t)((smem_byte_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- operands pre-packed in global memory, staged by TMA bulk copies ----------------------------------------------------
// (A first version had the CTA's threads convert and stage the operands themselves: 0.19 ms for 642 x 642 x 12288 with the tensor
// pipe idle 90 % of the time, ncu sm__pipe_tensor_cycles_active 9 %; this one takes 0.025 + 0.081 ms.) The operands are stored in
// global memory as ready-made shared-memory images of the stages -- the inverse once at setup (batch_pack_a_kernel), the residuals once per GEMM (batch_pack_b_kernel: the
// fp64 -> hi/lo TF32 conversion is done once instead of once per row tile) -- and one elected thread streams them in with
// cp.async.bulk (complete_tx on the stage's "full" mbarrier) while another issues the MMAs and releases the stage with
// tcgen05.commit on its "empty" mbarrier. No thread touches the operands; the CTA's warps only run the epilogue.
constexpr int kTcK = 32, kTcKcores = kTcK / 4, kTcStages = 2;
constexpr int kTcAFloats = kTcKcores * (kTcM / 8) * 32, kTcBFloats = kTcKcores * (kTcN / 8) * 32;     // per part (hi or lo)
constexpr uint32_t kTcABytes = 2u * kTcAFloats * 4u, kTcBBytes = 2u * kTcBFloats * 4u;                // hi + lo
constexpr size_t kTcSmemBytes = (size_t)kTcStages * (kTcABytes + kTcBBytes) + 128;

// a_pack[mt][ks] = [hi | lo], each [K-core][row][4 floats]  (rows and columns beyond V are zero)
__global__ void __launch_bounds__(kBlock) batch_pack_a_kernel(int V, int ld, int n_mt, int n_ks, const float *__restrict__ inv, float *__restrict__ a_pack) {
    const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;                 // one 16-byte chunk: (mt, ks, kc, row)
    const size_t total = (size_t)n_mt * n_ks * kTcKcores * kTcM;
    if (item >= total) return;
    const int row = (int)(item % kTcM);
    const int kc = (int)((item / kTcM) % kTcKcores);
    const int ks = (int)((item / ((size_t)kTcM * kTcKcores)) % n_ks);
    const int mt = (int)(item / ((size_t)kTcM * kTcKcores * n_ks));
    const int gi = mt * kTcM + row, gk = ks * kTcK + 4 * kc;
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float a = (gi < V && gk + q < V) ? inv[(size_t)gi * ld + gk + q] : 0.f;
        h[q] = tf32_round(a);
        l[q] = tf32_round(a - h[q]);
    }
    float *blk = a_pack + ((size_t)mt * n_ks + ks) * (2 * kTcAFloats);
    const int off = kc * (kTcM / 8) * 32 + row * 4;
    *reinterpret_cast<float4 *>(blk + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(blk + kTcAFloats + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// b_pack[nt][ks] = [hi | lo], each [K-core][column n = 3 member + c][4 floats], from the fp64 residual (member-major Vec3d)
__global__ void __launch_bounds__(kBlock) batch_pack_b_kernel(int V, int K, int n_nt, int n_ks, const Vec3d *__restrict__ r, float *__restrict__ b_pack,
                                                              const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;                 // one 16-byte chunk: (nt, ks, kc, n)
    const size_t total = (size_t)n_nt * n_ks * kTcKcores * kTcN;
    if (item >= total) return;
    const int n = (int)(item % kTcN);
    const int kc = (int)((item / kTcN) % kTcKcores);
    const int ks = (int)((item / ((size_t)kTcN * kTcKcores)) % n_ks);
    const int nt = (int)(item / ((size_t)kTcN * kTcKcores * n_ks));
    const int ml = n / 3, c = n - 3 * ml;
    const int gm = nt * kTcMembers + ml, gk = ks * kTcK + 4 * kc;
    const double *rd = reinterpret_cast<const double *>(r);
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double v = (gm < K && gk + q < V) ? rd[((size_t)gm * V + gk + q) * 3 + c] : 0.0;
        h[q] = tf32_round((float)v);
        l[q] = tf32_round((float)(v - (double)h[q]));
    }
    float *blk = b_pack + ((size_t)nt * n_ks + ks) * (2 * kTcBFloats);
    const int off = kc * (kTcN / 8) * 32 + n * 4;
    *reinterpret_cast<float4 *>(blk + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(blk + kTcBFloats + off) = make_float4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(128, 1) mg_batch_dense_tc_kernel(int V, int K, int n_ks, const float *__restrict__ a_pack, const float *__restrict__ b_pack,
                                                                   MgVec *__restrict__ z, const CgScalars *__restrict__ cg) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    if (cg->converged) return;
    unsigned char *stage_base = tc_smem;                                              // per stage: [A hi | A lo | B hi | B lo]
    uint64_t *bar = reinterpret_cast<uint64_t *>(tc_smem + (size_t)kTcStages * (kTcABytes + kTcBBytes));   // full[S], empty[S], done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2 * kTcStages + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int mt = blockIdx.x, nt = blockIdx.y;
    const int i0 = mt * kTcM, m0 = nt * kTcMembers;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_addr(tmem_slot)), "n"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int b = 0; b < 2 * kTcStages + 1; ++b) tma::mbar_init(&bar[b], 1);
        tma::fence_barrier_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    uint64_t *full = bar, *empty = bar + kTcStages, *done = bar + 2 * kTcStages;

    if (tid == 0) {                               // ---- producer: stream the stage images in
        const float *a_src = a_pack + (size_t)mt * n_ks * (2 * kTcAFloats);
        const float *b_src = b_pack + (size_t)nt * n_ks * (2 * kTcBFloats);
        for (int s = 0; s < n_ks; ++s) {
            const int buf = s % kTcStages;
            if (s >= kTcStages) tma::mbar_wait(&empty[buf], (uint32_t)((s / kTcStages - 1) & 1));
            unsigned char *dst = stage_base + (size_t)buf * (kTcABytes + kTcBBytes);
            tma::mbar_arrive_expect_tx(&full[buf], kTcABytes + kTcBBytes);
            tma::bulk_g2s(dst, a_src + (size_t)s * (2 * kTcAFloats), kTcABytes, &full[buf]);
            tma::bulk_g2s(dst + kTcABytes, b_src + (size_t)s * (2 * kTcBFloats), kTcBBytes, &full[buf]);
        }
    } else if (tid == 32) {                       // ---- MMA issuer
        constexpr uint32_t kALbo = (kTcM / 8) * 128, kBLbo = (kTcN / 8) * 128, kSbo = 128;       // bytes
        for (int s = 0; s < n_ks; ++s) {
            const int buf = s % kTcStages;
            tma::mbar_wait(&full[buf], (uint32_t)((s / kTcStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi_s = tma::smem_addr(stage_base + (size_t)buf * (kTcABytes + kTcBBytes)), a_lo_s = a_hi_s + 4 * kTcAFloats;
            const uint32_t b_hi_s = a_hi_s + kTcABytes, b_lo_s = b_hi_s + 4 * kTcBFloats;
#pragma unroll
            for (int kk = 0; kk < kTcK / 8; ++kk) {
                const uint64_t dah = umma_desc_kmajor(a_hi_s + kk * 2 * kALbo, kALbo, kSbo), dal = umma_desc_kmajor(a_lo_s + kk * 2 * kALbo, kALbo, kSbo);
                const uint64_t dbh = umma_desc_kmajor(b_hi_s + kk * 2 * kBLbo, kBLbo, kSbo), dbl = umma_desc_kmajor(b_lo_s + kk * 2 * kBLbo, kBLbo, kSbo);
                umma_tf32(tmem_d, dah, dbh, kTcInstrDesc, (s == 0 && kk == 0) ? 0u : 1u);
                umma_tf32(tmem_d, dal, dbh, kTcInstrDesc, 1u);
                umma_tf32(tmem_d, dah, dbl, kTcInstrDesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_addr(&empty[buf])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_addr(done)) : "memory");
    }
    // ---- epilogue (all four warps): TMEM lane = row of the tile, column = 3 * member + component
    tma::mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int gi = i0 + tid;
    for (int c0 = 0; c0 < kTcN; c0 += 24) {                                       // 8 members per round
        uint32_t v[24];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr + 8));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]) : "r"(taddr + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (gi < V) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int gm = m0 + c0 / 3 + q;
                if (gm < K) z[(size_t)gm * V + gi] = MgVec{__uint_as_float(v[3 * q]), __uint_as_float(v[3 * q + 1]), __uint_as_float(v[3 * q + 2]), 0.f};
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTcTmemCols) : "memory");
}

}  // namespace arap
