// batch_gemm_tc.cuh -- the batch preconditioner GEMM  Z (V x 3K) = Inv (V x V) . R (V x 3K)  on the 5th-generation tensor
// cores (tcgen05.mma, accumulator in tensor memory), as a drop-in for the SIMT mg_batch_dense_kernel (mg_kernels.cuh).
//
// One CTA (256 threads) computes a 128-row x 64-member tile: D (128 x 192, fp32) lives in 192 TMEM columns. Operands are
// staged by the CTA's threads (two stages: the next one is fetched and converted while the tensor core works) in shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core
// matrices; SBO = distance between core matrices along M/N, LBO = along K) and consumed by tcgen05.mma.kind::tf32 issued by
// one thread; tcgen05.commit + an mbarrier tell the CTA when a stage's operands may be overwritten.
// Precision: TF32 has a 10-bit mantissa, far too little for a preconditioner that should let CG finish in one or two
// iterations, so both operands are split hi + lo (hi = the nearest TF32 value, lo = the rest, rounded to TF32 as well) and
// three products are accumulated, hi.hi + lo.hi + hi.lo: ~1e-6 relative to sum |a b| (measured), close to fp32 (1e-7). Inv is split once at setup
// (tf32_split_kernel); R is split on the fly while it is converted from the fp64 CG residual.
#pragma once

#include "mg_kernels.cuh"
#include "tma_stage.cuh"

namespace arap {

constexpr int kTcM = 128, kTcMembers = 64, kTcN = kTcMembers * 3, kTcK = 32;     // tile: rows x members x k per stage
constexpr int kTcTmemCols = 256;                                                // >= kTcN, power of two
constexpr uint32_t kTcInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
// shared memory per stage (floats): A hi, A lo: kTcK/4 K-cores x 16 M-cores x 32 floats; B hi, B lo: kTcK/4 x 24 N-cores x 32
// (A's K-cores are 16 bytes further apart than they need to be -- the LBO is a free parameter of the descriptor -- so that the
// eight float4 stores of one row, one per K-core, fall into different banks)
constexpr int kTcAKcoreFloats = (kTcM / 8) * 32 + 4, kTcBKcoreFloats = (kTcN / 8) * 32;
constexpr int kTcAFloats = (kTcK / 4) * kTcAKcoreFloats, kTcBFloats = (kTcK / 4) * kTcBKcoreFloats;
constexpr size_t kTcSmemBytes = sizeof(float) * 2 * 2 * (size_t)(kTcAFloats + kTcBFloats) + 64;      // two stages + barriers

// nearest TF32 value of a (10-bit mantissa), as an fp32
__device__ __forceinline__ float tf32_round(float a) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a));
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(kBlock) tf32_split_kernel(size_t n, const float *__restrict__ in, float *__restrict__ hi, float *__restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = in[i];
    const float h = tf32_round(a);
    hi[i] = h;
    lo[i] = tf32_round(a - h);
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

constexpr int kTcThreads = 256;
constexpr int kTcStageFloats = 2 * (kTcAFloats + kTcBFloats);

__global__ void __launch_bounds__(kTcThreads, 1) mg_batch_dense_tc_kernel(int V, int ld, int K, const float *__restrict__ inv_hi,
                                                                         const float *__restrict__ inv_lo, const Vec3d *__restrict__ r,
                                                                         MgVec *__restrict__ z, const CgScalars *__restrict__ cg) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    if (cg->converged) return;
    float *stage_base = reinterpret_cast<float *>(tc_smem);                       // 2 stages of [A hi | A lo | B hi | B lo]
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage_base + 2 * kTcStageFloats);   // one mbarrier per stage
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int i0 = blockIdx.x * kTcM, m0 = blockIdx.y * kTcMembers;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_addr(tmem_slot)), "n"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { tma::mbar_init(&bar[0], 1); tma::mbar_init(&bar[1], 1); tma::fence_barrier_init(); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    constexpr uint32_t kALbo = kTcAKcoreFloats * 4, kBLbo = kTcBKcoreFloats * 4, kSbo = 128;       // bytes
    constexpr int kAItems = kTcM * (kTcK / 4) / kTcThreads;                                  // float4 pairs per thread: 4
    constexpr int kBItems = kTcN * (kTcK / 4) / kTcThreads;                                  // (column, K-core) pairs per thread: 6
    float4 a_h[kAItems], a_l[kAItems];
    double b_v[kBItems][4];
    // global -> registers (issued a stage ahead, so their latency hides behind the previous stage's MMAs)
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < kAItems; ++j) {
            const int t = tid + j * kTcThreads, row = t >> 3, kc = t & 7;                    // a row's 8 float4 are contiguous in global memory
            const int gi = i0 + row, gk = k0 + 4 * kc;
            const bool ok = gi < V && gk < ld;
            a_h[j] = ok ? __ldg(reinterpret_cast<const float4 *>(inv_hi + (size_t)gi * ld + gk)) : make_float4(0.f, 0.f, 0.f, 0.f);
            a_l[j] = ok ? __ldg(reinterpret_cast<const float4 *>(inv_lo + (size_t)gi * ld + gk)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const double *rd = reinterpret_cast<const double *>(r);
#pragma unroll
        for (int j = 0; j < kBItems; ++j) {
            const int t = tid + j * kTcThreads, kc = t / kTcN, n = t - kc * kTcN;            // consecutive threads: consecutive columns n = 3 m + c
            const int ml = n / 3, c = n - 3 * ml;
            const int gm = m0 + ml, gk = k0 + 4 * kc;
#pragma unroll
            for (int q = 0; q < 4; ++q) b_v[j][q] = (gm < K && gk + q < V) ? rd[((size_t)gm * V + gk + q) * 3 + c] : 0.0;
        }
    };
    // registers -> shared memory in the canonical K-major layout: the 16-byte chunk (row, K-core kc) of a tile sits at
    // kc * LBO + (row/8) * 128 + (row%8) * 16 = kc * LBO + row * 16 bytes
    auto stash = [&](int buf) {
        float *a_hi = stage_base + buf * kTcStageFloats, *a_lo = a_hi + kTcAFloats, *b_hi = a_lo + kTcAFloats, *b_lo = b_hi + kTcBFloats;
#pragma unroll
        for (int j = 0; j < kAItems; ++j) {
            const int t = tid + j * kTcThreads, row = t >> 3, kc = t & 7;
            const int off = kc * kTcAKcoreFloats + row * 4;
            *reinterpret_cast<float4 *>(a_hi + off) = a_h[j];
            *reinterpret_cast<float4 *>(a_lo + off) = a_l[j];
        }
#pragma unroll
        for (int j = 0; j < kBItems; ++j) {
            const int t = tid + j * kTcThreads, kc = t / kTcN, n = t - kc * kTcN;
            float h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float f = (float)b_v[j][q];
                h[q] = tf32_round(f);
                l[q] = tf32_round((float)(b_v[j][q] - (double)h[q]));
            }
            const int off = kc * kTcBKcoreFloats + n * 4;
            *reinterpret_cast<float4 *>(b_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(b_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
    };
    auto issue = [&](int buf, bool first) {         // one thread: 4 k-steps x (hi.hi + lo.hi + hi.lo), then commit to the stage's barrier
        const uint32_t a_hi_s = tma::smem_addr(stage_base + buf * kTcStageFloats), a_lo_s = a_hi_s + 4 * kTcAFloats;
        const uint32_t b_hi_s = a_lo_s + 4 * kTcAFloats, b_lo_s = b_hi_s + 4 * kTcBFloats;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < kTcK / 8; ++kk) {
            const uint64_t dah = umma_desc_kmajor(a_hi_s + kk * 2 * kALbo, kALbo, kSbo), dal = umma_desc_kmajor(a_lo_s + kk * 2 * kALbo, kALbo, kSbo);
            const uint64_t dbh = umma_desc_kmajor(b_hi_s + kk * 2 * kBLbo, kBLbo, kSbo), dbl = umma_desc_kmajor(b_lo_s + kk * 2 * kBLbo, kBLbo, kSbo);
            umma_tf32(tmem_d, dah, dbh, kTcInstrDesc, (first && kk == 0) ? 0u : 1u);
            umma_tf32(tmem_d, dal, dbh, kTcInstrDesc, 1u);
            umma_tf32(tmem_d, dah, dbl, kTcInstrDesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_addr(&bar[buf])) : "memory");
    };

    const int n_stages = (V + kTcK - 1) / kTcK;
    fetch(0);
    for (int s = 0; s < n_stages; ++s) {
        const int buf = s & 1;
        if (s >= 2) tma::mbar_wait(&bar[buf], (uint32_t)(((s >> 1) - 1) & 1));     // the MMAs of stage s-2 are done with this buffer
        stash(buf);
        if (s + 1 < n_stages) fetch((s + 1) * kTcK);                                // in flight while this stage's MMAs run
        tma::fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core's (async proxy) reads
        __syncthreads();
        if (tid == 0) issue(buf, s == 0);
    }
    // all MMAs done: the last commit covers every earlier one (they complete in issue order)
    tma::mbar_wait(&bar[(n_stages - 1) & 1], (uint32_t)(((n_stages - 1) >> 1) & 1));
    // ---- epilogue: TMEM lane = row of the tile, column = 3 * member + component. Warp w reads lanes 32 (w % 4) ..;
    //      warps 0-3 take the first half of the columns, warps 4-7 the second.
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lane_row = (warp & 3) * 32 + (tid & 31);
    const int gi = i0 + lane_row;
    const int c_begin = (warp >> 2) * (kTcN / 2);
    for (int c0 = c_begin; c0 < c_begin + kTcN / 2; c0 += 24) {                   // 8 members per round
        uint32_t v[24];
        const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
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
