// mg_kernels.cuh -- device side of the multigrid-preconditioned CG (global step, replaces the
// reference's SimplicialLDLT::solve, arap.h:418-421).
//
// One V(1,1)-cycle, damped Jacobi smoothing, applied to b_0 = r (the CG residual):
//   x_l = omega_l D_l^-1 b_l ; r_l = b_l - A_l x_l ; b_{l+1} = R_l r_l ; ... ; x_L = A_L^-1 b_L (dense)
//   x_l += P_l x_{l+1} ; x_l += omega_l D_l^-1 (b_l - A_l x_l)
// The CG itself (operator, residual, dot products, solution) is fp64; the V-cycle is only a preconditioner and runs
// in fp32: its vectors are float4 (x,y,z,-) -- three right-hand sides per row, ONE aligned 16-byte gather per
// neighbour -- and its matrices are int32 + float. That halves the bytes of every V-cycle kernel; the CG still
// converges to the fp64 solution of the exactly-weighted system. HBM/L2-bound row-gather kernels.
#pragma once

#include "kernels.cuh"

#include <cooperative_groups.h>

namespace arap {

typedef float4 MgVec;      // (x, y, z, unused)

__device__ __forceinline__ float gather_gate(const MgVec (&a)[kSpmvChunk]) {
    float s = a[0].x;
#pragma unroll
    for (int u = 1; u < kSpmvChunk; ++u) s += a[u].x;
    return 0.0f * s;
}

// (|z| / length)^8 of one vertex: the summand of the position-error stopping criterion (cg_finalize, CG_STAGE_RHO)
__device__ __forceinline__ double z_norm8(const MgVec &z, double inv_len2) {
    const double q = ((double)z.x * z.x + (double)z.y * z.y + (double)z.z * z.z) * inv_len2;
    const double q2 = q * q;
    return q2 * q2;
}

// ---- fine level (matrix-free): (A x)_i = sum_j w_ij (x_i - x_j) on free rows ---------------------------
__device__ __forceinline__ float3 fine_apply_row(int i, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                 const float *__restrict__ weight, const MgVec *__restrict__ x) {
    constexpr int CH = kSpmvChunk;
    const int k0 = rowptr[i], k1 = rowptr[i + 1];
    const MgVec xi = x[i];
    float3 out = {0.f, 0.f, 0.f};
    for (int k = k0; k < k1; k += CH) {
        int j[CH];
        float w[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const bool valid = k + u < k1;
            j[u] = valid ? __ldg(&colidx[k + u]) : i;
            w[u] = valid ? __ldg(&weight[k + u]) : 0.f;
        }
        MgVec xj[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) xj[u] = __ldg(&x[j[u]]);
        const float gate = gather_gate(xj);
#pragma unroll
        for (int u = 0; u < CH; ++u) w[u] += gate;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            out.x += w[u] * (xi.x - xj[u].x); out.y += w[u] * (xi.y - xj[u].y); out.z += w[u] * (xi.z - xj[u].z);
        }
    }
    return out;
}

// r0 = b - A x0 on free rows (0 elsewhere); b is the CG residual (fp64)
__global__ void __launch_bounds__(kBlock) mg_fine_residual_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                  const float *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                                  const Vec3d *__restrict__ b, const MgVec *__restrict__ x,
                                                                  MgVec *__restrict__ r, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    MgVec out = {0.f, 0.f, 0.f, 0.f};
    if (free_mask[i]) {
        const float3 ax = fine_apply_row(i, rowptr, colidx, weight, x);
        const Vec3d bi = b[i];
        out.x = (float)bi.x - ax.x; out.y = (float)bi.y - ax.y; out.z = (float)bi.z - ax.z;
    }
    r[i] = out;
}

// z = x + omega D^-1 (b - A x) on free rows; fused rho_new = b . z (b is the CG residual) -> beta.
__global__ void __launch_bounds__(kBlock) mg_fine_postsmooth_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                    const float *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                                    const double *__restrict__ inv_diag, double omega,
                                                                    const Vec3d *__restrict__ b, const MgVec *__restrict__ x,
                                                                    MgVec *__restrict__ z, double *__restrict__ partials,
                                                                    unsigned *__restrict__ counter, CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    double red[4] = {0, 0, 0, 0};
    const double inv_len2 = cg->inv_len2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        MgVec out = {0.f, 0.f, 0.f, 0.f};
        if (free_mask[i]) {
            const float3 ax = fine_apply_row(i, rowptr, colidx, weight, x);
            const Vec3d bi = b[i];
            const MgVec xi = x[i];
            const float s = (float)(omega * inv_diag[i]);
            out.x = xi.x + s * ((float)bi.x - ax.x); out.y = xi.y + s * ((float)bi.y - ax.y); out.z = xi.z + s * ((float)bi.z - ax.z);
            red[0] += bi.x * (double)out.x; red[1] += bi.y * (double)out.y; red[2] += bi.z * (double)out.z;
            red[3] += z_norm8(out, inv_len2);
        }
        z[i] = out;
    }
    double total[4];
    if (grid_sum_last_block<4>(red, partials, counter, total)) cg_finish_reduction<4>(cg, CG_STAGE_RHO, total);
}

// ---- generic CSR levels -------------------------------------------------------------------------------
// LANES (a power of two <= 32) consecutive threads share one row and reduce with shuffles, so that
// rows of ~10-25 entries (coarse operators, restriction) still spread over enough threads to fill the GPU.
template <int LANES>
__device__ __forceinline__ float3 csr_apply_row(int row, bool row_valid, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                const float *__restrict__ val, const MgVec *__restrict__ x) {
    float3 out = {0.f, 0.f, 0.f};
    const int sub = threadIdx.x & (LANES - 1);
    if (row_valid) {
        const int k0 = rowptr[row], k1 = rowptr[row + 1];
        for (int k = k0 + sub; k < k1; k += LANES) {
            const int j = __ldg(&colidx[k]);
            const float a = __ldg(&val[k]);
            const MgVec xj = __ldg(&x[j]);
            out.x += a * xj.x; out.y += a * xj.y; out.z += a * xj.z;
        }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        out.x += __shfl_down_sync(0xffffffffu, out.x, o, LANES);
        out.y += __shfl_down_sync(0xffffffffu, out.y, o, LANES);
        out.z += __shfl_down_sync(0xffffffffu, out.z, o, LANES);
    }
    return out;     // complete in the row's lane 0
}

// r = b - A x
template <int LANES>
__global__ void __launch_bounds__(kBlock) mg_csr_residual_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                 const float *__restrict__ val, const MgVec *__restrict__ b,
                                                                 const MgVec *__restrict__ x, MgVec *__restrict__ r,
                                                                 const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const float3 ax = csr_apply_row<LANES>(i, i < n, rowptr, colidx, val, x);
    if (i < n && (threadIdx.x & (LANES - 1)) == 0) {
        const MgVec bi = b[i];
        r[i] = MgVec{bi.x - ax.x, bi.y - ax.y, bi.z - ax.z, 0.f};
    }
}

// b_c = R r_f ; x_c = omega_c D_c^-1 b_c   (restriction fused with the coarse level's pre-smoothing from a zero guess)
template <int LANES>
__global__ void __launch_bounds__(kBlock) mg_restrict_presmooth_kernel(int nc, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                       const float *__restrict__ val, const MgVec *__restrict__ r_fine,
                                                                       const float *__restrict__ inv_diag_c, float omega_c,
                                                                       MgVec *__restrict__ b_c, MgVec *__restrict__ x_c,
                                                                       const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const float3 bc = csr_apply_row<LANES>(i, i < nc, rowptr, colidx, val, r_fine);
    if (i < nc && (threadIdx.x & (LANES - 1)) == 0) {
        b_c[i] = MgVec{bc.x, bc.y, bc.z, 0.f};
        const float s = omega_c * inv_diag_c[i];
        x_c[i] = MgVec{s * bc.x, s * bc.y, s * bc.z, 0.f};
    }
}

// x += P x_c   (P rows hold ~1-4 entries: one thread per row)
__global__ void __launch_bounds__(kBlock) mg_prolong_add_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                const float *__restrict__ val, const MgVec *__restrict__ x_c,
                                                                MgVec *__restrict__ x, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (rowptr[i] == rowptr[i + 1]) return;
    const float3 c = csr_apply_row<1>(i, true, rowptr, colidx, val, x_c);
    MgVec xi = x[i];
    xi.x += c.x; xi.y += c.y; xi.z += c.z;
    x[i] = xi;
}

// x_out = x + omega D^-1 (b - A x)
template <int LANES>
__global__ void __launch_bounds__(kBlock) mg_csr_postsmooth_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                   const float *__restrict__ val, const float *__restrict__ inv_diag,
                                                                   float omega, const MgVec *__restrict__ b, const MgVec *__restrict__ x,
                                                                   MgVec *__restrict__ x_out, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const float3 ax = csr_apply_row<LANES>(i, i < n, rowptr, colidx, val, x);
    if (i < n && (threadIdx.x & (LANES - 1)) == 0) {
        const MgVec bi = b[i], xi = x[i];
        const float s = omega * inv_diag[i];
        x_out[i] = MgVec{xi.x + s * (bi.x - ax.x), xi.y + s * (bi.y - ax.y), xi.z + s * (bi.z - ax.z), 0.f};
    }
}

// coarsest level: x = A^-1 b with the dense inverse; one warp per row. The right-hand side (n float4, <= 32 KB) is staged
// in shared memory once per CTA: read per warp from L2 it was 4x the traffic of the matrix itself (13 us at 1170 rows).
__global__ void __launch_bounds__(kBlock) mg_dense_solve_kernel(int n, int ld, const float *__restrict__ inv, const MgVec *__restrict__ b,
                                                                MgVec *__restrict__ x, const CgScalars *__restrict__ cg) {
    extern __shared__ __align__(16) unsigned char dense_smem[];
    MgVec *sb = reinterpret_cast<MgVec *>(dense_smem);                   // ld entries, zero beyond n
    if (cg->converged) return;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) sb[c] = c < n ? b[c] : MgVec{0.f, 0.f, 0.f, 0.f};
    __syncthreads();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    // rows are padded to ld (a multiple of 4) floats: 16-byte loads, two in flight per lane
    const float4 *arow = reinterpret_cast<const float4 *>(inv + (size_t)row * ld);
    const int n4 = ld >> 2;
    float s0 = 0, s1 = 0, s2 = 0;
    for (int q = lane; q < n4; q += 64) {
        const float4 a0 = __ldg(&arow[q]);
        const bool two = q + 32 < n4;
        const float4 a1 = two ? __ldg(&arow[q + 32]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const MgVec *b0 = sb + 4 * q, *b1 = sb + (two ? 4 * (q + 32) : 0);
        s0 += a0.x * b0[0].x + a0.y * b0[1].x + a0.z * b0[2].x + a0.w * b0[3].x;
        s1 += a0.x * b0[0].y + a0.y * b0[1].y + a0.z * b0[2].y + a0.w * b0[3].y;
        s2 += a0.x * b0[0].z + a0.y * b0[1].z + a0.z * b0[2].z + a0.w * b0[3].z;
        s0 += a1.x * b1[0].x + a1.y * b1[1].x + a1.z * b1[2].x + a1.w * b1[3].x;
        s1 += a1.x * b1[0].y + a1.y * b1[1].y + a1.z * b1[2].y + a1.w * b1[3].y;
        s2 += a1.x * b1[0].z + a1.y * b1[1].z + a1.z * b1[2].z + a1.w * b1[3].z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) x[row] = MgVec{s0, s1, s2, 0.f};
}

// ---- batches: one member's dense inverse applied to every member (a GEMM) --------------------------------------------
// arap_batch_*: K members share topology, rest pose and the constrained SET, so they share the operator L. For members of
// up to 2048 vertices the preconditioner is therefore ONE dense fp32 inverse (V x V, inverted once on the device) applied
// to all K residuals at once:  Z (V x 3K) = Inv (V x V) . R (V x 3K)  -- the one place on this path that is a dense
// contraction. SIMT fp32 tiles (128 rows x 32 members x 16 k), 8 x 6 accumulators per thread; the right-hand side is read
// straight from the fp64 CG residual (member-major Vec3d) and converted on the fly, the result is written as float4.
constexpr int kBgM = 128, kBgMembers = 32, kBgK = 16, kBgN = kBgMembers * 3;
__global__ void __launch_bounds__(256) mg_batch_dense_kernel(int V, int ld, int K, const float *__restrict__ inv, const Vec3d *__restrict__ r,
                                                             MgVec *__restrict__ z, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    __shared__ __align__(16) float As[2][kBgK][kBgM];        // As[k][i]
    __shared__ __align__(16) float Bs[2][kBgK][kBgN];        // Bs[k][member * 3 + c]
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * kBgM, m0 = blockIdx.y * kBgMembers;
    const int ti = tid & 15, tn = tid >> 4;                  // thread tile: rows ti*8 .. +8, members tn*2 .. +2
    float acc[8][6];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) acc[a][b] = 0.f;
    // global -> registers for one k-tile
    float4 a_reg[2];
    Vec3d b_reg[2];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = (tid >> 2) + 64 * j, q = tid & 3;                // 128 rows x 4 float4
            const int gi = i0 + row, gk = k0 + 4 * q;
            a_reg[j] = (gi < V && gk < ld) ? __ldg(reinterpret_cast<const float4 *>(inv + (size_t)gi * ld + gk)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int ml = (tid >> 4) + 16 * j, kl = tid & 15;              // 32 members x 16 k
            const int gm = m0 + ml, kk = k0 + kl;
            b_reg[j] = (gm < K && kk < V) ? r[(size_t)gm * V + kk] : Vec3d{0.0, 0.0, 0.0};
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = (tid >> 2) + 64 * j, q = tid & 3;
            As[buf][4 * q + 0][row] = a_reg[j].x; As[buf][4 * q + 1][row] = a_reg[j].y;
            As[buf][4 * q + 2][row] = a_reg[j].z; As[buf][4 * q + 3][row] = a_reg[j].w;
            const int ml = (tid >> 4) + 16 * j, kl = tid & 15;
            Bs[buf][kl][3 * ml + 0] = (float)b_reg[j].x; Bs[buf][kl][3 * ml + 1] = (float)b_reg[j].y; Bs[buf][kl][3 * ml + 2] = (float)b_reg[j].z;
        }
    };
    const int n_tiles = (V + kBgK - 1) / kBgK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) load_tile((t + 1) * kBgK);
#pragma unroll
        for (int k = 0; k < kBgK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ti * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ti * 8 + 4]);
            const float2 b0 = *reinterpret_cast<const float2 *>(&Bs[buf][k][tn * 6]);
            const float2 b1 = *reinterpret_cast<const float2 *>(&Bs[buf][k][tn * 6 + 2]);
            const float2 b2 = *reinterpret_cast<const float2 *>(&Bs[buf][k][tn * 6 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) acc[a][b] += av[a] * bv[b];
        }
        if (t + 1 < n_tiles) store_tile(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
        const int gm = m0 + tn * 2 + mm;
        if (gm >= K) continue;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int gi = i0 + ti * 8 + a;
            if (gi < V) z[(size_t)gm * V + gi] = MgVec{acc[a][3 * mm], acc[a][3 * mm + 1], acc[a][3 * mm + 2], 0.f};
        }
    }
}

// ---- the tail of the V-cycle in ONE kernel -----------------------------------------------------------------------
// On the coarse levels (a few thousand rows and fewer) every kernel above is pure launch latency: ~11 dependent launches of
// ~3-5 us each per CG iteration, a quarter of the iteration at 1M vertices. This kernel runs the whole tail -- restriction
// into the first small level, residual / restriction down to the coarsest level, the dense solve, prolongation and
// post-smoothing back up, and the prolongation out of the tail -- as phases of one thread-block CLUSTER separated by the
// hardware cluster barrier (~0.5 us instead of a kernel boundary). All vectors of the tail live in global memory / L2 and
// are read with ld.global.cg (never through L1: another CTA of the cluster wrote them in the previous phase).
constexpr int kTailMaxLevels = 12;
constexpr int kTailThreads = 1024;

struct MgTailLevel {
    int n, a_lanes, r_lanes;
    float omega;
    const int *a_rowptr, *a_colidx;
    const float *a_val, *inv_diag;
    const int *p_rowptr, *p_colidx;          // P: rows of this level -> columns of the next (coarser) level
    const float *p_val;
    const int *r_rowptr, *r_colidx;          // R: rows of the next level <- columns of this level
    const float *r_val;
    MgVec *b, *x, *x2, *r;
};
struct MgTailArgs {
    int n_levels;                            // lv[0] = the parent of the tail (only its r, x, P, R are used), lv[n_levels-1] = coarsest
    int dense;                               // coarsest level: dense inverse (else one more damped-Jacobi step)
    int coarse_ld;                           // row stride of the dense inverse (n rounded up to a multiple of 4)
    const float *coarse_inv;
    MgTailLevel lv[kTailMaxLevels];
};

__device__ __forceinline__ MgVec tail_load(const MgVec *p) { return __ldcg(p); }

// (M v)_row with `lanes` threads per row; complete in the row's lane 0. All threads of a warp must call it together.
__device__ __forceinline__ float3 tail_apply_row(int row, bool valid, int lanes, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                 const float *__restrict__ val, const MgVec *v) {
    float3 out = {0.f, 0.f, 0.f};
    if (valid) {
        const int k1 = __ldg(&rowptr[row + 1]);
        for (int k = __ldg(&rowptr[row]) + (int)(threadIdx.x & (lanes - 1)); k < k1; k += lanes) {
            const float a = __ldg(&val[k]);
            const MgVec xj = tail_load(&v[__ldg(&colidx[k])]);
            out.x += a * xj.x; out.y += a * xj.y; out.z += a * xj.z;
        }
    }
    for (int o = lanes >> 1; o > 0; o >>= 1) {
        out.x += __shfl_down_sync(0xffffffffu, out.x, o, lanes);
        out.y += __shfl_down_sync(0xffffffffu, out.y, o, lanes);
        out.z += __shfl_down_sync(0xffffffffu, out.z, o, lanes);
    }
    return out;
}

__global__ void __launch_bounds__(kTailThreads, 1) mg_tail_kernel(const MgTailArgs args, const CgScalars *__restrict__ cg) {
    namespace cgrp = cooperative_groups;
    cgrp::cluster_group cluster = cgrp::this_cluster();
    if (cg->converged) return;                                  // uniform over the cluster: nobody reaches a barrier
    const int nt = (int)cluster.num_blocks() * kTailThreads;
    const int tid = (int)cluster.block_rank() * kTailThreads + (int)threadIdx.x;
    const int L = args.n_levels;

    auto restrict_presmooth = [&](const MgTailLevel &f, const MgTailLevel &c) {      // b_c = R r_f ; x_c = omega_c D_c^-1 b_c
        const int lanes = f.r_lanes, per = nt / lanes;
        for (int base = 0; base < c.n; base += per) {
            const int row = base + tid / lanes;
            const float3 bc = tail_apply_row(row, row < c.n, lanes, f.r_rowptr, f.r_colidx, f.r_val, f.r);
            if (row < c.n && (threadIdx.x & (lanes - 1)) == 0) {
                c.b[row] = MgVec{bc.x, bc.y, bc.z, 0.f};
                const float s = c.omega * __ldg(&c.inv_diag[row]);
                c.x[row] = MgVec{s * bc.x, s * bc.y, s * bc.z, 0.f};
            }
        }
    };
    auto residual = [&](const MgTailLevel &f) {                                       // r = b - A x
        const int lanes = f.a_lanes, per = nt / lanes;
        for (int base = 0; base < f.n; base += per) {
            const int row = base + tid / lanes;
            const float3 ax = tail_apply_row(row, row < f.n, lanes, f.a_rowptr, f.a_colidx, f.a_val, f.x);
            if (row < f.n && (threadIdx.x & (lanes - 1)) == 0) {
                const MgVec bi = tail_load(&f.b[row]);
                f.r[row] = MgVec{bi.x - ax.x, bi.y - ax.y, bi.z - ax.z, 0.f};
            }
        }
    };
    auto postsmooth = [&](const MgTailLevel &f) {                                     // x2 = x + omega D^-1 (b - A x)
        const int lanes = f.a_lanes, per = nt / lanes;
        for (int base = 0; base < f.n; base += per) {
            const int row = base + tid / lanes;
            const float3 ax = tail_apply_row(row, row < f.n, lanes, f.a_rowptr, f.a_colidx, f.a_val, f.x);
            if (row < f.n && (threadIdx.x & (lanes - 1)) == 0) {
                const MgVec bi = tail_load(&f.b[row]), xi = tail_load(&f.x[row]);
                const float s = f.omega * __ldg(&f.inv_diag[row]);
                f.x2[row] = MgVec{xi.x + s * (bi.x - ax.x), xi.y + s * (bi.y - ax.y), xi.z + s * (bi.z - ax.z), 0.f};
            }
        }
    };
    auto prolong_add = [&](const MgTailLevel &f, const MgTailLevel &c) {              // x_f += P x2_c (P rows are short: one thread each)
        for (int row = tid; row < f.n; row += nt) {
            const int k0 = __ldg(&f.p_rowptr[row]), k1 = __ldg(&f.p_rowptr[row + 1]);
            if (k0 == k1) continue;
            MgVec xi = tail_load(&f.x[row]);
            for (int k = k0; k < k1; ++k) {
                const float a = __ldg(&f.p_val[k]);
                const MgVec xc = tail_load(&c.x2[__ldg(&f.p_colidx[k])]);
                xi.x += a * xc.x; xi.y += a * xc.y; xi.z += a * xc.z;
            }
            f.x[row] = xi;
        }
    };

    // down
    restrict_presmooth(args.lv[0], args.lv[1]);
    cluster.sync();
    for (int l = 1; l + 1 < L; ++l) {
        residual(args.lv[l]);
        cluster.sync();
        restrict_presmooth(args.lv[l], args.lv[l + 1]);
        cluster.sync();
    }
    // coarsest
    {
        const MgTailLevel &c = args.lv[L - 1];
        if (args.dense) {                                         // x2 = A^-1 b, one warp per row
            const int lane = threadIdx.x & 31, n = c.n;
            for (int row = tid >> 5; row < n; row += nt >> 5) {
                float s0 = 0, s1 = 0, s2 = 0;
                for (int k = lane; k < n; k += 32) {
                    const float a = __ldg(&args.coarse_inv[(size_t)row * args.coarse_ld + k]);
                    const MgVec bk = tail_load(&c.b[k]);
                    s0 += a * bk.x; s1 += a * bk.y; s2 += a * bk.z;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s0 += __shfl_down_sync(0xffffffffu, s0, o);
                    s1 += __shfl_down_sync(0xffffffffu, s1, o);
                    s2 += __shfl_down_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) c.x2[row] = MgVec{s0, s1, s2, 0.f};
            }
        } else {
            postsmooth(c);
        }
    }
    cluster.sync();
    // up
    for (int l = L - 2; l >= 1; --l) {
        prolong_add(args.lv[l], args.lv[l + 1]);
        cluster.sync();
        postsmooth(args.lv[l]);
        cluster.sync();
    }
    prolong_add(args.lv[0], args.lv[1]);
}

// ---- dense inverse of the coarsest operator on the device (setup, once per hierarchy) ---------------------------------
// The host inverts coarsest levels of up to a few hundred rows; with up to 2048 rows the hierarchy is one or two levels
// shorter (4-8 launches less per V-cycle), but O(n^3) scalar host code would take seconds. In-place Gauss-Jordan without
// pivoting in fp64 (the operator is SPD after the same tiny diagonal shift the host code applies): two launches per pivot.
__global__ void __launch_bounds__(kBlock) dense_from_csr_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                const double *__restrict__ val, double shift, double *__restrict__ M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double diag = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        M[(size_t)i * n + colidx[k]] += val[k];
        if (colidx[k] == i) diag += val[k];
    }
    if (diag == 0.0) M[(size_t)i * n + i] = 1.0;            // empty row: identity
    else M[(size_t)i * n + i] += shift;
}
// pivot c, part 1 (one CTA): save column c, scale row c by 1/pivot, put 1/pivot on the diagonal
__global__ void __launch_bounds__(1024) gj_pivot_kernel(int n, int c, double *__restrict__ M, double *__restrict__ colbuf, int *__restrict__ bad) {
    __shared__ double dinv;
    if (threadIdx.x == 0) {
        const double p = M[(size_t)c * n + c];
        if (!(p > 0.0) || !(p < 1e300)) { *bad = 1; dinv = 0.0; }
        else dinv = 1.0 / p;
    }
    __syncthreads();
    const double d = dinv;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        colbuf[k] = M[(size_t)k * n + c];
        const double v = M[(size_t)c * n + k];
        M[(size_t)c * n + k] = (k == c) ? d : v * d;
    }
}
// pivot c, part 2: every other row r: M[r][k] -= f M[c][k] (k != c), M[r][c] = -f / pivot, f = the saved M[r][c]
__global__ void __launch_bounds__(kBlock) gj_update_kernel(int n, int c, double *__restrict__ M, const double *__restrict__ colbuf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (k >= n || r == c) return;
    const double f = colbuf[r];
    if (f == 0.0) return;
    const double rc = M[(size_t)c * n + k];                 // row c is already scaled; its entry at k == c is 1/pivot
    M[(size_t)r * n + k] = (k == c) ? -f * rc : M[(size_t)r * n + k] - f * rc;
}
__global__ void __launch_bounds__(kBlock) dense_to_float_kernel(int n, int ld, const double *__restrict__ in, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over n x ld, padding columns = 0
    if (i >= (size_t)n * ld) return;
    const int r = (int)(i / ld), c = (int)(i - (size_t)r * ld);
    out[i] = c < n ? (float)in[(size_t)r * n + c] : 0.f;
}

// single-level hierarchies (tiny meshes): the V-cycle input/output live in fp64 CG vectors
__global__ void __launch_bounds__(kBlock) mg_to_float_kernel(int n, const Vec3d *__restrict__ in, MgVec *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const Vec3d v = in[i]; out[i] = MgVec{(float)v.x, (float)v.y, (float)v.z, 0.f}; }
}
// z = omega D^-1 b (plain damped Jacobi; fallback when a single-level hierarchy has no dense inverse)
__global__ void __launch_bounds__(kBlock) mg_jacobi_kernel(int n, const float *__restrict__ inv_diag, float omega, const MgVec *__restrict__ b,
                                                           MgVec *__restrict__ z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const MgVec v = b[i]; const float s = omega * inv_diag[i]; z[i] = MgVec{s * v.x, s * v.y, s * v.z, 0.f}; }
}
// rho_new = r . z with z in fp32 (used when the last V-cycle kernel cannot fuse the dot product)
__global__ void __launch_bounds__(kBlock) cg_dot_rho_f_kernel(int n, const Vec3d *__restrict__ r, const MgVec *__restrict__ z,
                                                              double *__restrict__ partials, unsigned *__restrict__ counter,
                                                              CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    double red[4] = {0, 0, 0, 0};
    const double inv_len2 = cg->inv_len2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Vec3d ri = r[i];
        const MgVec zi = z[i];
        red[0] += ri.x * (double)zi.x; red[1] += ri.y * (double)zi.y; red[2] += ri.z * (double)zi.z;
        red[3] += z_norm8(zi, inv_len2);
    }
    double total[4];
    if (grid_sum_last_block<4>(red, partials, counter, total)) cg_finish_reduction<4>(cg, CG_STAGE_RHO, total);
}

__device__ __forceinline__ double pick3(int c, double a0, double a1, double a2) { return c == 0 ? a0 : (c == 1 ? a1 : a2); }

// ---- CG pieces for a general preconditioner -------------------------------------------------------------
// The CG vectors are (x,y,z) triples, i.e. flat arrays of 3V doubles. The update and direction kernels are
// purely element-wise, so they stream those flat arrays with one coalesced 16-byte access per thread and array
// (element e belongs to vertex e/3, coordinate e%3) instead of three strided 8-byte accesses per vertex.
//
// x += alpha d ; r -= alpha Ad ; x0 = omega_0 D^-1 r (the V-cycle's pre-smoothed fine iterate) ; |r|^2 -> convergence
__global__ void __launch_bounds__(kBlock) cg_update_mg_kernel(int n3, const double *__restrict__ inv_diag, double omega0,
                                                              const double *__restrict__ d, const double *__restrict__ ad,
                                                              double *__restrict__ x, double *__restrict__ r, float *__restrict__ x0,
                                                              double *__restrict__ partials, unsigned *__restrict__ counter,
                                                              CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    double red[1] = {0};
    const double al0 = cg->alpha[0], al1 = cg->alpha[1], al2 = cg->alpha[2];
    for (int e = 2 * (blockIdx.x * blockDim.x + threadIdx.x); e < n3; e += 2 * gridDim.x * blockDim.x) {
        if (e + 1 < n3) {
            const double2 dv = *reinterpret_cast<const double2 *>(d + e), av = *reinterpret_cast<const double2 *>(ad + e);
            double2 xv = *reinterpret_cast<const double2 *>(x + e), rv = *reinterpret_cast<const double2 *>(r + e);
            const int c0 = e % 3, c1 = (e + 1) % 3;
            const double a0 = pick3(c0, al0, al1, al2), a1 = pick3(c1, al0, al1, al2);
            xv.x += a0 * dv.x; xv.y += a1 * dv.y;
            rv.x -= a0 * av.x; rv.y -= a1 * av.y;
            *reinterpret_cast<double2 *>(x + e) = xv;
            *reinterpret_cast<double2 *>(r + e) = rv;
            const int v0 = e / 3, v1 = (e + 1) / 3;
            x0[4 * v0 + c0] = (float)(omega0 * inv_diag[v0] * rv.x);          // x0 is a float4 per vertex: element (v, c) at 4 v + c
            x0[4 * v1 + c1] = (float)(omega0 * inv_diag[v1] * rv.y);
            red[0] += rv.x * rv.x + rv.y * rv.y;
        } else {
            const double a0 = pick3(e % 3, al0, al1, al2);
            const double xv = x[e] + a0 * d[e], rv = r[e] - a0 * ad[e];
            x[e] = xv; r[e] = rv;
            x0[4 * (e / 3) + e % 3] = (float)(omega0 * inv_diag[e / 3] * rv);
            red[0] += rv * rv;
        }
    }
    double total[1];
    if (grid_sum_last_block<1>(red, partials, counter, total)) cg_finish_reduction<1>(cg, CG_STAGE_UPDATE_MG, total);
}

// d = z + beta d
__global__ void __launch_bounds__(kBlock) cg_direction_mg_kernel(int n3, const float *__restrict__ z, double *__restrict__ d,
                                                                 const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int e = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (e >= n3) return;
    const double b0 = cg->beta[0], b1 = cg->beta[1], b2 = cg->beta[2];
    if (e + 1 < n3) {
        const int c0 = e % 3, c1 = (e + 1) % 3;
        double2 dv = *reinterpret_cast<const double2 *>(d + e);
        dv.x = (double)z[4 * (e / 3) + c0] + pick3(c0, b0, b1, b2) * dv.x;      // z is a float4 per vertex
        dv.y = (double)z[4 * ((e + 1) / 3) + c1] + pick3(c1, b0, b1, b2) * dv.y;
        *reinterpret_cast<double2 *>(d + e) = dv;
    } else {
        d[e] = (double)z[4 * (e / 3) + e % 3] + pick3(e % 3, b0, b1, b2) * d[e];
    }
}

}  // namespace arap
