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

// (|z| / length)^8 of one vertex: the summand of the position-error stopping criterion (cg_finalize, CG_STAGE_GAMMA)
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
    pdl_enter();
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
    pdl_enter();
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
    if (grid_sum_last_block<4>(red, partials, counter, total)) cg_finish_reduction<4>(cg, CG_STAGE_GAMMA, total, 0);
}

// ---- generic CSR levels -------------------------------------------------------------------------------
// LANES (a power of two <= 32) consecutive threads share one row and reduce with shuffles, so that
// rows of ~10-25 entries (coarse operators, restriction) still spread over enough threads to fill the GPU.
// These kernels are a few microseconds of dependent round trips each (rowptr -> colidx/val -> x gather), so the part of
// the chain that does not depend on the previous kernel -- the matrix, which is constant across iterations -- is loaded
// BEFORE the programmatic-dependency wait (CsrRowHead) and overlaps the predecessor's tail.
constexpr int kCsrHead = 4;          // matrix entries per lane held in registers across the wait
struct CsrRowHead {
    int k0, k1;                      // this lane's first entry, the row's end
    int j[kCsrHead];
    float a[kCsrHead];
};
template <int LANES>
__device__ __forceinline__ CsrRowHead csr_row_head(int row, bool row_valid, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                   const float *__restrict__ val) {
    CsrRowHead h;
    h.k0 = 0; h.k1 = 0;
    if (row_valid) { h.k0 = __ldg(&rowptr[row]) + (int)(threadIdx.x & (LANES - 1)); h.k1 = __ldg(&rowptr[row + 1]); }
#pragma unroll
    for (int u = 0; u < kCsrHead; ++u) {
        const int k = h.k0 + u * LANES;
        const bool valid = k < h.k1;
        h.j[u] = valid ? __ldg(&colidx[k]) : 0;
        h.a[u] = valid ? __ldg(&val[k]) : 0.f;
    }
    return h;
}
template <int LANES>
__device__ __forceinline__ float3 csr_apply_row(const CsrRowHead &h, const int *__restrict__ colidx, const float *__restrict__ val,
                                                const MgVec *__restrict__ x) {
    float3 out = {0.f, 0.f, 0.f};
    MgVec xj[kCsrHead];
#pragma unroll
    for (int u = 0; u < kCsrHead; ++u) xj[u] = (h.k0 + u * LANES < h.k1) ? __ldg(&x[h.j[u]]) : MgVec{0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < kCsrHead; ++u) { out.x += h.a[u] * xj[u].x; out.y += h.a[u] * xj[u].y; out.z += h.a[u] * xj[u].z; }
    for (int k = h.k0 + kCsrHead * LANES; k < h.k1; k += LANES) {
        const int j = __ldg(&colidx[k]);
        const float a = __ldg(&val[k]);
        const MgVec v = __ldg(&x[j]);
        out.x += a * v.x; out.y += a * v.y; out.z += a * v.z;
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
    pdl_trigger();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const CsrRowHead h = csr_row_head<LANES>(i, i < n, rowptr, colidx, val);
    pdl_wait();
    if (cg->converged) return;
    const float3 ax = csr_apply_row<LANES>(h, colidx, val, x);
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
    pdl_trigger();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const CsrRowHead h = csr_row_head<LANES>(i, i < nc, rowptr, colidx, val);
    const float s = (i < nc) ? omega_c * __ldg(&inv_diag_c[i]) : 0.f;
    pdl_wait();
    if (cg->converged) return;
    const float3 bc = csr_apply_row<LANES>(h, colidx, val, r_fine);
    if (i < nc && (threadIdx.x & (LANES - 1)) == 0) {
        b_c[i] = MgVec{bc.x, bc.y, bc.z, 0.f};
        x_c[i] = MgVec{s * bc.x, s * bc.y, s * bc.z, 0.f};
    }
}

// x += P x_c   (P rows hold ~1-4 entries: one thread per row)
__global__ void __launch_bounds__(kBlock) mg_prolong_add_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                const float *__restrict__ val, const MgVec *__restrict__ x_c,
                                                                MgVec *__restrict__ x, const CgScalars *__restrict__ cg) {
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const CsrRowHead h = csr_row_head<1>(i, i < n, rowptr, colidx, val);
    pdl_wait();
    if (cg->converged) return;
    if (i >= n || h.k0 == h.k1) return;
    const float3 c = csr_apply_row<1>(h, colidx, val, x_c);
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
    pdl_trigger();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const CsrRowHead h = csr_row_head<LANES>(i, i < n, rowptr, colidx, val);
    const float s = (i < n) ? omega * __ldg(&inv_diag[i]) : 0.f;
    pdl_wait();
    if (cg->converged) return;
    const float3 ax = csr_apply_row<LANES>(h, colidx, val, x);
    if (i < n && (threadIdx.x & (LANES - 1)) == 0) {
        const MgVec bi = b[i], xi = x[i];
        x_out[i] = MgVec{xi.x + s * (bi.x - ax.x), xi.y + s * (bi.y - ax.y), xi.z + s * (bi.z - ax.z), 0.f};
    }
}

// coarsest level: x = A^-1 b with the dense inverse; one warp per row. The right-hand side (n float4, <= 32 KB) is staged
// in shared memory once per CTA: read per warp from L2 it was 4x the traffic of the matrix itself (13 us at 1170 rows).
// The matrix row (<= 2048 floats = 16 float4 per lane) is loaded in batches of 8 independent 16-byte loads per lane BEFORE
// any arithmetic: the kernel is one dependent L2 round trip long instead of one per 128 columns.
__global__ void __launch_bounds__(kBlock) mg_dense_solve_kernel(int n, int ld, const float *__restrict__ inv, const MgVec *__restrict__ b,
                                                                MgVec *__restrict__ x, const CgScalars *__restrict__ cg) {
    extern __shared__ __align__(16) unsigned char dense_smem[];
    MgVec *sb = reinterpret_cast<MgVec *>(dense_smem);                   // ld entries, zero beyond n
    pdl_trigger();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    // rows are padded to ld (a multiple of 4) floats
    const float4 *arow = reinterpret_cast<const float4 *>(inv + (size_t)(row < n ? row : 0) * ld);
    const int n4 = ld >> 2;
    constexpr int B = 8;
    float4 a[B];
#pragma unroll
    for (int u = 0; u < B; ++u) { const int q = lane + 32 * u; a[u] = q < n4 ? __ldg(&arow[q]) : make_float4(0.f, 0.f, 0.f, 0.f); }
    pdl_wait();                    // the matrix is constant: its first batch is in flight while the producer of b finishes
    if (cg->converged) return;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) sb[c] = c < n ? b[c] : MgVec{0.f, 0.f, 0.f, 0.f};
    __syncthreads();
    if (row >= n) return;
    float s0 = 0, s1 = 0, s2 = 0;
    for (int base = 0; base < n4; base += 32 * B) {
        if (base > 0) {
#pragma unroll
            for (int u = 0; u < B; ++u) { const int q = base + lane + 32 * u; a[u] = q < n4 ? __ldg(&arow[q]) : make_float4(0.f, 0.f, 0.f, 0.f); }
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int q = base + lane + 32 * u;
            if (q < n4) {
                const MgVec *bq = sb + 4 * q;
                s0 += a[u].x * bq[0].x + a[u].y * bq[1].x + a[u].z * bq[2].x + a[u].w * bq[3].x;
                s1 += a[u].x * bq[0].y + a[u].y * bq[1].y + a[u].z * bq[2].y + a[u].w * bq[3].y;
                s2 += a[u].x * bq[0].z + a[u].y * bq[1].z + a[u].z * bq[2].z + a[u].w * bq[3].z;
            }
        }
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
    pdl_enter();
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
    pdl_enter();
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

// ---- the whole coarse part of the V-cycle in ONE cooperative kernel ---------------------------------------------------
// Same phases as mg_tail_kernel, but run by a grid that fills the GPU (cudaLaunchCooperativeKernel: all CTAs co-resident)
// and separated by a grid-wide barrier in global memory instead of a cluster barrier: with ~150k threads the 100k-row level 1
// is processed as fast as by its own kernels, the small levels cost one barrier (~1.5 us) instead of a launch each, and the
// restriction out of / prolongation into the fine level move in here too. One CG iteration at 1M vertices then is 5 launches
// (fine residual, this kernel, fine post-smoothing, w = A z, the fused update) instead of 17.
struct GridBarrier { unsigned count; unsigned generation; };

__device__ __forceinline__ void grid_barrier(GridBarrier *b, unsigned n_blocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = *(volatile unsigned *)&b->generation;
        __threadfence();                                          // this CTA's writes before the arrival
        if (atomicAdd(&b->count, 1u) == n_blocks - 1) {
            b->count = 0;
            __threadfence();
            atomicAdd(&b->generation, 1u);
        } else {
            while (*(volatile unsigned *)&b->generation == gen) { }
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kBlock) mg_tail_grid_kernel(const MgTailArgs args, const CgScalars *__restrict__ cg, GridBarrier *bar) {
    extern __shared__ __align__(16) unsigned char tail_smem[];
    if (cg->converged) return;                                  // uniform over the grid: nobody reaches a barrier
    const int nt = (int)gridDim.x * kBlock;
    const int tid = (int)blockIdx.x * kBlock + (int)threadIdx.x;
    const unsigned nb = gridDim.x;
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
    grid_barrier(bar, nb);
    for (int l = 1; l + 1 < L; ++l) {
        residual(args.lv[l]);
        grid_barrier(bar, nb);
        restrict_presmooth(args.lv[l], args.lv[l + 1]);
        grid_barrier(bar, nb);
    }
    // coarsest
    {
        const MgTailLevel &c = args.lv[L - 1];
        if (args.dense) {                                         // x2 = A^-1 b: one warp per row, b staged in shared memory by the CTAs that have rows
            const int n = c.n, ld = args.coarse_ld, n4 = ld >> 2;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if ((int)blockIdx.x * kWarpsPerBlock < n) {
                MgVec *sb = reinterpret_cast<MgVec *>(tail_smem);
                for (int k = threadIdx.x; k < ld; k += kBlock) sb[k] = k < n ? tail_load(&c.b[k]) : MgVec{0.f, 0.f, 0.f, 0.f};
                __syncthreads();
                for (int row = (int)blockIdx.x * kWarpsPerBlock + warp; row < n; row += (int)gridDim.x * kWarpsPerBlock) {
                    const float4 *arow = reinterpret_cast<const float4 *>(args.coarse_inv + (size_t)row * ld);
                    float s0 = 0, s1 = 0, s2 = 0;
                    for (int base = 0; base < n4; base += 32 * 8) {
                        float4 a[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) { const int q = base + lane + 32 * u; a[u] = q < n4 ? __ldg(&arow[q]) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int q = base + lane + 32 * u;
                            if (q < n4) {
                                const MgVec *bq = sb + 4 * q;
                                s0 += a[u].x * bq[0].x + a[u].y * bq[1].x + a[u].z * bq[2].x + a[u].w * bq[3].x;
                                s1 += a[u].x * bq[0].y + a[u].y * bq[1].y + a[u].z * bq[2].y + a[u].w * bq[3].y;
                                s2 += a[u].x * bq[0].z + a[u].y * bq[1].z + a[u].z * bq[2].z + a[u].w * bq[3].z;
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s0 += __shfl_down_sync(0xffffffffu, s0, o);
                        s1 += __shfl_down_sync(0xffffffffu, s1, o);
                        s2 += __shfl_down_sync(0xffffffffu, s2, o);
                    }
                    if (lane == 0) c.x2[row] = MgVec{s0, s1, s2, 0.f};
                }
            }
        } else {
            postsmooth(c);
        }
    }
    grid_barrier(bar, nb);
    // up
    for (int l = L - 2; l >= 1; --l) {
        prolong_add(args.lv[l], args.lv[l + 1]);
        grid_barrier(bar, nb);
        postsmooth(args.lv[l]);
        grid_barrier(bar, nb);
    }
    prolong_add(args.lv[0], args.lv[1]);
}

// ---- dense inverse of the coarsest operator on the device (setup, once per hierarchy) ---------------------------------
// The host inverts coarsest levels of up to a few hundred rows; with up to 2048 rows the hierarchy is one or two levels
// shorter (4-8 launches less per V-cycle), but O(n^3) scalar host code would take seconds. In-place Gauss-Jordan without
// pivoting in fp64 (the operator is SPD after the same tiny diagonal shift the host code applies), blocked by panels.
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
// Blocked in-place Gauss-Jordan, panels of kGjB pivots, three launches per panel K = [k0, k0 + nb):
//   gj_panel_kernel   D = A[K,K]^-1 (one CTA, unblocked Gauss-Jordan in shared memory); C = the column panel A[:,K] saved
//   gj_row_kernel     A[K,J] <- D A[K,J] for the columns J outside K; A[K,K] <- D
//   gj_update_kernel  A[I,J] -= C[I,:] A[K,J] and A[I,K] <- -C[I,:] D for the rows I outside K
// (n / 32 * 3 launches instead of the 2 n of the pivot-by-pivot version: 111 instead of 2,340 at 1,170 rows.)
constexpr int kGjB = 32;
__global__ void __launch_bounds__(1024) gj_panel_kernel(int n, int k0, int nb, const double *__restrict__ M, double *__restrict__ D,
                                                        double *__restrict__ C, int *__restrict__ bad) {
    if (blockIdx.x > 0) {                                       // every other CTA: save the column panel C[r][c] = M[r][k0 + c]
        for (int t = (blockIdx.x - 1) * blockDim.x + threadIdx.x; t < n * nb; t += (gridDim.x - 1) * blockDim.x) {
            const int r = t / nb, c = t - r * nb;
            C[(size_t)r * kGjB + c] = M[(size_t)r * n + k0 + c];
        }
        return;
    }
    __shared__ double a[kGjB][kGjB + 1];
    __shared__ double colbuf[kGjB];
    __shared__ double dinv;
    const int r = threadIdx.x / kGjB, c = threadIdx.x % kGjB;   // 32 x 32 threads
    if (r < nb && c < nb) a[r][c] = M[(size_t)(k0 + r) * n + k0 + c];
    __syncthreads();
    for (int p = 0; p < nb; ++p) {
        if (threadIdx.x == 0) {
            const double piv = a[p][p];
            if (!(piv > 0.0) || !(piv < 1e300)) { *bad = 1; dinv = 0.0; }
            else dinv = 1.0 / piv;
        }
        if (r == 0 && c < nb) colbuf[c] = a[c][p];
        __syncthreads();
        const double d = dinv;
        if (r == p && c < nb) a[p][c] = (c == p) ? d : a[p][c] * d;
        __syncthreads();
        if (r < nb && c < nb && r != p) {
            const double f = colbuf[r];
            a[r][c] = (c == p) ? -f * a[p][c] : a[r][c] - f * a[p][c];
        }
        __syncthreads();
    }
    if (r < nb && c < nb) D[r * kGjB + c] = a[r][c];
}
__global__ void __launch_bounds__(kBlock) gj_row_kernel(int n, int k0, int nb, double *__restrict__ M, const double *__restrict__ D,
                                                        double *__restrict__ R) {
    __shared__ double d[kGjB][kGjB];
    for (int t = threadIdx.x; t < kGjB * kGjB; t += blockDim.x) d[t / kGjB][t % kGjB] = (t / kGjB < nb && t % kGjB < nb) ? D[t] : 0.0;
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;       // one column per thread
    if (j >= n) return;
    const bool inside = j >= k0 && j < k0 + nb;
    double col[kGjB];
#pragma unroll
    for (int q = 0; q < kGjB; ++q) col[q] = (q < nb) ? M[(size_t)(k0 + q) * n + j] : 0.0;
#pragma unroll 4
    for (int r = 0; r < nb; ++r) {
        double v;
        if (inside) v = d[r][j - k0];
        else {
            v = 0.0;
#pragma unroll
            for (int q = 0; q < kGjB; ++q) v += d[r][q] * col[q];
        }
        M[(size_t)(k0 + r) * n + j] = v;
        R[(size_t)r * n + j] = v;                                // the scaled row panel, read by the update kernel
    }
}
// 16 x 16 threads, each one entry of a 16 x 16 tile of the rows outside the panel
__global__ void __launch_bounds__(256) gj_update_kernel(int n, int k0, int nb, double *__restrict__ M, const double *__restrict__ C,
                                                        const double *__restrict__ R, const double *__restrict__ D) {
    __shared__ double cs[16][kGjB + 1];
    __shared__ double rs[kGjB][16 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    for (int t = threadIdx.x; t < 16 * kGjB; t += 256) {
        const int rr = t / kGjB, q = t % kGjB, gi = blockIdx.y * 16 + rr;
        cs[rr][q] = (gi < n && q < nb) ? C[(size_t)gi * kGjB + q] : 0.0;
        const int q2 = t / 16, cc = t % 16, gj = blockIdx.x * 16 + cc;
        const bool in_panel = gj >= k0 && gj < k0 + nb;
        // columns inside the panel use D (A[I,K] <- -C D); R holds D there as well, so one source serves both cases
        rs[q2][cc] = (gj < n && q2 < nb) ? (in_panel ? D[q2 * kGjB + (gj - k0)] : R[(size_t)q2 * n + gj]) : 0.0;
    }
    __syncthreads();
    if (i >= n || j >= n || (i >= k0 && i < k0 + nb)) return;
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kGjB; ++q) acc += cs[ty][q] * rs[q][tx];
    const bool in_panel = j >= k0 && j < k0 + nb;
    M[(size_t)i * n + j] = in_panel ? -acc : M[(size_t)i * n + j] - acc;
}
__global__ void __launch_bounds__(kBlock) dense_to_float_kernel(int n, int ld, const double *__restrict__ in, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over n x ld, padding columns = 0
    if (i >= (size_t)n * ld) return;
    const int r = (int)(i / ld), c = (int)(i - (size_t)r * ld);
    out[i] = c < n ? (float)in[(size_t)r * n + c] : 0.f;
}

// single-level hierarchies (tiny meshes): the V-cycle input/output live in fp64 CG vectors
__global__ void __launch_bounds__(kBlock) mg_to_float_kernel(int n, const Vec3d *__restrict__ in, MgVec *__restrict__ out) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const Vec3d v = in[i]; out[i] = MgVec{(float)v.x, (float)v.y, (float)v.z, 0.f}; }
}
// z = omega D^-1 b (plain damped Jacobi; fallback when a single-level hierarchy has no dense inverse)
__global__ void __launch_bounds__(kBlock) mg_jacobi_kernel(int n, const float *__restrict__ inv_diag, float omega, const MgVec *__restrict__ b,
                                                           MgVec *__restrict__ z) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const MgVec v = b[i]; const float s = omega * inv_diag[i]; z[i] = MgVec{s * v.x, s * v.y, s * v.z, 0.f}; }
}
// rho_new = r . z with z in fp32 (used when the last V-cycle kernel cannot fuse the dot product)
__global__ void __launch_bounds__(kBlock) cg_dot_rho_f_kernel(int n, const Vec3d *__restrict__ r, const MgVec *__restrict__ z,
                                                              double *__restrict__ partials, unsigned *__restrict__ counter,
                                                              CgScalars *__restrict__ cg) {
    pdl_enter();
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
    if (grid_sum_last_block<4>(red, partials, counter, total)) cg_finish_reduction<4>(cg, CG_STAGE_GAMMA, total, 0);
}

}  // namespace arap
