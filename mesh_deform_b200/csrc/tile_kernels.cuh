// tile_kernels.cuh -- the one-ring kernels of the ARAP iteration with the neighbourhood staged through shared memory.
//
// Every hot kernel of the iteration walks the one-ring CSR and gathers a per-vertex record of each neighbour (positions,
// rotations, a CG / multigrid vector). Done straight from global memory that is six scattered 16- or 32-byte requests per
// vertex and array: the L1 pipeline serves a warp's scattered request one cache line per cycle, and the kernels of round 1
// sat at 0.5-0.6 of the HBM roofline for exactly that reason (profiles/r01_g_spmv_stalls.txt) although their DRAM traffic was
// already minimal. Here the vertices are numbered in compact patches (Morton order of the rest pose, engine.cu) and the
// rows are processed in TILES of kTile consecutive rows by one CTA:
//
//   stage   the tile's own records are read with fully coalesced vector loads, the records of the tile's HALO -- the
//           distinct neighbours outside the tile, a precomputed list of ~0.3-0.6 x kTile vertices -- with one scattered
//           load each, all into shared memory                                   (global requests per row: ~1.5 instead of 7)
//   gather  each thread then walks its row with tile-LOCAL 16-bit column indices and gathers from shared memory.
//
// The tile structure (halo lists + local column indices) depends only on the topology and the vertex order and is built
// once per handle on the device (build_tiles_kernel). Arithmetic and summation order are exactly those of the kernels in
// kernels.cuh / mg_kernels.cuh (rows keep the reference's column order), so results are bit-identical to the untiled path,
// which remains the fallback for meshes whose tiles do not fit (a vertex of enormous valence) and for tiny meshes.
#pragma once

#include "kernels.cuh"
#include "mg_kernels.cuh"

namespace arap {

constexpr int kTile = kBlock;            // rows per tile = threads per CTA
constexpr int kTileHaloCap = 512;        // capacity of a tile's halo list: (256 + 512) staged records keep 3-4 CTAs of the fp64 kernels on an SM
constexpr int kTileEdgeCap = 4096;       // candidate slots while building a tile (entries with a column outside the tile)
constexpr int kTileChunk = 2;            // fp64 kernels: row entries whose index / weight loads are batched

// One CTA per tile: collect the columns outside [t0, t0 + kTile), sort + unique them (the tile's halo, ascending), and
// rewrite every column index of the tile's rows as a local index: j - t0 for own rows, kTile + rank in the halo otherwise.
// A tile whose halo does not fit gets count -1 (and *bad is raised): the engine then keeps the untiled kernels.
__global__ void __launch_bounds__(kBlock) build_tiles_kernel(int n_rows, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                             int *__restrict__ tile_halo, int *__restrict__ tile_halo_count,
                                                             unsigned short *__restrict__ tile_colidx, int *__restrict__ max_halo,
                                                             int *__restrict__ bad) {
    __shared__ int ext[kTileEdgeCap];
    __shared__ int uniq[kTileHaloCap];
    __shared__ int n_ext, n_uniq;
    const int tile = blockIdx.x;
    const int t0 = tile * kTile, t1 = min(n_rows, t0 + kTile);
    const int k0 = rowptr[t0], k1 = rowptr[t1];
    if (threadIdx.x == 0) { n_ext = 0; n_uniq = 0; }
    __syncthreads();
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        const int j = colidx[k];
        if (j < t0 || j >= t1) {
            const int slot = atomicAdd(&n_ext, 1);
            if (slot < kTileEdgeCap) ext[slot] = j;
        }
    }
    __syncthreads();
    const int ne = n_ext;
    bool ok = ne <= kTileEdgeCap;
    if (ok) {
        int pow2 = 1;
        while (pow2 < ne) pow2 <<= 1;
        for (int i = ne + threadIdx.x; i < pow2; i += blockDim.x) ext[i] = 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= pow2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
                    const int l = i ^ j;
                    if (l > i) {
                        const int a = ext[i], b = ext[l];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { ext[i] = b; ext[l] = a; }
                    }
                }
                __syncthreads();
            }
        // unique (ascending): heads are entries that differ from their predecessor
        for (int i = threadIdx.x; i < ne; i += blockDim.x)
            if (i == 0 || ext[i] != ext[i - 1]) {      // compacted in arrival order, sorted again below (the list is short)
                const int slot = atomicAdd(&n_uniq, 1);
                if (slot < kTileHaloCap) uniq[slot] = ext[i];
            }
        __syncthreads();
        ok = n_uniq <= kTileHaloCap;
    }
    if (!ok) {
        if (threadIdx.x == 0) { tile_halo_count[tile] = -1; *bad = 1; }
        return;
    }
    const int nu = n_uniq;
    {   // the compacted heads arrived in atomic order: sort them (<= kTileHaloCap entries)
        int pow2 = 1;
        while (pow2 < nu) pow2 <<= 1;
        // sorted in ext[] (free again)
        for (int i = threadIdx.x; i < pow2; i += blockDim.x) ext[i] = i < nu ? uniq[i] : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= pow2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
                    const int l = i ^ j;
                    if (l > i) {
                        const int a = ext[i], b = ext[l];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { ext[i] = b; ext[l] = a; }
                    }
                }
                __syncthreads();
            }
    }
    for (int i = threadIdx.x; i < nu; i += blockDim.x) tile_halo[(size_t)tile * kTileHaloCap + i] = ext[i];
    if (threadIdx.x == 0) { tile_halo_count[tile] = nu; atomicMax(max_halo, nu); }
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        const int j = colidx[k];
        int local;
        if (j >= t0 && j < t1) local = j - t0;
        else {
            int lo = 0, hi = nu - 1;                         // binary search in the sorted halo
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ext[mid] < j) lo = mid + 1; else hi = mid;
            }
            local = kTile + lo;
        }
        tile_colidx[k] = (unsigned short)local;
    }
}

struct TileView {
    const int *halo;                     // [n_tiles][kTileHaloCap]
    const int *halo_count;               // [n_tiles]
    const unsigned short *colidx;        // [nnz] tile-local column indices
    int n_tiles;
};

// Stage one array of per-vertex records for tile `tile`: own rows coalesced, halo rows through the list. No barrier inside.
template <typename V>
__device__ __forceinline__ void stage_records(V *__restrict__ s, const V *__restrict__ g, int t0, int rows, const int *__restrict__ halo, int n_halo) {
    if ((int)threadIdx.x < rows) s[threadIdx.x] = g[t0 + threadIdx.x];
    for (int h = threadIdx.x; h < n_halo; h += blockDim.x) s[kTile + h] = g[__ldg(&halo[h])];
}

// 32-byte records (Vec4T<double>) are staged as two 16-byte halves so that the shared-memory gathers are LDS.128
struct alignas(16) Half32 { double a, b; };
template <typename S> struct StageRec;
template <> struct StageRec<float> {
    typedef Vec4T<float> type;
    static __device__ __forceinline__ Vec4T<float> get(const Vec4T<float> *s, int i) {
        const float4 v = *reinterpret_cast<const float4 *>(s + i);
        return Vec4T<float>{v.x, v.y, v.z, v.w};
    }
};
template <> struct StageRec<double> {
    typedef Vec4T<double> type;
    static __device__ __forceinline__ Vec4T<double> get(const Vec4T<double> *s, int i) {
        const Half32 *p = reinterpret_cast<const Half32 *>(s + i);
        const Half32 lo = p[0], hi = p[1];
        return Vec4T<double>{lo.a, lo.b, hi.a, hi.b};
    }
};

// ---- local step (arap.h:354-384), tiled -----------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(kBlock, ARAP_LOCAL_MIN_BLOCKS) local_step_tiled_kernel(int n, TileView tv, const int *__restrict__ rowptr,
                                                                  const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                                  const Vec4T<S> *__restrict__ cur4, Vec4T<S> *__restrict__ quat,
                                                                  int *__restrict__ redo_list, int *__restrict__ redo_count) {
    extern __shared__ __align__(32) unsigned char tile_smem[];
    pdl_enter();
    for (int tile = blockIdx.x; tile < tv.n_tiles; tile += gridDim.x) {
        const int t0 = tile * kTile, rows = min(kTile, n - t0);
        const int nh = tv.halo_count[tile];
        const int *halo = tv.halo + (size_t)tile * kTileHaloCap;
        Vec4T<S> *s_rest = reinterpret_cast<Vec4T<S> *>(tile_smem);
        Vec4T<S> *s_cur = s_rest + (kTile + nh);
        const int i = t0 + (int)threadIdx.x;
        const bool active = (int)threadIdx.x < rows;
        Vec4T<S> pi = Vec4T<S>{0, 0, 0, 0}, ci = pi;
        if (active) { pi = load4<S>(&rest4[i]); ci = load4<S>(&cur4[i]); s_rest[threadIdx.x] = pi; s_cur[threadIdx.x] = ci; }
        for (int h = threadIdx.x; h < nh; h += blockDim.x) {
            const int j = __ldg(&halo[h]);
            s_rest[kTile + h] = load4<S>(&rest4[j]);
            s_cur[kTile + h] = load4<S>(&cur4[j]);
        }
        int k0 = 0, k1 = 0;
        Vec4T<S> qprev = Vec4T<S>{1, 0, 0, 0};
        if (active) { k0 = rowptr[i]; k1 = rowptr[i + 1]; qprev = load4<S>(&quat[i]); }
        __syncthreads();
        if (active) {
            S cov[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) cov[c] = S(0);
            constexpr int CH = kTileChunk;               // index / weight loads of CH entries in flight before the first use
            for (int k = k0; k < k1; k += CH) {
                int lj[CH];
                S w[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const bool valid = k + u < k1;
                    lj[u] = valid ? (int)tv.colidx[k + u] : (int)threadIdx.x;
                    w[u] = valid ? __ldg(&weight[k + u]) : S(0);
                }
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const Vec4T<S> pj = StageRec<S>::get(s_rest, lj[u]), cj = StageRec<S>::get(s_cur, lj[u]);
                    const S ex = w[u] * (pi.x - pj.x), ey = w[u] * (pi.y - pj.y), ez = w[u] * (pi.z - pj.z);
                    const S dx = ci.x - cj.x, dy = ci.y - cj.y, dz = ci.z - cj.z;
                    cov[0] += ex * dx; cov[1] += ex * dy; cov[2] += ex * dz;
                    cov[3] += ey * dx; cov[4] += ey * dy; cov[5] += ey * dz;
                    cov[6] += ez * dx; cov[7] += ez * dy; cov[8] += ez * dz;
                }
            }
            const S qp[4] = {qprev.x, qprev.y, qprev.z, qprev.w};
            S q[4];
            if (rotation_from_covariance_newton_only<S>(cov, qp, q)) store4<S>(&quat[i], q[0], q[1], q[2], q[3]);
            else redo_list[atomicAdd(redo_count, 1)] = i;
        }
        __syncthreads();                     // the next tile overwrites the staged records
    }
}

// ---- right-hand side + residual + CG start (arap.h:393-414), tiled ----------------------------------------------------
template <typename S, bool MG>
__global__ void __launch_bounds__(kBlock, 3) rhs_residual_tiled_kernel(int n, TileView tv, const int *__restrict__ rowptr,
                                                                    const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                                    const Vec4T<S> *__restrict__ cur4, const Vec4T<S> *__restrict__ quat,
                                                                    const double *__restrict__ inv_diag, double omega0,
                                                                    Vec3d *__restrict__ r_out, Vec3d *__restrict__ d_out,
                                                                    Vec3d *__restrict__ x_out, float4 *__restrict__ x0_out,
                                                                    double *__restrict__ partials, unsigned *__restrict__ counter,
                                                                    CgScalars *__restrict__ cg) {
    extern __shared__ __align__(32) unsigned char tile_smem[];
    pdl_enter();
    double red[5] = {0, 0, 0, 0, 0};   // rho x,y,z ; rr ; ref2
    for (int tile = blockIdx.x; tile < tv.n_tiles; tile += gridDim.x) {
        const int t0 = tile * kTile, rows = min(kTile, n - t0);
        const int nh = tv.halo_count[tile];
        const int *halo = tv.halo + (size_t)tile * kTileHaloCap;
        Vec4T<S> *s_rest = reinterpret_cast<Vec4T<S> *>(tile_smem);
        Vec4T<S> *s_cur = s_rest + (kTile + nh);
        Vec4T<S> *s_quat = s_cur + (kTile + nh);
        const int i = t0 + (int)threadIdx.x;
        const bool active = (int)threadIdx.x < rows;
        Vec4T<S> pi = Vec4T<S>{0, 0, 0, 0}, ci = pi, qi = pi;
        if (active) {
            pi = load4<S>(&rest4[i]); ci = load4<S>(&cur4[i]); qi = load4<S>(&quat[i]);
            s_rest[threadIdx.x] = pi; s_cur[threadIdx.x] = ci; s_quat[threadIdx.x] = qi;
        }
        for (int h = threadIdx.x; h < nh; h += blockDim.x) {
            const int j = __ldg(&halo[h]);
            s_rest[kTile + h] = load4<S>(&rest4[j]);
            s_cur[kTile + h] = load4<S>(&cur4[j]);
            s_quat[kTile + h] = load4<S>(&quat[j]);
        }
        int k0 = 0, k1 = 0;
        double idg = 0.0;
        if (active && pi.w != S(0)) { k0 = rowptr[i]; k1 = rowptr[i + 1]; idg = inv_diag[i]; }
        __syncthreads();
        if (active) {
            Vec3d r = {0, 0, 0}, z = {0, 0, 0};
            if (pi.w != S(0)) {
                double rot_j[3] = {0, 0, 0};     // sum_j (w/2) R_j e_ij
                double se[3] = {0, 0, 0};        // sum_j (w/2) e_ij
                double lap[3] = {0, 0, 0};       // sum_j w (p'_i - p'_j)
                constexpr int CH = kTileChunk;
                for (int k = k0; k < k1; k += CH) {
                    int lj[CH];
                    S w[CH];
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const bool valid = k + u < k1;
                        lj[u] = valid ? (int)tv.colidx[k + u] : (int)threadIdx.x;
                        w[u] = valid ? __ldg(&weight[k + u]) : S(0);
                    }
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const Vec4T<S> pj = StageRec<S>::get(s_rest, lj[u]), cj = StageRec<S>::get(s_cur, lj[u]), qj = StageRec<S>::get(s_quat, lj[u]);
                        const S hw = S(0.5) * w[u];
                        const S ex = hw * (pi.x - pj.x), ey = hw * (pi.y - pj.y), ez = hw * (pi.z - pj.z);
                        S rx, ry, rz;                                      // R_j e_ij, quat stored as (w,x,y,z) in (.x,.y,.z,.w)
                        quat_rotate<S>(qj.x, qj.y, qj.z, qj.w, ex, ey, ez, rx, ry, rz);
                        rot_j[0] += (double)rx; rot_j[1] += (double)ry; rot_j[2] += (double)rz;
                        se[0] += (double)ex; se[1] += (double)ey; se[2] += (double)ez;
                        lap[0] += (double)w[u] * ((double)ci.x - (double)cj.x);
                        lap[1] += (double)w[u] * ((double)ci.y - (double)cj.y);
                        lap[2] += (double)w[u] * ((double)ci.z - (double)cj.z);
                    }
                }
                double ox, oy, oz;                                                 // R_i sum_j (w/2) e_ij
                quat_rotate<double>((double)qi.x, (double)qi.y, (double)qi.z, (double)qi.w, se[0], se[1], se[2], ox, oy, oz);
                const double rhs0 = rot_j[0] + ox, rhs1 = rot_j[1] + oy, rhs2 = rot_j[2] + oz;
                r.x = rhs0 - lap[0]; r.y = rhs1 - lap[1]; r.z = rhs2 - lap[2];
                z.x = r.x * idg; z.y = r.y * idg; z.z = r.z * idg;
                red[0] += r.x * z.x; red[1] += r.y * z.y; red[2] += r.z * z.z;
                red[3] += r.x * r.x + r.y * r.y + r.z * r.z;
                red[4] += rhs0 * rhs0 + rhs1 * rhs1 + rhs2 * rhs2;
            }
            r_out[i] = r;
            if (MG) {
                x0_out[i] = make_float4((float)(omega0 * z.x), (float)(omega0 * z.y), (float)(omega0 * z.z), 0.f);
            } else {
                x_out[i] = Vec3d{0, 0, 0};
                d_out[i] = z;
            }
        }
        __syncthreads();
    }
    double total[5];
    if (grid_sum_last_block<5>(red, partials, counter, total))
        cg_finish_reduction<5>(cg, MG ? CG_STAGE_START_MG : CG_STAGE_START_JACOBI, total);
}

// ---- fine level of the V-cycle and the CG's matrix-vector product, tiled (fp32 records, 16 bytes) -------------------------
// (A x)_i = sum_j w_ij (x_i - x_j) over the row, x gathered from the staged tile
__device__ __forceinline__ float3 fine_apply_row_tiled(int k0, int k1, const unsigned short *__restrict__ lcol, const float *__restrict__ weight,
                                                       const MgVec *__restrict__ s_x, const MgVec xi) {
    float3 out = {0.f, 0.f, 0.f};
    constexpr int CH = 3;
    for (int k = k0; k < k1; k += CH) {
        int lj[CH];
        float w[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const bool valid = k + u < k1;
            lj[u] = valid ? (int)lcol[k + u] : (int)threadIdx.x;
            w[u] = valid ? __ldg(&weight[k + u]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const MgVec xj = s_x[lj[u]];
            out.x += w[u] * (xi.x - xj.x); out.y += w[u] * (xi.y - xj.y); out.z += w[u] * (xi.z - xj.z);
        }
    }
    return out;
}

__global__ void __launch_bounds__(kBlock) mg_fine_residual_tiled_kernel(int n, TileView tv, const int *__restrict__ rowptr,
                                                                        const float *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                                        const Vec3d *__restrict__ b, const MgVec *__restrict__ x,
                                                                        MgVec *__restrict__ r, const CgScalars *__restrict__ cg) {
    extern __shared__ __align__(32) unsigned char tile_smem[];
    MgVec *s_x = reinterpret_cast<MgVec *>(tile_smem);
    pdl_enter();
    if (cg->converged) return;
    for (int tile = blockIdx.x; tile < tv.n_tiles; tile += gridDim.x) {
        const int t0 = tile * kTile, rows = min(kTile, n - t0);
        const int nh = tv.halo_count[tile];
        stage_records<MgVec>(s_x, x, t0, rows, tv.halo + (size_t)tile * kTileHaloCap, nh);
        const int i = t0 + (int)threadIdx.x;
        const bool active = (int)threadIdx.x < rows;
        const bool is_free = active && free_mask[i];
        int k0 = 0, k1 = 0;
        Vec3d bi = {0, 0, 0};
        if (is_free) { k0 = rowptr[i]; k1 = rowptr[i + 1]; bi = b[i]; }
        __syncthreads();
        if (active) {
            MgVec out = {0.f, 0.f, 0.f, 0.f};
            if (is_free) {
                const float3 ax = fine_apply_row_tiled(k0, k1, tv.colidx, weight, s_x, s_x[threadIdx.x]);
                out.x = (float)bi.x - ax.x; out.y = (float)bi.y - ax.y; out.z = (float)bi.z - ax.z;
            }
            r[i] = out;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock) mg_fine_postsmooth_tiled_kernel(int n, TileView tv, const int *__restrict__ rowptr,
                                                                          const float *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                                          const double *__restrict__ inv_diag, double omega,
                                                                          const Vec3d *__restrict__ b, const MgVec *__restrict__ x,
                                                                          MgVec *__restrict__ z, double *__restrict__ partials,
                                                                          unsigned *__restrict__ counter, CgScalars *__restrict__ cg) {
    extern __shared__ __align__(32) unsigned char tile_smem[];
    MgVec *s_x = reinterpret_cast<MgVec *>(tile_smem);
    pdl_enter();
    if (cg->converged) return;
    double red[4] = {0, 0, 0, 0};
    const double inv_len2 = cg->inv_len2;
    for (int tile = blockIdx.x; tile < tv.n_tiles; tile += gridDim.x) {
        const int t0 = tile * kTile, rows = min(kTile, n - t0);
        const int nh = tv.halo_count[tile];
        stage_records<MgVec>(s_x, x, t0, rows, tv.halo + (size_t)tile * kTileHaloCap, nh);
        const int i = t0 + (int)threadIdx.x;
        const bool active = (int)threadIdx.x < rows;
        const bool is_free = active && free_mask[i];
        int k0 = 0, k1 = 0;
        Vec3d bi = {0, 0, 0};
        float s = 0.f;
        if (is_free) { k0 = rowptr[i]; k1 = rowptr[i + 1]; bi = b[i]; s = (float)(omega * inv_diag[i]); }
        __syncthreads();
        if (active) {
            MgVec out = {0.f, 0.f, 0.f, 0.f};
            if (is_free) {
                const MgVec xi = s_x[threadIdx.x];
                const float3 ax = fine_apply_row_tiled(k0, k1, tv.colidx, weight, s_x, xi);
                out.x = xi.x + s * ((float)bi.x - ax.x); out.y = xi.y + s * ((float)bi.y - ax.y); out.z = xi.z + s * ((float)bi.z - ax.z);
                red[0] += bi.x * (double)out.x; red[1] += bi.y * (double)out.y; red[2] += bi.z * (double)out.z;
                red[3] += z_norm8(out, inv_len2);
            }
            z[i] = out;
        }
        __syncthreads();
    }
    double total[4];
    if (grid_sum_last_block<4>(red, partials, counter, total)) cg_finish_reduction<4>(cg, CG_STAGE_GAMMA, total, 0);
}

template <typename S>
__global__ void __launch_bounds__(kBlock, ARAP_SPMV_MIN_BLOCKS) cg_spmv_z_tiled_kernel(int n, TileView tv, const int *__restrict__ rowptr,
                                                                 const S *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                                 const float4 *__restrict__ z, Vec3d *__restrict__ w_out,
                                                                 double *__restrict__ partials, unsigned *__restrict__ counter,
                                                                 CgScalars *__restrict__ cg) {
    extern __shared__ __align__(32) unsigned char tile_smem[];
    float4 *s_z = reinterpret_cast<float4 *>(tile_smem);
    pdl_enter();
    if (cg->converged) return;
    double red[3] = {0, 0, 0};
    for (int tile = blockIdx.x; tile < tv.n_tiles; tile += gridDim.x) {
        const int t0 = tile * kTile, rows = min(kTile, n - t0);
        const int nh = tv.halo_count[tile];
        stage_records<float4>(s_z, z, t0, rows, tv.halo + (size_t)tile * kTileHaloCap, nh);
        const int i = t0 + (int)threadIdx.x;
        const bool active = (int)threadIdx.x < rows;
        const bool is_free = active && free_mask[i];
        int k0 = 0, k1 = 0;
        if (is_free) { k0 = rowptr[i]; k1 = rowptr[i + 1]; }
        __syncthreads();
        if (active) {
            Vec3d out = {0, 0, 0};
            if (is_free) {
                const float4 zi = s_z[threadIdx.x];
                constexpr int CH = 3;
                for (int k = k0; k < k1; k += CH) {
                    int lj[CH];
                    double w[CH];
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const bool valid = k + u < k1;
                        lj[u] = valid ? (int)tv.colidx[k + u] : (int)threadIdx.x;
                        w[u] = valid ? (double)__ldg(&weight[k + u]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        const float4 zj = s_z[lj[u]];
                        out.x += w[u] * ((double)zi.x - (double)zj.x);
                        out.y += w[u] * ((double)zi.y - (double)zj.y);
                        out.z += w[u] * ((double)zi.z - (double)zj.z);
                    }
                }
                red[0] += (double)zi.x * out.x; red[1] += (double)zi.y * out.y; red[2] += (double)zi.z * out.z;
            }
            w_out[i] = out;
        }
        __syncthreads();
    }
    double total[3];
    if (grid_sum_last_block<3>(red, partials, counter, total)) cg_finish_reduction<3>(cg, CG_STAGE_DELTA, total, 4);
}

}  // namespace arap
