// kernels.cuh -- the CUDA kernels of the ARAP hot path (sm_100a). One thread per face / vertex / row;
// all are HBM-bound gather/stream kernels (no tensor cores: there is no dense contraction here).
//
// Data layout in HBM (S = float|double = the reference's PrecisionType):
//   rest4[V], cur4[V]  Vec4T<S>: (x,y,z,mask) rest pose p, (x,y,z,-) current pose p'.  rest4.w = 1 free / 0 constrained.
//   quat[V]            Vec4T<S>: unit quaternion (w,x,y,z) of R_i (the reference stores 3x3 matrices, arap.h:454).
//   rowptr[V+1], colidx[nnz] int32, weight[nnz] S: _edgeWeights CSR (arap.h:453), columns ascending.
//   cg vectors         Vec3d[V]: one (x,y,z) triple of doubles per vertex (three right-hand sides solved together).
#pragma once

#include "arap_math.cuh"
#include "device_utils.cuh"
#include "tma_stage.cuh"

namespace arap {

template <typename S> __device__ __forceinline__ Vec4T<S> load4(const Vec4T<S> *p);
template <> __device__ __forceinline__ Vec4T<float> load4<float>(const Vec4T<float> *p) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    return Vec4T<float>{v.x, v.y, v.z, v.w};
}
// sm_100 has 256-bit global loads/stores (SASS LDG.E.256 / STG.E.256): one request per 32-byte element, i.e. per
// 32-byte sector, instead of two 128-bit requests that each touch the same sector -- halves the L1 request traffic of
// the double-precision gathers, which is what bounds the local step and the RHS kernel (profiles/README.md).
template <> __device__ __forceinline__ Vec4T<double> load4<double>(const Vec4T<double> *p) {
    Vec4T<double> r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
template <typename S> __device__ __forceinline__ void store4(Vec4T<S> *p, S x, S y, S z, S w);
template <> __device__ __forceinline__ void store4<float>(Vec4T<float> *p, float x, float y, float z, float w) {
    *reinterpret_cast<float4 *>(p) = make_float4(x, y, z, w);
}
template <> __device__ __forceinline__ void store4<double>(Vec4T<double> *p, double x, double y, double z, double w) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
}

// ---- load-batching gates ---------------------------------------------------------------------------
// ptxas interleaves each gather with the arithmetic that consumes it (to shorten live ranges), which
// leaves only one neighbour's loads in flight per thread.
// `0 * x` cannot be folded under IEEE rules (x could be NaN/Inf), so adding gate = 0 * (sum of one word
// of every gathered value) to the chunk's weights makes every product depend on ALL gathers of the
// chunk: the loads are then issued back to back and a whole chunk is in flight. Numerically a no-op
// (w + 0.0 == w) for finite data. Measured on B200 (profiles/r01_d_variants.txt): batching 2 neighbours
// is the sweet spot for the SpMV-like kernels (40 registers, full occupancy); batching 6 costs 78
// registers and is 15 % SLOWER, and the local step / RHS kernels are best without a gate at all.
template <int N>
__device__ __forceinline__ double gather_gate(const Vec3d (&a)[N]) {
    double s = a[0].z;
#pragma unroll
    for (int u = 1; u < N; ++u) s += a[u].z;
    return 0.0 * s;
}
template <int N>
__device__ __forceinline__ double gather_gate(const Vec4T<double> (&a)[N]) {   // one 32-byte load per element
    double s = a[0].x;
#pragma unroll
    for (int u = 1; u < N; ++u) s += a[u].x;
    return 0.0 * s;
}
template <int N>
__device__ __forceinline__ float gather_gate(const Vec4T<float> (&a)[N]) {     // one 16-byte load per element
    float s = a[0].x;
#pragma unroll
    for (int u = 1; u < N; ++u) s += a[u].x;
    return 0.0f * s;
}

// =================================================================================================
// Setup path: cotan weights + CSR (reference arap.h:182-239), free map (:261-272), constraints (:277-281)
// =================================================================================================

// K1a: how many raw triplets land in each row. Each face emits, per edge, (lo,hi) and (hi,lo)  (arap.h:225-232).
__global__ void __launch_bounds__(kBlock) weights_count_kernel(const int *__restrict__ faces, int n_faces, int *__restrict__ row_count) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const int v0 = faces[3 * (size_t)f], v1 = faces[3 * (size_t)f + 1], v2 = faces[3 * (size_t)f + 2];
    atomicAdd(&row_count[v0], 2);   // v0 is an end point of e0 and e2
    atomicAdd(&row_count[v1], 2);   // e0, e1
    atomicAdd(&row_count[v2], 2);   // e1, e2
}

// K1b: compute the three half-cotans of every face and scatter the 6 triplets into their rows.
// tag = insertion index of the triplet in the reference's triplet list (6 f + slot), used to sum
// duplicates in the reference's order.
template <typename S>
__global__ void __launch_bounds__(kBlock) weights_fill_kernel(const int *__restrict__ faces, int n_faces, const S *__restrict__ rest_xyz,
                                                              const int *__restrict__ raw_rowptr, int *__restrict__ row_cursor,
                                                              int *__restrict__ raw_col, S *__restrict__ raw_val, unsigned *__restrict__ raw_tag) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    int vid[3];
    S pos[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vid[k] = faces[3 * (size_t)f + k];
#pragma unroll
        for (int d = 0; d < 3; ++d) pos[k][d] = rest_xyz[3 * (size_t)vid[k] + d];
    }
    S hw[3];
    cotan_half_weights<S>(pos[0], pos[1], pos[2], hw);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = vid[k], b = vid[(k + 1) % 3];
        const int lo = a > b ? b : a, hi = a > b ? a : b;           // undirectedEdge (arap.h:435-440)
        int slot = raw_rowptr[lo] + atomicAdd(&row_cursor[lo], 1);
        raw_col[slot] = hi; raw_val[slot] = hw[k]; raw_tag[slot] = 6u * (unsigned)f + (unsigned)k;
        slot = raw_rowptr[hi] + atomicAdd(&row_cursor[hi], 1);
        raw_col[slot] = lo; raw_val[slot] = hw[k]; raw_tag[slot] = 6u * (unsigned)f + 3u + (unsigned)k;
    }
}

constexpr int kLongRow = 64;     // raw triplets per row above which one thread's insertion sort (O(d^2)) is handed to a CTA

// K2a: setFromTriplets for one row (arap.h:238): sort the row's raw triplets by (column, insertion order),
// sum duplicates in insertion order, leave the unique entries at the front of the raw segment.
template <typename S>
__global__ void __launch_bounds__(kBlock) row_sort_merge_kernel(int n_rows, const int *__restrict__ raw_rowptr, int *__restrict__ raw_col,
                                                                S *__restrict__ raw_val, unsigned *__restrict__ raw_tag,
                                                                int *__restrict__ unique_count, int *__restrict__ long_rows,
                                                                int *__restrict__ long_count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int lo = raw_rowptr[r], hi = raw_rowptr[r + 1];
    if (hi - lo > kLongRow) {                                        // fan centre / pole: sorted by a whole CTA (row_sort_long_kernel)
        long_rows[atomicAdd(long_count, 1)] = r;
        return;
    }
    for (int a = lo + 1; a < hi; ++a) {                              // insertion sort: rows are ~12 entries long
        const int cj = raw_col[a]; const S cv = raw_val[a]; const unsigned ct = raw_tag[a];
        int b = a - 1;
        while (b >= lo && (raw_col[b] > cj || (raw_col[b] == cj && raw_tag[b] > ct))) {
            raw_col[b + 1] = raw_col[b]; raw_val[b + 1] = raw_val[b]; raw_tag[b + 1] = raw_tag[b];
            --b;
        }
        raw_col[b + 1] = cj; raw_val[b + 1] = cv; raw_tag[b + 1] = ct;
    }
    int out = lo;
    for (int a = lo; a < hi;) {
        const int cj = raw_col[a];
        S s = raw_val[a];
        ++a;
        while (a < hi && raw_col[a] == cj) { s = add_rn(s, raw_val[a]); ++a; }
        raw_col[out] = cj; raw_val[out] = s; ++out;
    }
    unique_count[r] = out - lo;
}

// K2a for the rows row_sort_merge_kernel skipped (vertices of very high valence: the centre of a triangle fan, the pole
// of a UV sphere): one CTA per row sorts the row's triplets by (column, insertion order) with a bitonic network whose
// comparators all point the same way ("flip" variant), so a row of any length is sorted as if it were padded with +inf
// to the next power of two; then one thread sums the duplicates in insertion order, as above. O(d log^2 d / 256).
template <typename S>
__global__ void __launch_bounds__(kBlock) row_sort_long_kernel(const int *__restrict__ long_rows, const int *__restrict__ long_count,
                                                               const int *__restrict__ raw_rowptr, int *__restrict__ raw_col,
                                                               S *__restrict__ raw_val, unsigned *__restrict__ raw_tag,
                                                               int *__restrict__ unique_count) {
    for (int t = blockIdx.x; t < *long_count; t += gridDim.x) {
        const int r = long_rows[t];
        const int lo = raw_rowptr[r], n = raw_rowptr[r + 1] - lo;
        int *col = raw_col + lo;
        S *val = raw_val + lo;
        unsigned *tag = raw_tag + lo;
        int pow2 = 1;
        while (pow2 < n) pow2 <<= 1;
        for (int k = 2; k <= pow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const bool flip = (j == (k >> 1));
                for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
                    const int l = flip ? (i ^ (k - 1)) : (i ^ j);
                    if (l > i && l < n) {
                        const int ci = col[i], cl = col[l];
                        const unsigned ti = tag[i], tl = tag[l];
                        if (ci > cl || (ci == cl && ti > tl)) {
                            col[i] = cl; col[l] = ci; tag[i] = tl; tag[l] = ti;
                            const S v = val[i]; val[i] = val[l]; val[l] = v;
                        }
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            int out = 0;
            for (int a = 0; a < n;) {
                const int cj = col[a];
                S sum = val[a];
                ++a;
                while (a < n && col[a] == cj) { sum = add_rn(sum, val[a]); ++a; }
                col[out] = cj; val[out] = sum; ++out;
            }
            unique_count[r] = out;
        }
        __syncthreads();
    }
}

// K2b: pack the unique entries into the final CSR.
template <typename S>
__global__ void __launch_bounds__(kBlock) csr_compact_kernel(int n_rows, const int *__restrict__ raw_rowptr, const int *__restrict__ raw_col,
                                                             const S *__restrict__ raw_val, const int *__restrict__ rowptr,
                                                             int *__restrict__ colidx, S *__restrict__ weight) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int src = raw_rowptr[r], dst = rowptr[r], n = rowptr[r + 1] - dst;
    for (int k = 0; k < n; ++k) { colidx[dst + k] = raw_col[src + k]; weight[dst + k] = raw_val[src + k]; }
}

// setConstraint (arap.h:81-85): scatter n (index, location) pairs into the dense constraint table.
template <typename S, typename T>
__global__ void __launch_bounds__(kBlock) set_constraints_kernel(int n, const int *__restrict__ idx, const T *__restrict__ xyz,
                                                                 int n_vertices, unsigned char *__restrict__ is_constrained,
                                                                 S *__restrict__ target_xyz) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int v = idx[k];
    if (v < 0 || v >= n_vertices) return;
    is_constrained[v] = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) target_xyz[3 * (size_t)v + d] = (S)xyz[3 * (size_t)k + d];
}

// DeformationUtil::updateConstraints (deformation_util.h:48-57) for a whole batch: member m's handle k goes to
// transform_m * rest_k. transforms: 12 doubles per member (the top three rows of the 4x4, row-major).
template <typename S, typename T>
__global__ void __launch_bounds__(kBlock) set_rigid_constraints_kernel(int n, int batch, int member_stride, const int *__restrict__ idx,
                                                                       const T *__restrict__ rest, const double *__restrict__ transforms,
                                                                       int n_vertices, unsigned char *__restrict__ is_constrained,
                                                                       S *__restrict__ target_xyz) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * batch) return;
    const int m = (int)(t / n), k = (int)(t - (size_t)m * n);
    const long long v = (long long)idx[k] + (long long)m * member_stride;
    if (idx[k] < 0 || v >= n_vertices) return;
    const double *M = transforms + 12 * (size_t)m;
    const double x = (double)rest[3 * (size_t)k], y = (double)rest[3 * (size_t)k + 1], z = (double)rest[3 * (size_t)k + 2];
    is_constrained[v] = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) target_xyz[3 * (size_t)v + d] = (S)(M[4 * d] * x + M[4 * d + 1] * y + M[4 * d + 2] * z + M[4 * d + 3]);
}

template <typename S, typename T>
__global__ void __launch_bounds__(kBlock) cast_xyz_kernel(size_t n, const T *__restrict__ in, S *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (S)in[i];
}

__global__ void __launch_bounds__(kBlock) free_flag_kernel(int n, const unsigned char *__restrict__ is_constrained, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = is_constrained[i] ? 0 : 1;
}

// initializeFreeVariableMapping (arap.h:261-272): prefix -> free index, -1 for constrained.
__global__ void __launch_bounds__(kBlock) free_map_kernel(int n, const unsigned char *__restrict__ is_constrained,
                                                          const int *__restrict__ prefix, int *__restrict__ free_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) free_idx[i] = is_constrained[i] ? -1 : prefix[i];
}

// initializeMeshGeometry + initializeRotations + initializeConstraints (arap.h:162-168, 246-249, 277-281)
// plus the Jacobi preconditioner 1 / L_ii = 1 / sum_j w_ij (the diagonal of arap.h:332).
// Internal vertex order. The engine renumbers vertices once (Morton order of the rest pose, see engine.cu) so that a
// CTA's rows and their one-ring neighbours sit close together in memory; perm[internal] = user index. Everything at the
// C-ABI boundary (constraints, CSR export, positions, rotations) stays in the user's numbering.
__global__ void __launch_bounds__(kBlock) invert_perm_kernel(int n, const int *__restrict__ perm, int *__restrict__ iperm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) iperm[perm[i]] = i;
}
__global__ void __launch_bounds__(kBlock) perm_row_count_kernel(int n, const int *__restrict__ perm, const int *__restrict__ rowptr_user,
                                                                int *__restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int u = perm[i]; count[i] = rowptr_user[u + 1] - rowptr_user[u]; }
}
template <typename S>
__global__ void __launch_bounds__(kBlock) perm_csr_fill_kernel(int n, const int *__restrict__ perm, const int *__restrict__ iperm,
                                                               const int *__restrict__ rowptr_user, const int *__restrict__ colidx_user,
                                                               const S *__restrict__ weight_user, const int *__restrict__ rowptr_hot,
                                                               int *__restrict__ colidx_hot, S *__restrict__ weight_hot,
                                                               float *__restrict__ weight_hot_f32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int u = perm[i];
    const int src = rowptr_user[u], len = rowptr_user[u + 1] - src, dst = rowptr_hot[i];
    for (int k = 0; k < len; ++k) {                 // the user row's column order is kept, so every sum runs in the reference's order
#if defined(ARAP_DIAG_SELFGATHER)
        colidx_hot[dst + k] = (ARAP_DIAG_SELFGATHER == 1) ? i : min(n - 1, i + 1 + k);   /* diagnostic: gathers without inter-CTA reuse */
#else
        colidx_hot[dst + k] = iperm[colidx_user[src + k]];
#endif
        weight_hot[dst + k] = weight_user[src + k];
        weight_hot_f32[dst + k] = (float)weight_user[src + k];      // the fp32 multigrid preconditioner's copy
    }
}

// initializeMeshGeometry + initializeRotations + initializeConstraints (arap.h:162-168, 246-249, 277-281) into the
// solver layout (internal order), plus the Jacobi preconditioner 1 / L_ii = 1 / sum_j w_ij (the diagonal of arap.h:332).
template <typename S>
__global__ void __launch_bounds__(kBlock) init_state_kernel(int n, const int *__restrict__ perm, const S *__restrict__ rest_xyz,
                                                            const unsigned char *__restrict__ is_constrained,
                                                            const S *__restrict__ target_xyz, const int *__restrict__ rowptr,
                                                            const S *__restrict__ weight, Vec4T<S> *__restrict__ rest4,
                                                            Vec4T<S> *__restrict__ cur4, Vec4T<S> *__restrict__ quat,
                                                            double *__restrict__ inv_diag, unsigned char *__restrict__ free_mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t u = (size_t)perm[i];
    const bool con = is_constrained[u] != 0;
    free_mask[i] = con ? 0 : 1;      // one byte per row for the solver kernels (reading rest4[i].w would pull a 32-byte sector)
    const S x = rest_xyz[3 * u], y = rest_xyz[3 * u + 1], z = rest_xyz[3 * u + 2];
    store4<S>(&rest4[i], x, y, z, con ? S(0) : S(1));
    if (con) store4<S>(&cur4[i], target_xyz[3 * u], target_xyz[3 * u + 1], target_xyz[3 * u + 2], S(0));
    else store4<S>(&cur4[i], x, y, z, S(0));
    store4<S>(&quat[i], S(1), S(0), S(0), S(0));
    double diag = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) diag += (double)weight[k];
    inv_diag[i] = diag > 0.0 ? 1.0 / diag : 0.0;
}

// =================================================================================================
// Local step (reference arap.h:354-384): S_i = sum_j w_ij (p_i-p_j)(p'_i-p'_j)^T, R_i from its SVD.
// =================================================================================================
// Neighbours are gathered in chunks of kGather<S> so that all index loads, then all position gathers
// of a chunk are in flight together (memory-level parallelism) before the arithmetic starts.
#ifndef ARAP_LOCAL_CHUNK_F64
#define ARAP_LOCAL_CHUNK_F64 2
#endif
#ifndef ARAP_RHS_CHUNK_F64
#define ARAP_RHS_CHUNK_F64 3
#endif
#ifndef ARAP_SPMV_CHUNK
#define ARAP_SPMV_CHUNK 3
#endif
#ifndef ARAP_SPMV_MIN_BLOCKS
#define ARAP_SPMV_MIN_BLOCKS 4
#endif
#ifndef ARAP_LOCAL_GATE
#define ARAP_LOCAL_GATE 0
#endif
#ifndef ARAP_RHS_GATE
#define ARAP_RHS_GATE 0
#endif
#ifndef ARAP_LOCAL_MIN_BLOCKS
#define ARAP_LOCAL_MIN_BLOCKS 4
#endif
#ifndef ARAP_RHS_MIN_BLOCKS
#define ARAP_RHS_MIN_BLOCKS 2      /* 128 registers, no spills, 2 CTAs per SM: 60.9 us; 3 CTAs (80 registers, 164 B of spills): 65.9 us */
#endif
template <typename S> struct GatherChunk;
template <> struct GatherChunk<float> { static constexpr int value = 6; };
template <> struct GatherChunk<double> { static constexpr int value = ARAP_LOCAL_CHUNK_F64; };
template <typename S> struct RhsChunk;
template <> struct RhsChunk<float> { static constexpr int value = 4; };
template <> struct RhsChunk<double> { static constexpr int value = ARAP_RHS_CHUNK_F64; };
constexpr int kSpmvChunk = ARAP_SPMV_CHUNK;

// S_i of one vertex: neighbours gathered in chunks of GatherChunk<S> (index/weight loads, then position gathers, then FMAs).
template <typename S>
__device__ __forceinline__ void one_ring_covariance(int i, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                    const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                    const Vec4T<S> *__restrict__ cur4, S cov[9]) {
    constexpr int CH = GatherChunk<S>::value;
    const int k0 = rowptr[i], k1 = rowptr[i + 1];
    const Vec4T<S> pi = load4<S>(&rest4[i]);
    const Vec4T<S> ci = load4<S>(&cur4[i]);
#pragma unroll
    for (int c = 0; c < 9; ++c) cov[c] = S(0);
    for (int k = k0; k < k1; k += CH) {
        int j[CH];
        S w[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const bool valid = k + u < k1;
            j[u] = valid ? __ldg(&colidx[k + u]) : i;
            w[u] = valid ? __ldg(&weight[k + u]) : S(0);
        }
        Vec4T<S> pj[CH], cj[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) { pj[u] = load4<S>(&rest4[j[u]]); cj[u] = load4<S>(&cur4[j[u]]); }
#if ARAP_LOCAL_GATE
        {
            const S gate = gather_gate(pj) + gather_gate(cj);
#pragma unroll
            for (int u = 0; u < CH; ++u) w[u] += gate;
        }
#endif
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const S ex = w[u] * (pi.x - pj[u].x), ey = w[u] * (pi.y - pj[u].y), ez = w[u] * (pi.z - pj[u].z);
            const S dx = ci.x - cj[u].x, dy = ci.y - cj[u].y, dz = ci.z - cj[u].z;
            cov[0] += ex * dx; cov[1] += ex * dy; cov[2] += ex * dz;
            cov[3] += ey * dx; cov[4] += ey * dy; cov[5] += ey * dz;
            cov[6] += ez * dx; cov[7] += ez * dy; cov[8] += ez * dz;
        }
    }
}

// Local step, hot kernel: R_i from the previous iteration seeds a Newton iteration on SO(3) (see arap_math.cuh).
// Vertices whose Newton iteration does not certify (far from the previous rotation, degenerate covariance) are appended
// to `redo_list`; local_step_redo_kernel gives them the full Jacobi SVD. Keeping the SVD out of this kernel keeps its
// register count -- and with it the occupancy of the 99.9 % case -- down.
template <typename S>
__global__ void __launch_bounds__(kBlock, ARAP_LOCAL_MIN_BLOCKS) local_step_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                            const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                            const Vec4T<S> *__restrict__ cur4, Vec4T<S> *__restrict__ quat,
                                                            int *__restrict__ redo_list, int *__restrict__ redo_count) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    S cov[9];
#if defined(ARAP_DIAG_LOCAL) && ARAP_DIAG_LOCAL == 2          /* diagnostic build: no gathers, compute only */
    {
        const Vec4T<S> pi = load4<S>(&rest4[i]);
        const Vec4T<S> ci = load4<S>(&cur4[i]);
        cov[0] = pi.x + S(2); cov[1] = pi.y * S(0.1); cov[2] = pi.z * S(0.1); cov[3] = ci.y * S(0.1); cov[4] = ci.x + S(2.5);
        cov[5] = ci.z * S(0.1); cov[6] = pi.z * S(0.05); cov[7] = ci.y * S(0.05); cov[8] = S(3) + pi.y;
    }
#else
    one_ring_covariance<S>(i, rowptr, colidx, weight, rest4, cur4, cov);
#endif
    const Vec4T<S> qprev = load4<S>(&quat[i]);
    const S qp[4] = {qprev.x, qprev.y, qprev.z, qprev.w};
    S q[4];
#if defined(ARAP_DIAG_LOCAL) && ARAP_DIAG_LOCAL == 1          /* diagnostic build: gathers only, no rotation extraction */
    store4<S>(&quat[i], cov[0] + cov[4] + cov[8] + qp[0], cov[1] + cov[2] + cov[3], cov[5] + cov[6], cov[7]);
    (void)q; (void)redo_list; (void)redo_count;
#else
    if (rotation_from_covariance_newton_only<S>(cov, qp, q)) store4<S>(&quat[i], q[0], q[1], q[2], q[3]);
    else redo_list[atomicAdd(redo_count, 1)] = i;
#endif
}

// Jacobi SVD (reference arap.h:376-382) for the vertices the Newton kernel handed back. Grid-stride over the list.
template <typename S>
__global__ void __launch_bounds__(kBlock) local_step_redo_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                 const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                                 const Vec4T<S> *__restrict__ cur4, Vec4T<S> *__restrict__ quat,
                                                                 const int *__restrict__ redo_list, int *__restrict__ redo_count,
                                                                 unsigned *__restrict__ done_counter) {
    pdl_enter();
    const int count = *redo_count;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const int i = redo_list[t];
        S cov[9];
        one_ring_covariance<S>(i, rowptr, colidx, weight, rest4, cur4, cov);
        S q[4];
        rotation_from_covariance<S>(cov, q);
        store4<S>(&quat[i], q[0], q[1], q[2], q[3]);
    }
    // the last CTA to finish resets the list for the next local step
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) { *redo_count = 0; *done_counter = 0u; }
}

// =================================================================================================
// Global step, part 1 (reference arap.h:393-414 + the residual of L p' = b at the current p').
//   rhs_i = sum_j (w_ij/2) (R_i + R_j) (p_i - p_j)                     (arap.h:406-413)
//   r_i   = rhs_i - sum_j w_ij (p'_i - p'_j)      for free i; 0 for constrained i.
// The second sum runs over ALL neighbours with constrained p'_j = their targets, which is exactly
// bFixed (arap.h:327) moved back to the left-hand side, so no separate bFixed array is read.
// Also starts the CG: x = 0, d = z = r / L_ii, rho = r.z, and the reference norm |rhs|^2.
// =================================================================================================
struct CgScalars {
    double rho[3];       // r.z per coordinate (of the previous iteration once gamma[] holds the current one)
    double alpha[3];
    double beta[3];
    double gamma[3];     // multigrid CG: r.z of the V-cycle just finished (becomes rho when the iteration's alpha is made)
    double rr;           // |r|^2 over the three coordinates
    double ref2;         // |rhs|^2 over the three coordinates
    double tol2;         // tolerance^2 on |r| / |rhs| (0 = residual criterion off)
    double z8_tol;       // multigrid only: stop when sum_i (|z_i| / length)^8 <= z8_tol, z = M^-1 r (0 = off); see cg_finalize
    double inv_len2;     // 1 / length^2, length = bounding-box diagonal of the rest pose
    double z8;           // last value of that sum (diagnostics)
    double red[8];       // partitioned mode: this rank's partial sums, all-reduced in place before cg_finalize_kernel
    int converged;
    int iterations;
    int distributed;     // != 0: reduction kernels only deposit their sums in red[]; cg_finalize_kernel finishes the stage
    int max_iterations;  // per global step (the device-side loop of the step graph stops there)
    // bookkeeping of the global steps since the host last read it (one host round trip per arap_iterate, not per step)
    long long iterations_total;
    long long body_runs_total;   // executions of the CG iteration's kernel sequence (iterations + the passes that only noticed convergence)
    int steps, unconverged_steps, first_step_iterations, pad;
};

// What each grid-wide reduction of the CG turns into. On one GPU the last CTA of the reducing kernel calls this
// directly; in partitioned mode the sums first go through an all-reduce over the ranks (see partition.cuh).
//
// Multigrid-preconditioned CG in the single-reduction form (Chronopoulos & Gear 1989): per iteration
//   z = M^-1 r (V-cycle) ; w = A z ; gamma = r.z ; delta = z.w ;
//   beta = gamma / gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old) ;
//   d = z + beta d ; s = w + beta s ; x += alpha d ; r -= alpha s
// so the only matrix-vector product gathers the fp32 V-cycle output (one aligned 16-byte load per neighbour), every
// vector update is ONE streaming kernel, and all scalars of an iteration come from sums that can travel in one all-reduce
// (gamma, the position-error norm, delta, and |r|^2 of the previous update: CG_STAGE_MERGED).
enum CgStage { CG_STAGE_START_JACOBI = 0, CG_STAGE_START_MG, CG_STAGE_ALPHA, CG_STAGE_UPDATE_JACOBI, CG_STAGE_UPDATE_MG, CG_STAGE_GAMMA,
               CG_STAGE_DELTA, CG_STAGE_MERGED };

__device__ __forceinline__ void cg_finalize(CgScalars *cg, int stage, const double *t) {
    switch (stage) {
        case CG_STAGE_START_JACOBI:      // t = rho x,y,z ; |r|^2 ; |rhs|^2
        case CG_STAGE_START_MG:
            for (int c = 0; c < 3; ++c) {
                cg->rho[c] = (stage == CG_STAGE_START_MG) ? 0.0 : t[c];
                cg->alpha[c] = 0.0;
                cg->beta[c] = 0.0;
            }
            cg->rr = t[3];
            cg->ref2 = t[4];
            cg->iterations = 0;
            cg->converged = (t[3] <= cg->tol2 * t[4]) ? 1 : 0;
            cg->red[7] = -1.0;           // CG_STAGE_MERGED: no |r|^2 of a previous update yet
            break;
        case CG_STAGE_ALPHA:             // t = d.Ad per coordinate
            for (int c = 0; c < 3; ++c) cg->alpha[c] = (t[c] > 0.0) ? cg->rho[c] / t[c] : 0.0;
            break;
        case CG_STAGE_UPDATE_JACOBI:     // t = r.D^-1 r per coordinate ; |r|^2
            for (int c = 0; c < 3; ++c) {
                cg->beta[c] = (cg->rho[c] > 0.0) ? t[c] / cg->rho[c] : 0.0;
                cg->rho[c] = t[c];
            }
            cg->rr = t[3];
            cg->iterations += 1;
            if (t[3] <= cg->tol2 * cg->ref2) cg->converged = 1;
            break;
        case CG_STAGE_UPDATE_MG:         // t = |r|^2
            cg->rr = t[0];
            if (t[0] <= cg->tol2 * cg->ref2) cg->converged = 1;
            break;
        case CG_STAGE_GAMMA:             // t = r.z per coordinate ; sum_i (|z_i| / length)^8
            for (int c = 0; c < 3; ++c) cg->gamma[c] = t[c];
            // z = M^-1 r with M^-1 one multigrid V-cycle is, up to the quality of the preconditioner (~ +-40 %), the ERROR
            // A^-1 r of the current iterate, in position units. Its 8-norm is a smooth stand-in for the largest per-vertex
            // error (max <= 8-norm <= V^(1/8) max) that can be summed -- and all-reduced -- like every other CG scalar.
            // A residual tolerance cannot play this role: the same |r|/|rhs| means 8e-10 of the bounding box on a regular
            // sphere and 4e-5 on a Delaunay patch with sliver triangles (weights from 5e-11 to 3e3).
            // (An energy rule on r.z ~ |e|_A^2, the energy excess of the iterate, was tried on top of this one and never bit:
            // the energy deviation that remains after 20 ARAP iterations comes from the drift of the ARAP state, i.e. from the
            // position error, not from the last solve -- profiles/r01_h_stopping_rule.txt.)
            cg->z8 = t[3];
            if (cg->z8_tol > 0.0 && t[3] <= cg->z8_tol) cg->converged = 1;
            break;
        case CG_STAGE_DELTA:             // t = z.w per coordinate -> beta, alpha of this iteration
            for (int c = 0; c < 3; ++c) {
                const double g = cg->gamma[c], a_old = cg->alpha[c];
                const double beta = (cg->rho[c] > 0.0 && a_old != 0.0) ? g / cg->rho[c] : 0.0;
                const double denom = t[c] - (beta != 0.0 ? beta * g / a_old : 0.0);      // = d.Ad of the new direction
                cg->beta[c] = beta;
                cg->alpha[c] = (denom > 0.0) ? g / denom : 0.0;
                cg->rho[c] = g;
            }
            cg->iterations += 1;
            break;
        case CG_STAGE_MERGED:            // partitioned mode: t = gamma[3], z8, delta[3], |r|^2 after the previous update (< 0: none)
            if (t[7] >= 0.0) {
                cg->rr = t[7];
                if (t[7] <= cg->tol2 * cg->ref2) cg->converged = 1;
            }
            if (!cg->converged) {
                cg_finalize(cg, CG_STAGE_GAMMA, t);
                if (!cg->converged) cg_finalize(cg, CG_STAGE_DELTA, t + 4);
            }
            break;
    }
}

// `slot`: where a distributed stage keeps its sums in red[] until the all-reduce (CG_STAGE_MERGED packs three kernels' sums)
template <int N>
__device__ __forceinline__ void cg_finish_reduction(CgScalars *cg, int stage, const double (&total)[N], int slot = 0) {
    if (cg->distributed) {
#pragma unroll
        for (int c = 0; c < N; ++c) cg->red[slot + c] = total[c];
    } else {
        cg_finalize(cg, stage, total);
    }
}

__global__ void cg_finalize_kernel(CgScalars *cg, int stage) {
    pdl_enter();
    if (cg->converged && stage != CG_STAGE_START_JACOBI && stage != CG_STAGE_START_MG) return;
    cg_finalize(cg, stage, cg->red);
}

// MG = false: Jacobi-preconditioned start (d = z = D^-1 r, rho = r.z).
// MG = true : multigrid start (rho = 0 so the first beta is 0, x0 = omega0 D^-1 r feeds the first V-cycle).
// (A variant with 8 lanes per vertex and a shuffle reduction of the nine partial sums measured 4x slower: 383 us.)
template <typename S, bool MG>
__global__ void __launch_bounds__(kBlock, ARAP_RHS_MIN_BLOCKS) rhs_residual_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                              const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                              const Vec4T<S> *__restrict__ cur4, const Vec4T<S> *__restrict__ quat,
                                                              const double *__restrict__ inv_diag, double omega0,
                                                              Vec3d *__restrict__ r_out, Vec3d *__restrict__ d_out,
                                                              Vec3d *__restrict__ x_out, float4 *__restrict__ x0_out,
                                                              double *__restrict__ partials, unsigned *__restrict__ counter,
                                                              CgScalars *__restrict__ cg) {
    pdl_enter();
    double red[5] = {0, 0, 0, 0, 0};   // rho x,y,z ; rr ; ref2
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Vec4T<S> pi = load4<S>(&rest4[i]);
        Vec3d r = {0, 0, 0}, z = {0, 0, 0};
        if (pi.w != S(0)) {
            const Vec4T<S> ci = load4<S>(&cur4[i]);
            const Vec4T<S> qi = load4<S>(&quat[i]);
            double rot_j[3] = {0, 0, 0};     // sum_j (w/2) R_j e_ij
            double se[3] = {0, 0, 0};        // sum_j (w/2) e_ij
            double lap[3] = {0, 0, 0};       // sum_j w (p'_i - p'_j)
            constexpr int CH = RhsChunk<S>::value;
            const int k0 = rowptr[i], k1 = rowptr[i + 1];
            for (int k = k0; k < k1; k += CH) {
                int j[CH];
                S w[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const bool valid = k + u < k1;
                    j[u] = valid ? __ldg(&colidx[k + u]) : i;
                    w[u] = valid ? __ldg(&weight[k + u]) : S(0);
                }
                Vec4T<S> pj[CH], cj[CH], qj[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) { pj[u] = load4<S>(&rest4[j[u]]); cj[u] = load4<S>(&cur4[j[u]]); qj[u] = load4<S>(&quat[j[u]]); }
#if ARAP_RHS_GATE
                {
                    const S gate = gather_gate(pj) + gather_gate(cj) + gather_gate(qj);
#pragma unroll
                    for (int u = 0; u < CH; ++u) w[u] += gate;
                }
#endif
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const S hw = S(0.5) * w[u];
                    const S ex = hw * (pi.x - pj[u].x), ey = hw * (pi.y - pj[u].y), ez = hw * (pi.z - pj[u].z);
                    S rx, ry, rz;                                                  // R_j e_ij, quat stored as (w,x,y,z) in (.x,.y,.z,.w)
                    quat_rotate<S>(qj[u].x, qj[u].y, qj[u].z, qj[u].w, ex, ey, ez, rx, ry, rz);
                    rot_j[0] += (double)rx; rot_j[1] += (double)ry; rot_j[2] += (double)rz;
                    se[0] += (double)ex; se[1] += (double)ey; se[2] += (double)ez;
                    lap[0] += (double)w[u] * ((double)ci.x - (double)cj[u].x);
                    lap[1] += (double)w[u] * ((double)ci.y - (double)cj[u].y);
                    lap[2] += (double)w[u] * ((double)ci.z - (double)cj[u].z);
                }
            }
            double ox, oy, oz;                                                     // R_i sum_j (w/2) e_ij
            quat_rotate<double>((double)qi.x, (double)qi.y, (double)qi.z, (double)qi.w, se[0], se[1], se[2], ox, oy, oz);
            const double rhs0 = rot_j[0] + ox, rhs1 = rot_j[1] + oy, rhs2 = rot_j[2] + oz;
            r.x = rhs0 - lap[0]; r.y = rhs1 - lap[1]; r.z = rhs2 - lap[2];
            const double idg = inv_diag[i];
            z.x = r.x * idg; z.y = r.y * idg; z.z = r.z * idg;
            red[0] += r.x * z.x; red[1] += r.y * z.y; red[2] += r.z * z.z;
            red[3] += r.x * r.x + r.y * r.y + r.z * r.z;
            red[4] += rhs0 * rhs0 + rhs1 * rhs1 + rhs2 * rhs2;
        }
        r_out[i] = r;
        if (MG) {      // x, d, s need no reset: the first update of a solve overwrites them (cg_fused_update_kernel)
            x0_out[i] = make_float4((float)(omega0 * z.x), (float)(omega0 * z.y), (float)(omega0 * z.z), 0.f);
        } else {
            x_out[i] = Vec3d{0, 0, 0};
            d_out[i] = z;
        }
    }
    double total[5];
    if (grid_sum_last_block<5>(red, partials, counter, total))
        cg_finish_reduction<5>(cg, MG ? CG_STAGE_START_MG : CG_STAGE_START_JACOBI, total);
}

// =================================================================================================
// Global step, part 2: Jacobi-preconditioned CG on L (free x free), three right-hand sides at once,
// matrix-free on the one-ring CSR: (L d)_i = sum_j w_ij (d_i - d_j) for free i (d_j = 0 on constrained j).
// Replaces SimplicialLDLT::solve (reference arap.h:418-421).
// =================================================================================================
template <typename S>
__global__ void __launch_bounds__(kBlock, ARAP_SPMV_MIN_BLOCKS) cg_spmv_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                         const S *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                         const Vec3d *__restrict__ d, Vec3d *__restrict__ ad,
                                                         double *__restrict__ partials, unsigned *__restrict__ counter,
                                                         CgScalars *__restrict__ cg) {
    pdl_enter();
    if (cg->converged) return;
    double red[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Vec3d out = {0, 0, 0};
        if (free_mask[i]) {
            constexpr int CH = kSpmvChunk;
            const int k0 = rowptr[i], k1 = rowptr[i + 1];
            const Vec3d di = d[i];
            for (int k = k0; k < k1; k += CH) {
                int j[CH];
                double w[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const bool valid = k + u < k1;
                    j[u] = valid ? __ldg(&colidx[k + u]) : i;
                    w[u] = valid ? (double)__ldg(&weight[k + u]) : 0.0;
                }
                Vec3d dj[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) dj[u] = d[j[u]];
                const double gate = gather_gate(dj);
#pragma unroll
                for (int u = 0; u < CH; ++u) w[u] += gate;
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    out.x += w[u] * (di.x - dj[u].x); out.y += w[u] * (di.y - dj[u].y); out.z += w[u] * (di.z - dj[u].z);
                }
            }
            red[0] += di.x * out.x; red[1] += di.y * out.y; red[2] += di.z * out.z;
        }
        ad[i] = out;
    }
    double total[3];
    if (grid_sum_last_block<3>(red, partials, counter, total)) cg_finish_reduction<3>(cg, CG_STAGE_ALPHA, total);
}

// ---- TMA-staged variant of cg_spmv ---------------------------------------------------------------------------------
// Same arithmetic as cg_spmv_kernel. The CTA walks row tiles of kBlock rows; the tile's contiguous colidx / weight spans
// are bulk-copied (cp.async.bulk + mbarrier, see tma_stage.cuh) into one of two shared-memory stages while the
// previous tile is being processed. Tiles whose span does not fit a stage (very high valence) read the CSR from global.
constexpr int kTmaStageEntries = 1792;      // per stage: 7 entries per row on average; 2 stages x 1792 x (4 + 8) B = 43 KB
template <typename S>
constexpr size_t tma_spmv_smem_bytes() { return 2 * (size_t)kTmaStageEntries * (sizeof(int) + sizeof(S)) + 64; }

template <typename S>
__global__ void __launch_bounds__(kBlock) cg_spmv_tma_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                             const S *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                             const Vec3d *__restrict__ d, Vec3d *__restrict__ ad,
                                                             double *__restrict__ partials, unsigned *__restrict__ counter,
                                                             CgScalars *__restrict__ cg) {
    pdl_enter();
    if (cg->converged) return;
    extern __shared__ __align__(16) unsigned char tma_smem[];
    S *const s_w0 = reinterpret_cast<S *>(tma_smem);                                                       // [2][kTmaStageEntries]
    int *const s_c0 = reinterpret_cast<int *>(tma_smem + 2 * (size_t)kTmaStageEntries * sizeof(S));        // [2][kTmaStageEntries]
    uint64_t *bar = reinterpret_cast<uint64_t *>(tma_smem + 2 * (size_t)kTmaStageEntries * (sizeof(int) + sizeof(S)));
    int *s_meta = reinterpret_cast<int *>(bar + 2);          // [stage][2]: first staged entry (k0a), staged flag
    const int n_tiles = (n + kBlock - 1) / kBlock;
    if (threadIdx.x == 0) {
        tma::mbar_init(&bar[0], 1);
        tma::mbar_init(&bar[1], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int stage) {                  // one thread: start the bulk copies of `tile` into `stage`
        const int r0 = tile * kBlock, r1 = min(n, r0 + kBlock);
        const int k0 = rowptr[r0], k1 = rowptr[r1];
        const int k0a = k0 & ~3, cnt = ((k1 + 3) & ~3) - k0a;
        const bool staged = cnt > 0 && cnt <= kTmaStageEntries;
        s_meta[2 * stage] = k0a;
        s_meta[2 * stage + 1] = staged ? 1 : 0;
        if (staged) {
            tma::mbar_arrive_expect_tx(&bar[stage], (uint32_t)cnt * (uint32_t)(sizeof(int) + sizeof(S)));
            tma::bulk_g2s(s_c0 + stage * kTmaStageEntries, colidx + k0a, (uint32_t)cnt * (uint32_t)sizeof(int), &bar[stage]);
            tma::bulk_g2s(s_w0 + stage * kTmaStageEntries, weight + k0a, (uint32_t)cnt * (uint32_t)sizeof(S), &bar[stage]);
        } else {
            tma::mbar_arrive_expect_tx(&bar[stage], 0u);
        }
    };
    int stage = 0;
    uint32_t phase_bits = 0u;                                // bit s = parity to wait for on stage s
    if (threadIdx.x == 0 && (int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    __syncthreads();
    double red[3] = {0, 0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < n_tiles) issue(next, stage ^ 1);
        tma::mbar_wait(&bar[stage], (phase_bits >> stage) & 1u);
        phase_bits ^= 1u << stage;
        const int k0a = s_meta[2 * stage];
        const bool staged = s_meta[2 * stage + 1] != 0;
        const int i = tile * kBlock + threadIdx.x;
        if (i < n) {
            Vec3d out = {0, 0, 0};
            if (free_mask[i]) {
                const int ka = rowptr[i], kb = rowptr[i + 1];
                const Vec3d di = d[i];
                const int *cc = staged ? (s_c0 + stage * kTmaStageEntries - k0a) : colidx;      // both indexed by the global entry number
                const S *ww = staged ? (s_w0 + stage * kTmaStageEntries - k0a) : weight;
                for (int kk = ka; kk < kb; ++kk) {
                    const int j = cc[kk];
                    const double w = (double)ww[kk];
                    const Vec3d dj = d[j];
                    out.x += w * (di.x - dj.x); out.y += w * (di.y - dj.y); out.z += w * (di.z - dj.z);
                }
                red[0] += di.x * out.x; red[1] += di.y * out.y; red[2] += di.z * out.z;
            }
            ad[i] = out;
        }
        __syncthreads();        // everyone is done with this stage before the copy of tile + 2*grid lands in it
        stage ^= 1;
    }
    double total[3];
    if (grid_sum_last_block<3>(red, partials, counter, total)) cg_finish_reduction<3>(cg, CG_STAGE_ALPHA, total);
}

__global__ void __launch_bounds__(kBlock) cg_update_kernel(int n, const double *__restrict__ inv_diag, const Vec3d *__restrict__ d,
                                                           const Vec3d *__restrict__ ad, Vec3d *__restrict__ x, Vec3d *__restrict__ r,
                                                           double *__restrict__ partials, unsigned *__restrict__ counter,
                                                           CgScalars *__restrict__ cg) {
    pdl_enter();
    if (cg->converged) return;
    double red[4] = {0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double a0 = cg->alpha[0], a1 = cg->alpha[1], a2 = cg->alpha[2];
        const Vec3d di = d[i], adi = ad[i];
        Vec3d xi = x[i], ri = r[i];
        xi.x += a0 * di.x; xi.y += a1 * di.y; xi.z += a2 * di.z;
        ri.x -= a0 * adi.x; ri.y -= a1 * adi.y; ri.z -= a2 * adi.z;
        x[i] = xi; r[i] = ri;
        const double idg = inv_diag[i];
        red[0] += ri.x * ri.x * idg; red[1] += ri.y * ri.y * idg; red[2] += ri.z * ri.z * idg;
        red[3] += ri.x * ri.x + ri.y * ri.y + ri.z * ri.z;
    }
    double total[4];
    if (grid_sum_last_block<4>(red, partials, counter, total)) cg_finish_reduction<4>(cg, CG_STAGE_UPDATE_JACOBI, total);
}

// d = z + beta d. Launched after cg_update; skipped (like everything else) once converged.
__global__ void __launch_bounds__(kBlock) cg_direction_kernel(int n, const double *__restrict__ inv_diag, const Vec3d *__restrict__ r,
                                                              Vec3d *__restrict__ d, const CgScalars *__restrict__ cg) {
    pdl_enter();
    if (cg->converged) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double idg = inv_diag[i];
    const Vec3d ri = r[i];
    Vec3d di = d[i];
    di.x = ri.x * idg + cg->beta[0] * di.x;
    di.y = ri.y * idg + cg->beta[1] * di.y;
    di.z = ri.z * idg + cg->beta[2] * di.z;
    d[i] = di;
}


// ---- multigrid-preconditioned CG, single-reduction form (see CgStage above) -----------------------------------------
// w = A z on the free rows, A matrix-free on the one-ring CSR with the exact (handle precision) weights; z is the fp32
// V-cycle output, so every neighbour is ONE aligned 16-byte gather; the sum runs in fp64. Fused delta = z . w.
template <typename S>
__global__ void __launch_bounds__(kBlock, ARAP_SPMV_MIN_BLOCKS) cg_spmv_z_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                           const S *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                           const float4 *__restrict__ z, Vec3d *__restrict__ w_out,
                                                           double *__restrict__ partials, unsigned *__restrict__ counter,
                                                           CgScalars *__restrict__ cg) {
    pdl_enter();
    if (cg->converged) return;
    double red[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Vec3d out = {0, 0, 0};
        if (free_mask[i]) {
            constexpr int CH = kSpmvChunk;
            const int k0 = rowptr[i], k1 = rowptr[i + 1];
            const float4 zi = z[i];
            for (int k = k0; k < k1; k += CH) {
                int j[CH];
                double w[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const bool valid = k + u < k1;
                    j[u] = valid ? __ldg(&colidx[k + u]) : i;
                    w[u] = valid ? (double)__ldg(&weight[k + u]) : 0.0;
                }
                float4 zj[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) zj[u] = z[j[u]];
                float gate = zj[0].x;
#pragma unroll
                for (int u = 1; u < CH; ++u) gate += zj[u].x;
                const double g = (double)(0.0f * gate);      // load-batching gate, see gather_gate()
#pragma unroll
                for (int u = 0; u < CH; ++u) w[u] += g;
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    out.x += w[u] * ((double)zi.x - (double)zj[u].x);
                    out.y += w[u] * ((double)zi.y - (double)zj[u].y);
                    out.z += w[u] * ((double)zi.z - (double)zj[u].z);
                }
            }
            red[0] += (double)zi.x * out.x; red[1] += (double)zi.y * out.y; red[2] += (double)zi.z * out.z;
        }
        w_out[i] = out;
    }
    double total[3];
    if (grid_sum_last_block<3>(red, partials, counter, total)) cg_finish_reduction<3>(cg, CG_STAGE_DELTA, total, 4);
}

__device__ __forceinline__ double pick3(int c, double a0, double a1, double a2) { return c == 0 ? a0 : (c == 1 ? a1 : a2); }

// d = z + beta d ; s = w + beta s ; x += alpha d ; r -= alpha s ; x0 = omega_0 D^-1 r (the next V-cycle's pre-smoothed fine
// iterate) ; |r|^2. The CG vectors are flat arrays of 3V doubles streamed with one coalesced 16-byte access per thread and
// array (element e = vertex e/3, coordinate e%3). The first update of a solve (iterations == 1) overwrites d, s and x
// instead of accumulating, so nobody has to zero them. `loop` != 0: this kernel closes the body of the step graph's
// device-side WHILE node and decides whether the body runs again.
__global__ void __launch_bounds__(kBlock) cg_fused_update_kernel(int n3, const double *__restrict__ inv_diag, double omega0,
                                                                 const float *__restrict__ z, const double *__restrict__ w,
                                                                 double *__restrict__ d, double *__restrict__ s, double *__restrict__ x,
                                                                 double *__restrict__ r, float *__restrict__ x0,
                                                                 double *__restrict__ partials, unsigned *__restrict__ counter,
                                                                 CgScalars *__restrict__ cg, unsigned long long loop) {
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) cg->body_runs_total += 1;
    if (cg->converged) {
        if (loop && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional((cudaGraphConditionalHandle)loop, 0u);
        return;
    }
    double red[1] = {0};
    const bool first = cg->iterations == 1;
    const double al0 = cg->alpha[0], al1 = cg->alpha[1], al2 = cg->alpha[2];
    const double be0 = cg->beta[0], be1 = cg->beta[1], be2 = cg->beta[2];
    for (int e = 2 * (blockIdx.x * blockDim.x + threadIdx.x); e < n3; e += 2 * gridDim.x * blockDim.x) {
        if (e + 1 < n3) {
            const int c0 = e % 3, c1 = (e + 1) % 3;
            const int v0 = e / 3, v1 = (e + 1) / 3;
            const double a0 = pick3(c0, al0, al1, al2), a1 = pick3(c1, al0, al1, al2);
            const double b0 = pick3(c0, be0, be1, be2), b1 = pick3(c1, be0, be1, be2);
            const double z0 = (double)z[4 * v0 + c0], z1 = (double)z[4 * v1 + c1];      // z is a float4 per vertex
            const double2 wv = *reinterpret_cast<const double2 *>(w + e);
            double2 dv, sv, xv;
            if (first) {
                dv = make_double2(z0, z1);
                sv = wv;
                xv = make_double2(a0 * z0, a1 * z1);
            } else {
                dv = *reinterpret_cast<const double2 *>(d + e);
                sv = *reinterpret_cast<const double2 *>(s + e);
                xv = *reinterpret_cast<const double2 *>(x + e);
                dv.x = z0 + b0 * dv.x; dv.y = z1 + b1 * dv.y;
                sv.x = wv.x + b0 * sv.x; sv.y = wv.y + b1 * sv.y;
                xv.x += a0 * dv.x; xv.y += a1 * dv.y;
            }
            double2 rv = *reinterpret_cast<const double2 *>(r + e);
            rv.x -= a0 * sv.x; rv.y -= a1 * sv.y;
            *reinterpret_cast<double2 *>(d + e) = dv;
            *reinterpret_cast<double2 *>(s + e) = sv;
            *reinterpret_cast<double2 *>(x + e) = xv;
            *reinterpret_cast<double2 *>(r + e) = rv;
            x0[4 * v0 + c0] = (float)(omega0 * inv_diag[v0] * rv.x);
            x0[4 * v1 + c1] = (float)(omega0 * inv_diag[v1] * rv.y);
            red[0] += rv.x * rv.x + rv.y * rv.y;
        } else {
            const int c0 = e % 3, v0 = e / 3;
            const double a0 = pick3(c0, al0, al1, al2), b0 = pick3(c0, be0, be1, be2);
            const double z0 = (double)z[4 * v0 + c0];
            const double dv = first ? z0 : z0 + b0 * d[e];
            const double sv = first ? w[e] : w[e] + b0 * s[e];
            const double xv = first ? a0 * dv : x[e] + a0 * dv;
            const double rv = r[e] - a0 * sv;
            d[e] = dv; s[e] = sv; x[e] = xv; r[e] = rv;
            x0[4 * v0 + c0] = (float)(omega0 * inv_diag[v0] * rv);
            red[0] += rv * rv;
        }
    }
    double total[1];
    if (grid_sum_last_block<1>(red, partials, counter, total)) {
        cg_finish_reduction<1>(cg, CG_STAGE_UPDATE_MG, total, 7);
        if (loop) cudaGraphSetConditional((cudaGraphConditionalHandle)loop, (!cg->converged && cg->iterations < cg->max_iterations) ? 1u : 0u);
    }
}

// p' += x on the free vertices (the scatter of arap.h:423-428); constrained vertices keep their targets.
// Also books the finished global step (iteration counts, convergence) in CgScalars, so that the host reads the solver's
// statistics once per arap_iterate instead of once per step.
template <typename S>
__global__ void __launch_bounds__(kBlock) apply_update_kernel(int n, const unsigned char *__restrict__ free_mask, const Vec3d *__restrict__ x,
                                                              Vec4T<S> *__restrict__ cur4, CgScalars *__restrict__ cg, int book) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int its = cg->iterations;
    if (book && i == 0) {
        if (cg->steps == 0) cg->first_step_iterations = its;
        cg->steps += 1;
        cg->iterations_total += its;
        if (!cg->converged) cg->unconverged_steps += 1;
    }
    if (i >= n) return;
    if (its == 0) return;              // the start residual already met the stopping rule: x was never written
    if (!free_mask[i]) return;
    const Vec4T<S> c = cur4[i];
    const Vec3d xi = x[i];
    store4<S>(&cur4[i], (S)((double)c.x + xi.x), (S)((double)c.y + xi.y), (S)((double)c.z + xi.z), c.w);
}

// =================================================================================================
// ARAP energy (not in the reference): E = sum_i sum_j w_ij |(p'_i-p'_j) - R_i (p_i-p_j)|^2, fp64.
// =================================================================================================
template <typename S>
__global__ void __launch_bounds__(kBlock) energy_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                        const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                        const Vec4T<S> *__restrict__ cur4, const Vec4T<S> *__restrict__ quat,
                                                        double *__restrict__ partials, unsigned *__restrict__ counter,
                                                        double *__restrict__ energy_out) {
    double red[1] = {0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Vec4T<S> pi = load4<S>(&rest4[i]);
        const Vec4T<S> ci = load4<S>(&cur4[i]);
        const Vec4T<S> qi = load4<S>(&quat[i]);
        double ri[9];
        quat_to_matrix<double>((double)qi.x, (double)qi.y, (double)qi.z, (double)qi.w, ri);
        double e = 0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colidx[k];
            const Vec4T<S> pj = load4<S>(&rest4[j]);
            const Vec4T<S> cj = load4<S>(&cur4[j]);
            const double ex = (double)pi.x - (double)pj.x, ey = (double)pi.y - (double)pj.y, ez = (double)pi.z - (double)pj.z;
            const double dx = (double)ci.x - (double)cj.x - (ri[0] * ex + ri[1] * ey + ri[2] * ez);
            const double dy = (double)ci.y - (double)cj.y - (ri[3] * ex + ri[4] * ey + ri[5] * ez);
            const double dz = (double)ci.z - (double)cj.z - (ri[6] * ex + ri[7] * ey + ri[8] * ez);
            e += (double)weight[k] * (dx * dx + dy * dy + dz * dz);
        }
        red[0] += e;
    }
    double total[1];
    if (grid_sum_last_block<1>(red, partials, counter, total)) *energy_out = total[0];
}

// _b of the reference (arap.h:393-414: bFixed + the rotated edge sums) for the free vertices, in free-index order, rebuilt from
// what the hot kernel computed: rhs_residual_kernel leaves r = b - L p' (with the constrained neighbours' targets folded into
// the left-hand side), so b_i = r_i + sum_j w_ij p'_i - sum_{j free} w_ij p'_j.
template <typename S>
__global__ void __launch_bounds__(kBlock) export_rhs_kernel(int n, const int *__restrict__ perm, const int *__restrict__ free_idx_user,
                                                            const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                            const S *__restrict__ weight, const unsigned char *__restrict__ free_mask,
                                                            const Vec4T<S> *__restrict__ cur4, const Vec3d *__restrict__ r,
                                                            double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !free_mask[i]) return;
    const Vec4T<S> ci = cur4[i];
    Vec3d b = r[i];
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        const double w = (double)weight[k];
        b.x += w * (double)ci.x; b.y += w * (double)ci.y; b.z += w * (double)ci.z;
        if (free_mask[j]) {
            const Vec4T<S> cj = cur4[j];
            b.x -= w * (double)cj.x; b.y -= w * (double)cj.y; b.z -= w * (double)cj.z;
        }
    }
    const size_t f = (size_t)free_idx_user[perm[i]];
    out[3 * f] = b.x; out[3 * f + 1] = b.y; out[3 * f + 2] = b.z;
}

// ---- readout helpers (internal order -> the user's numbering) -------------------------------------------------
template <typename S, typename T>
__global__ void __launch_bounds__(kBlock) export_positions_kernel(int n, const int *__restrict__ perm, const Vec4T<S> *__restrict__ cur4,
                                                                  T *__restrict__ out_xyz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec4T<S> c = cur4[i];
    const size_t u = (size_t)perm[i];
    out_xyz[3 * u] = (T)c.x; out_xyz[3 * u + 1] = (T)c.y; out_xyz[3 * u + 2] = (T)c.z;
}

template <typename S>
__global__ void __launch_bounds__(kBlock) export_rotations_kernel(int n, const int *__restrict__ perm, const Vec4T<S> *__restrict__ quat,
                                                                  S *__restrict__ rot9) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec4T<S> q = quat[i];
    S r[9];
    quat_to_matrix<S>(q.x, q.y, q.z, q.w, r);
    const size_t u = (size_t)perm[i];
#pragma unroll
    for (int k = 0; k < 9; ++k) rot9[9 * u + k] = r[k];
}

// =================================================================================================
// Viewer interop (reference examples/osg_viewer.cpp:45-72): what the viewer's update callback needs every frame --
// float positions and per-vertex normals -- produced on the device, optionally straight into a caller-owned DEVICE buffer
// (a mapped OpenGL vertex buffer), so a frame needs no host round trip of the geometry.
// Normals as OpenMesh's update_normals() makes them: unit face normals, vertex normal = normalised sum of the incident ones.
// =================================================================================================
// vertex -> incident faces, built once per handle: counts, (scan), fill with cursors, rows sorted by face id
__global__ void __launch_bounds__(kBlock) vf_count_kernel(int n_faces, const int *__restrict__ faces, int *__restrict__ count) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(&count[faces[3 * (size_t)f + k]], 1);
}
__global__ void __launch_bounds__(kBlock) vf_fill_kernel(int n_faces, const int *__restrict__ faces, const int *__restrict__ vf_ptr,
                                                         int *__restrict__ cursor, int *__restrict__ vf_face) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int v = faces[3 * (size_t)f + k];
        vf_face[vf_ptr[v] + atomicAdd(&cursor[v], 1)] = f;
    }
}
__global__ void __launch_bounds__(kBlock) vf_sort_kernel(int n_vertices, const int *__restrict__ vf_ptr, int *__restrict__ vf_face) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vertices) return;
    const int lo = vf_ptr[v], hi = vf_ptr[v + 1];
    for (int a = lo + 1; a < hi; ++a) {
        const int f = vf_face[a];
        int b = a - 1;
        while (b >= lo && vf_face[b] > f) { vf_face[b + 1] = vf_face[b]; --b; }
        vf_face[b + 1] = f;
    }
}
// unit face normals of the current pose ((p1 - p0) x (p2 - p0), the mesh's own winding); degenerate faces get (0,0,0)
template <typename S>
__global__ void __launch_bounds__(kBlock) face_normals_kernel(int n_faces, const int *__restrict__ faces, const int *__restrict__ iperm,
                                                              const Vec4T<S> *__restrict__ cur4, float4 *__restrict__ face_normal) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const Vec4T<S> a = cur4[iperm[faces[3 * (size_t)f]]], b = cur4[iperm[faces[3 * (size_t)f + 1]]], c = cur4[iperm[faces[3 * (size_t)f + 2]]];
    const double ux = (double)b.x - (double)a.x, uy = (double)b.y - (double)a.y, uz = (double)b.z - (double)a.z;
    const double vx = (double)c.x - (double)a.x, vy = (double)c.y - (double)a.y, vz = (double)c.z - (double)a.z;
    const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    const double inv = len > 0.0 ? 1.0 / len : 0.0;
    face_normal[f] = make_float4((float)(nx * inv), (float)(ny * inv), (float)(nz * inv), 0.f);
}
// float positions + vertex normals in the caller's vertex numbering (one thread per user vertex)
template <typename S>
__global__ void __launch_bounds__(kBlock) render_buffers_kernel(int n_vertices, const int *__restrict__ iperm, const Vec4T<S> *__restrict__ cur4,
                                                                const int *__restrict__ vf_ptr, const int *__restrict__ vf_face,
                                                                const float4 *__restrict__ face_normal, float *__restrict__ out_pos,
                                                                float *__restrict__ out_nrm) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vertices) return;
    const Vec4T<S> p = cur4[iperm[v]];
    out_pos[3 * (size_t)v] = (float)p.x; out_pos[3 * (size_t)v + 1] = (float)p.y; out_pos[3 * (size_t)v + 2] = (float)p.z;
    if (!out_nrm) return;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for (int k = vf_ptr[v]; k < vf_ptr[v + 1]; ++k) { const float4 n = face_normal[vf_face[k]]; nx += n.x; ny += n.y; nz += n.z; }
    const float len = sqrtf(nx * nx + ny * ny + nz * nz);
    const float inv = len > 0.f ? 1.f / len : 0.f;
    out_nrm[3 * (size_t)v] = nx * inv; out_nrm[3 * (size_t)v + 1] = ny * inv; out_nrm[3 * (size_t)v + 2] = nz * inv;
}

}  // namespace arap
