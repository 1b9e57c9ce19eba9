// mg_setup_device.cuh -- the smoothed-aggregation setup of mg_setup.cpp as CUDA kernels (the role of the reference's
// setupLinearSystem + SimplicialLDLT::compute, arap.h:292-340: a one-off analysis of L per change of the constrained set).
//
// Same algorithm as the host version (Vanek, Mandel, Brezina): strength graph |a_ij| >= theta sqrt(a_ii a_jj), aggregates =
// root + strong neighbours, P = (I - omega D^-1 A) T, R = P^T, A_c = R A P, all in fp64. What differs is how the aggregates
// are found: the host walks the vertices one after the other (greedy, in Morton order); here the roots are a maximal
// independent set of the SQUARE of the strength graph (no two roots within distance 2 -- exactly the property the greedy
// walk guarantees), found in a handful of parallel rounds: an undecided vertex becomes a root when its hashed key is the
// largest among the undecided vertices within distance 2. Every sparse product is built row by row, one thread per row,
// into a small sorted accumulator in the thread's local memory: rows come out sorted by column and every sum runs in a
// fixed order, so the hierarchy is bit-reproducible from run to run and from rank to rank.
#pragma once

#include "device_utils.cuh"

namespace arap {
namespace mgdev {

constexpr int kAggUnset = -2;

__device__ __forceinline__ unsigned hash_u32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// key of a vertex in the root election, largest wins: [number of strong neighbours already decided : 8][hash : 24][index : 32].
// The index makes keys unique; the hash spreads the first round's roots; the count makes later rounds elect roots right at
// the rim of what is already covered, the way the sequential greedy walk of the host version packs its aggregates (measured with
// a numpy model of this election, CG iterations to 1e-8 on a 300 x 300 grid / an 81k icosphere: greedy 15 / 16, hash alone
// 18 / 17, with the count 16 / 16 -- profiles/r02_experiments.txt). 0 = "not a candidate".
__device__ __forceinline__ unsigned long long root_key(int i, int decided_neighbours) {
    const unsigned long long c = (unsigned long long)(decided_neighbours > 255 ? 255 : decided_neighbours);
    return (c << 56) | ((unsigned long long)((hash_u32((unsigned)i) & 0xffffffu) | 1u) << 32) | (unsigned)i;
}

// ---- level 0: L = D - W on the free rows as an explicit CSR (full vertex index space, constrained rows empty) ---------
// pass 1 (fill == 0): row lengths; pass 2: entries sorted by column with the diagonal in place.
template <typename S>
__global__ void __launch_bounds__(kBlock) level0_rows_kernel(int n, int n_cols_valid, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                             const S *__restrict__ w, const unsigned char *__restrict__ is_free, int fill,
                                                             int *__restrict__ out_len, const int *__restrict__ a_rowptr,
                                                             int *__restrict__ a_col, double *__restrict__ a_val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!is_free[i]) { if (!fill) out_len[i] = 0; return; }
    int cnt = 1;
    double diag = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        diag += (double)w[k];
        if (j != i && j < n_cols_valid && is_free[j]) ++cnt;
    }
    if (!fill) { out_len[i] = cnt; return; }
    const int base = a_rowptr[i];
    int q = 0;
    a_col[base] = i; a_val[base] = diag; q = 1;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        if (j == i || j >= n_cols_valid || !is_free[j]) continue;
        // insertion into the sorted prefix (rows have ~7 entries); a repeated column (cannot happen in a merged CSR) would add up
        int p = q;
        while (p > 0 && a_col[base + p - 1] > j) { a_col[base + p] = a_col[base + p - 1]; a_val[base + p] = a_val[base + p - 1]; --p; }
        a_col[base + p] = j; a_val[base + p] = -(double)w[k];
        ++q;
    }
}

__global__ void __launch_bounds__(kBlock) inv_diag_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                          const double *__restrict__ val, double *__restrict__ inv_diag, int *__restrict__ active) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) if (colidx[k] == i && val[k] > 0.0) d = 1.0 / val[k];
    inv_diag[i] = d;
    if (d > 0.0) atomicAdd(active, 1);
}

// ---- largest eigenvalue of D^-1 A: power iteration (Rayleigh quotient x'Ax / x'Dx) + Gershgorin bound -----------------
__global__ void __launch_bounds__(kBlock) rho_init_kernel(int n, const double *__restrict__ inv_diag, double *__restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = inv_diag[i] > 0 ? 1.0 + 0.37 * (double)(((unsigned)i * 2654435761u) % 97u) / 97.0 * ((i & 1) ? 1.0 : -1.0) : 0.0;
}
// y = D^-1 A x ; sums[0..2] = x'Ax, x'Dx, |y|^2 ; sums[3] = max_i sum_j |a_ij| / a_ii (as a sum-free max via atomics on bits)
__global__ void __launch_bounds__(kBlock) rho_step_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                          const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                          const double *__restrict__ x, double *__restrict__ y,
                                                          double *__restrict__ partials, unsigned *__restrict__ counter,
                                                          double *__restrict__ sums, unsigned long long *__restrict__ gersh_bits) {
    double red[3] = {0, 0, 0};
    double g = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double s = 0.0, a = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) { s += val[k] * x[colidx[k]]; a += fabs(val[k]); }
        const double d = inv_diag[i], xi = x[i];
        red[0] += xi * s;
        if (d > 0) red[1] += xi * xi / d;
        const double yi = s * d;
        y[i] = yi;
        red[2] += yi * yi;
        g = fmax(g, a * d);
    }
    if (gersh_bits) atomicMax(gersh_bits, (unsigned long long)__double_as_longlong(g));      // g >= 0: bit order == value order
    double total[3];
    if (grid_sum_last_block<3>(red, partials, counter, total)) { sums[0] = total[0]; sums[1] = total[1]; sums[2] = total[2]; }
}
__global__ void __launch_bounds__(kBlock) rho_scale_kernel(int n, const double *__restrict__ y, const double *__restrict__ sums, double *__restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double nrm = sqrt(sums[2]);
    if (nrm > 0) x[i] = y[i] / nrm;
}

// ---- aggregation ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_strong(int i, int j, double a, const double *__restrict__ inv_diag, const int *__restrict__ block, double theta2) {
    if (j == i) return false;
    const double dj = inv_diag[j];
    if (!(dj > 0)) return false;
    if (block && block[i] != block[j]) return false;                        // aggregates stay inside one partition block
    return a * a * inv_diag[i] * dj >= theta2;
}
// agg = -1 where the row takes no part in the coarse level (empty row or no strong neighbour), UNSET elsewhere
__global__ void __launch_bounds__(kBlock) agg_init_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                          const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                          const int *__restrict__ block, double theta2, int *__restrict__ agg,
                                                          int *__restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool has = false;
    if (inv_diag[i] > 0)
        for (int k = rowptr[i]; k < rowptr[i + 1] && !has; ++k) has = is_strong(i, colidx[k], val[k], inv_diag, block, theta2);
    agg[i] = has ? kAggUnset : -1;
    status[i] = has ? 0 : 2;              // 0 undecided, 1 root, 2 out of the election
}
// number of strong connections of the level (summed over the rows): its mean tells a quad-like strength graph (a plane of right
// triangles: the diagonals carry cot(90 deg) = 0) from a triangle-like one, which take different election keys (agg_key_kernel).
// With positions (3 doubles per row) also the summed length of those connections: the mean edge sizes the sweep cells.
__global__ void __launch_bounds__(kBlock) agg_strong_count_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                  const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                                  const int *__restrict__ block, double theta2, const double *__restrict__ pos,
                                                                  unsigned long long *__restrict__ total, double *__restrict__ length_total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    double len = 0.0;
    if (i < n && inv_diag[i] > 0)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colidx[k];
            if (!is_strong(i, j, val[k], inv_diag, block, theta2)) continue;
            ++cnt;
            if (pos) {
                const double dx = pos[3 * (size_t)i] - pos[3 * (size_t)j], dy = pos[3 * (size_t)i + 1] - pos[3 * (size_t)j + 1],
                             dz = pos[3 * (size_t)i + 2] - pos[3 * (size_t)j + 2];
                len += sqrt(dx * dx + dy * dy + dz * dz);
            }
        }
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, o); len += __shfl_down_sync(0xffffffffu, len, o); }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(total, (unsigned long long)cnt); if (pos) atomicAdd(length_total, len); }
}
// debugging aid (ARAP_MG_FAKE_BLOCKS=k): k artificial partition blocks of consecutive rows for an unpartitioned solver
__global__ void __launch_bounds__(kBlock) fake_block_kernel(int n, int k, int *__restrict__ block) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) block[i] = (int)((long long)i * k / n);
}
// positions of the finest level as 3 doubles per row, from the engine's own arrays (scalar type and stride vary)
__global__ void __launch_bounds__(kBlock) pos_gather_kernel(int n, const void *__restrict__ src, int scalar_bytes, int stride, double *__restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < 3; ++c)
        pos[3 * (size_t)i + c] = scalar_bytes == 4 ? (double)((const float *)src)[(size_t)i * stride + c] : ((const double *)src)[(size_t)i * stride + c];
}
// a coarse row sits where the root of its aggregate does
__global__ void __launch_bounds__(kBlock) pos_coarse_kernel(int n, const int *__restrict__ status, const int *__restrict__ root_id,
                                                            const double *__restrict__ pos, double *__restrict__ pos_c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] != 1) return;
    const size_t a = (size_t)root_id[i];
    pos_c[3 * a] = pos[3 * (size_t)i]; pos_c[3 * a + 1] = pos[3 * (size_t)i + 1]; pos_c[3 * a + 2] = pos[3 * (size_t)i + 2];
}
// sweep key of row i: rows of one CELL (a cube of 1 / inv_cell, about 8 mean edges: ~64 rows of a surface mesh) share a hashed
// priority, inside the cell the Z-curve position on a 16^3 sub-grid decides, then the lower index. Without positions the cells
// are runs of 64 consecutive rows (compact only if the numbering is).
__device__ __forceinline__ unsigned long long sweep_key(int i, const double *__restrict__ pos, double inv_cell) {
    if (!pos)
        return (1ULL << 63) | ((unsigned long long)(hash_u32((unsigned)(i >> 6)) & 0xffffffu) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    unsigned cell = 0x9e3779b9u, sub = 0u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double x = pos[3 * (size_t)i + c] * inv_cell, fx = floor(x);
        cell = hash_u32(cell ^ (unsigned)(long long)fx);
        unsigned q = (unsigned)((x - fx) * 16.0) & 15u;                  // 4 bits of this axis, spread to every third bit
        q = (q | (q << 4)) & 0x0c3u;
        q = (q | (q << 2)) & 0x249u;
        sub |= q << c;
    }
    return (1ULL << 63) | ((unsigned long long)(cell & 0x7ffffu) << 44) | ((unsigned long long)(0xfffu - sub) << 32) |
           (unsigned long long)(0xffffffffu - (unsigned)i);
}
// this round's keys, in one of two forms (0 = "not a candidate"):
//  * rim growth (sweep_shift == 0): candidates are the undecided vertices on the rim of what is already decided, plus a sparse set
//    of seeds (hash & seed_mask == 0) that start the growth, ordered by root_key(): aggregates pack tightly, ring after ring,
//    around few centres instead of around the ~n/13 random roots a first round open to every vertex elects;
//  * ordered sweeps (sweep != 0): every undecided vertex is a candidate, keyed by sweep_key(): compact cells of ~64 rows take
//    hashed priorities and inside a cell the order along a Z-curve decides, so every cell is swept in order the way the sequential
//    greedy pass of the host setup sweeps the whole level in Morton order, and the aggregates come out as a regular pattern.
// Measured, CG iterations per ARAP iteration (profiles/r02_experiments.txt): 1M icosphere 8.7 hashed keys / 8.7 host greedy /
// 7.4 rim growth / 9.8 sweeps; 2000 x 2000 plane (4-neighbour strength graph) 6.5 hashed / 5.35 host greedy / 6.1 rim / 5.4 sweeps.
__global__ void __launch_bounds__(kBlock) agg_key_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                         const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                         const int *__restrict__ block, double theta2, const int *__restrict__ status,
                                                         unsigned long long *__restrict__ key, unsigned seed_mask, int sweep,
                                                         const double *__restrict__ pos, double inv_cell) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k0 = 0ULL;
    if (status[i] == 0) {
        if (sweep) {
            k0 = sweep_key(i, pos, inv_cell);
        } else {
            int decided = 0;
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                const int j = colidx[k];
                if (is_strong(i, j, val[k], inv_diag, block, theta2) && status[j] != 0) ++decided;
            }
            if (decided > 0 || ((hash_u32((unsigned)i) >> 8) & seed_mask) == 0u) k0 = root_key(i, decided);
        }
    }
    key[i] = k0;
}
// m1[i] = largest key in the closed strong neighbourhood of i (0 if none is undecided)
__global__ void __launch_bounds__(kBlock) agg_max1_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                          const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                          const int *__restrict__ block, double theta2, const unsigned long long *__restrict__ key,
                                                          unsigned long long *__restrict__ m1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long m = key[i];
    if (inv_diag[i] > 0)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colidx[k];
            if (is_strong(i, j, val[k], inv_diag, block, theta2) && key[j] > m) m = key[j];
        }
    m1[i] = m;
}
// an undecided vertex whose key is the largest within distance 2 becomes a root
__global__ void __launch_bounds__(kBlock) agg_elect_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                           const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                           const int *__restrict__ block, double theta2, const unsigned long long *__restrict__ key,
                                                           const unsigned long long *__restrict__ m1, int *__restrict__ status, int *__restrict__ n_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] != 0) return;
    unsigned long long m = m1[i];
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        if (is_strong(i, j, val[k], inv_diag, block, theta2) && m1[j] > m) m = m1[j];
    }
    if (m == key[i] && m != 0ULL) { status[i] = 1; atomicAdd(n_new, 1); }
}
// after an election: neighbours of roots leave the election (status 3 = adjacent to a root), then their neighbours do (2)
__global__ void __launch_bounds__(kBlock) agg_cover1_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                            const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                            const int *__restrict__ block, double theta2, int *__restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] != 0) return;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        if (is_strong(i, j, val[k], inv_diag, block, theta2) && status[j] == 1) { status[i] = 3; return; }
    }
}
__global__ void __launch_bounds__(kBlock) agg_cover2_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                            const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                            const int *__restrict__ block, double theta2, int *__restrict__ status,
                                                            int *__restrict__ n_undecided) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] != 0) return;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        if (is_strong(i, j, val[k], inv_diag, block, theta2) && status[j] == 3) { status[i] = 2; return; }
    }
    atomicAdd(n_undecided, 1);
}
// ---- the lexicographically first maximal independent set of the squared strength graph, as a wavefront ---------------------------
// What the host setup's sequential greedy pass computes (visit the rows in order; a row none of whose strong neighbours is
// aggregated yet becomes a root and takes them): row i is a root iff no LOWER row within distance 2 is one. In parallel: i is decided
// as soon as every lower row within distance 2 is, so each round elects the undecided rows whose lower distance-2 neighbourhood is
// fully decided, covers their surroundings, and hands the undecided rows within distance 2 of anything decided this round to the
// next round as its worklist. One coherent sweep over the whole level -- no seams between independently grown patches, which is what
// the hashed / rim-growth / cell-sweep elections lose 20-40 % of the preconditioner's quality to on regular quad-like meshes --
// at the price of as many rounds as the longest dependency chain (~ the side length of the mesh): the rounds work on worklists
// (the front), with device-side counts and no host synchronisation except a termination check every 64 rounds.
// One relaxation (tests/test_aggregation_model.py models both forms): lex_cover1_kernel takes only UNDECIDED neighbours of a new
// root; a neighbour that an earlier root had marked "distance 2" keeps that mark, so the rows behind it are not covered by the new
// root and stay electable. The set is a few per cent denser than the walk's (roots at distance 2 through an already covered row),
// never has adjacent roots, is deterministic, and is what round 2 measured; ARAP_MG_WAVEFRONT_EXACT=1 re-marks (`exact`) and gives the
// walk's set (not yet measured on the hardware).
struct LexLists {
    int *work, *work_next, *roots, *adj, *far;      // this round's candidates, next round's, newly: roots / next to a root / distance 2
    int *count;                                       // [0] work [1] work_next [2] roots [3] adj [4] far
    int *stamp;                                       // round in which a row was last put on a worklist
};
__global__ void __launch_bounds__(kBlock) lex_fill_kernel(int n, const int *__restrict__ status, int *__restrict__ work, int *__restrict__ count,
                                                          int *__restrict__ stamp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    stamp[i] = -1;
    if (status[i] == 0) work[atomicAdd(&count[0], 1)] = i;
}
// One WARP per candidate: a single thread would walk ~50 dependent, uncached loads (rows of the neighbours of the neighbours) one
// after the other, ~100 us per round whatever the size of the front; the lanes take the neighbours' rows in parallel.
__global__ void __launch_bounds__(kBlock) lex_elect_kernel(LexLists L, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                           const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                           const int *__restrict__ block, double theta2, const int *__restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_work = L.count[0];
    for (int t = warp; t < n_work; t += n_warps) {
        const int i = L.work[t];
        if (status[i] != 0) continue;                                   // uniform over the warp
        const int b = rowptr[i], e = rowptr[i + 1];
        bool blocked = false;
        for (int base = b; base < e && !blocked; base += 32) {
            // every lane fetches one neighbour u of i and the extent of u's row
            const int k = base + lane;
            int u = -1, ub = 0, ue = 0;
            if (k < e) {
                u = colidx[k];
                if (is_strong(i, u, val[k], inv_diag, block, theta2)) { ub = rowptr[u]; ue = rowptr[u + 1]; } else u = -1;
            }
            bool bl = u >= 0 && u < i && status[u] == 0;
            const int cnt = min(32, e - base);
            for (int l = 0; l < cnt; ++l) {                              // then the warp walks u's row together
                const int uu = __shfl_sync(0xffffffffu, u, l);
                if (uu < 0) continue;
                const int qb = __shfl_sync(0xffffffffu, ub, l), qe = __shfl_sync(0xffffffffu, ue, l);
                for (int q = qb + lane; q < qe; q += 32) {
                    const int w = colidx[q];
                    if (w < i && status[w] == 0 && is_strong(uu, w, val[q], inv_diag, block, theta2)) bl = true;
                }
            }
            blocked = __any_sync(0xffffffffu, bl);
        }
        if (!blocked && lane == 0) L.roots[atomicAdd(&L.count[2], 1)] = i;
    }
}
// new roots take their undecided strong neighbours (two roots of one round are at distance >= 3: no neighbour is shared).
// One warp per root, one lane per row entry; the list append is one atomic per warp.
__device__ __forceinline__ void lex_append(bool take, int value, int *__restrict__ list, int *__restrict__ counter) {
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}
__global__ void __launch_bounds__(kBlock) lex_cover1_kernel(LexLists L, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                            const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                            const int *__restrict__ block, double theta2, int *__restrict__ status, int exact) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_roots = L.count[2];
    for (int t = warp; t < n_roots; t += n_warps) {
        const int r = L.roots[t];
        if (lane == 0) status[r] = 1;
        const int b = rowptr[r], e = rowptr[r + 1];
        for (int base = b; base < e; base += 32) {
            const int k = base + lane;
            bool take = false;
            int u = -1;
            if (k < e) {
                u = colidx[k];
                take = is_strong(r, u, val[k], inv_diag, block, theta2) && (status[u] == 0 || (exact && status[u] == 2));
                if (take) status[u] = 3;
            }
            lex_append(take, u, L.adj, &L.count[3]);
        }
    }
}
// ... and the undecided rows next to those leave the election (distance 2 from a root)
__global__ void __launch_bounds__(kBlock) lex_cover2_kernel(LexLists L, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                            const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                            const int *__restrict__ block, double theta2, int *__restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_adj = L.count[3];
    for (int t = warp; t < n_adj; t += n_warps) {
        const int u = L.adj[t];
        const int b = rowptr[u], e = rowptr[u + 1];
        for (int base = b; base < e; base += 32) {
            const int k = base + lane;
            bool take = false;
            int w = -1;
            if (k < e) {
                w = colidx[k];
                take = is_strong(u, w, val[k], inv_diag, block, theta2) && atomicCAS(&status[w], 0, 2) == 0;
            }
            lex_append(take, w, L.far, &L.count[4]);
        }
    }
}
// next round's worklist: the undecided rows within distance 2 of a row decided in this round (one warp per decided row)
__global__ void __launch_bounds__(kBlock) lex_next_kernel(LexLists L, int round, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                          const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                          const int *__restrict__ block, double theta2, const int *__restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_roots = L.count[2], n_adj = L.count[3], n_far = L.count[4];
    const int total = n_roots + n_adj + n_far;
    for (int t = warp; t < total; t += n_warps) {
        const int v = t < n_roots ? L.roots[t] : t < n_roots + n_adj ? L.adj[t - n_roots] : L.far[t - n_roots - n_adj];
        const int b = rowptr[v], e = rowptr[v + 1];
        for (int base = b; base < e; base += 32) {
            const int k = base + lane;
            int u = -1, ub = 0, ue = 0;
            if (k < e) {
                u = colidx[k];
                if (is_strong(v, u, val[k], inv_diag, block, theta2)) { ub = rowptr[u]; ue = rowptr[u + 1]; } else u = -1;
            }
            lex_append(u >= 0 && status[u] == 0 && atomicExch(&L.stamp[u], round) != round, u, L.work_next, &L.count[1]);
            const int cnt = min(32, e - base);
            for (int l = 0; l < cnt; ++l) {
                const int uu = __shfl_sync(0xffffffffu, u, l);
                if (uu < 0) continue;
                const int qb = __shfl_sync(0xffffffffu, ub, l), qe = __shfl_sync(0xffffffffu, ue, l);
                for (int q0 = qb; q0 < qe; q0 += 32) {
                    const int q = q0 + lane;
                    int w = -1;
                    bool take = false;
                    if (q < qe) {
                        w = colidx[q];
                        take = status[w] == 0 && is_strong(uu, w, val[q], inv_diag, block, theta2) && atomicExch(&L.stamp[w], round) != round;
                    }
                    lex_append(take, w, L.work_next, &L.count[1]);
                }
            }
        }
    }
}
__global__ void lex_advance_kernel(int *__restrict__ count, int *__restrict__ rounds_with_work) {
    count[0] = count[1];
    count[1] = count[2] = count[3] = count[4] = 0;
    if (count[0] > 0) *rounds_with_work += 1;
}
__global__ void __launch_bounds__(kBlock) agg_root_flag_kernel(int n, const int *__restrict__ status, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = status[i] == 1 ? 1 : 0;
}
// roots get their id (prefix of the root flags: coarse numbering follows the fine numbering); a vertex next to a root joins the
// root it is most strongly connected to (ties: the smaller root id)
__global__ void __launch_bounds__(kBlock) agg_assign1_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                             const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                             const int *__restrict__ block, double theta2, const int *__restrict__ status,
                                                             const int *__restrict__ root_id, int *__restrict__ agg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || agg[i] == -1) return;
    if (status[i] == 1) { agg[i] = root_id[i]; return; }
    double best = -1.0;
    int pick = kAggUnset;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = colidx[k];
        if (is_strong(i, j, val[k], inv_diag, block, theta2) && status[j] == 1) {
            const double a = fabs(val[k]);
            const int id = root_id[j];
            if (a > best || (a == best && id < pick)) { best = a; pick = id; }
        }
    }
    agg[i] = pick;
}
// leftovers (distance 2 from the nearest root) join the aggregate of pass 1 they are most strongly connected to
__global__ void __launch_bounds__(kBlock) agg_assign2_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                             const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                             const int *__restrict__ block, double theta2, const int *__restrict__ agg,
                                                             int *__restrict__ joined) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int pick = agg[i];
    if (pick == kAggUnset) {
        double best = -1.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colidx[k];
            if (is_strong(i, j, val[k], inv_diag, block, theta2) && agg[j] >= 0) {
                const double a = fabs(val[k]);
                if (a > best || (a == best && agg[j] < pick)) { best = a; pick = agg[j]; }
            }
        }
    }
    joined[i] = pick;
}
// coarse block (owner) = block of the aggregate's root
__global__ void __launch_bounds__(kBlock) agg_block_kernel(int n, const int *__restrict__ status, const int *__restrict__ root_id,
                                                           const int *__restrict__ block, int *__restrict__ block_c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && status[i] == 1) block_c[root_id[i]] = block[i];
}

// ---- sorted accumulator in local memory ------------------------------------------------------------------------------
template <int CAP>
struct RowAcc {
    int col[CAP];
    double val[CAP];
    int n = 0;
    bool overflow = false;
    __device__ __forceinline__ void add(int c, double v) {
        int lo = 0, hi = n;                      // first position with col >= c
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] < c) lo = mid + 1; else hi = mid; }
        if (lo < n && col[lo] == c) { val[lo] += v; return; }
        if (n == CAP) { overflow = true; return; }
        for (int p = n; p > lo; --p) { col[p] = col[p - 1]; val[p] = val[p - 1]; }
        col[lo] = c; val[lo] = v; ++n;
    }
};

// P = (I - omega D^-1 A) T, T piecewise constant over the aggregates. fill == 0: row lengths.
template <int CAP>
__global__ void __launch_bounds__(kBlock) prolongator_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                             const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                             const int *__restrict__ agg, double omega, int fill, int *__restrict__ out_len,
                                                             const int *__restrict__ p_rowptr, int *__restrict__ p_col, double *__restrict__ p_val,
                                                             int *__restrict__ error) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RowAcc<CAP> acc;
    const double d = inv_diag[i];
    if (d > 0) {
        if (agg[i] >= 0) acc.add(agg[i], 1.0);
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int J = agg[colidx[k]];
            if (J >= 0) acc.add(J, -omega * d * val[k]);
        }
    }
    if (acc.overflow) *error = 1;
    if (!fill) { out_len[i] = acc.n; return; }
    const int base = p_rowptr[i];
    for (int q = 0; q < acc.n; ++q) { p_col[base + q] = acc.col[q]; p_val[base + q] = acc.val[q]; }
}

// transpose: column counts, then a scatter with per-column cursors, then every row of the result sorted by column
__global__ void __launch_bounds__(kBlock) col_count_kernel(int nnz, const int *__restrict__ colidx, int *__restrict__ count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) atomicAdd(&count[colidx[k]], 1);
}
__global__ void __launch_bounds__(kBlock) transpose_fill_kernel(int n_rows, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                const double *__restrict__ val, const int *__restrict__ t_rowptr,
                                                                int *__restrict__ cursor, int *__restrict__ t_col, double *__restrict__ t_val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int c = colidx[k];
        const int q = t_rowptr[c] + atomicAdd(&cursor[c], 1);
        t_col[q] = i; t_val[q] = val[k];
    }
}
__global__ void __launch_bounds__(kBlock) sort_rows_kernel(int n_rows, const int *__restrict__ rowptr, int *__restrict__ colidx, double *__restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int lo = rowptr[i], hi = rowptr[i + 1];
    for (int a = lo + 1; a < hi; ++a) {
        const int cj = colidx[a];
        const double cv = val[a];
        int q = a - 1;
        while (q >= lo && colidx[q] > cj) { colidx[q + 1] = colidx[q]; val[q + 1] = val[q]; --q; }
        colidx[q + 1] = cj; val[q + 1] = cv;
    }
}

// C = A * B, one thread per row of A (rows of a few dozen products). fill == 0: row lengths.
template <int CAP>
__global__ void __launch_bounds__(kBlock) spgemm_rows_kernel(int n, const int *__restrict__ a_rowptr, const int *__restrict__ a_col,
                                                             const double *__restrict__ a_val, const int *__restrict__ b_rowptr,
                                                             const int *__restrict__ b_col, const double *__restrict__ b_val, int fill,
                                                             int *__restrict__ out_len, const int *__restrict__ c_rowptr, int *__restrict__ c_col,
                                                             double *__restrict__ c_val, int *__restrict__ error) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RowAcc<CAP> acc;
    for (int k = a_rowptr[i]; k < a_rowptr[i + 1]; ++k) {
        const int j = a_col[k];
        const double a = a_val[k];
        for (int q = b_rowptr[j]; q < b_rowptr[j + 1]; ++q) acc.add(b_col[q], a * b_val[q]);
    }
    if (acc.overflow) *error = 1;
    if (!fill) { out_len[i] = acc.n; return; }
    const int base = c_rowptr[i];
    for (int q = 0; q < acc.n; ++q) { c_col[base + q] = acc.col[q]; c_val[base + q] = acc.val[q]; }
}

// fp64 setup arrays -> the fp32 arrays the V-cycle reads
__global__ void __launch_bounds__(kBlock) to_float_kernel(size_t n, const double *__restrict__ in, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

}  // namespace mgdev
}  // namespace arap
