// mg_kernels.cuh -- device side of the multigrid-preconditioned CG (global step, replaces the
// reference's SimplicialLDLT::solve, arap.h:418-421). All vectors are Vec3d (three right-hand sides
// at once, fp64). The fine level is matrix-free on the one-ring CSR; coarse levels are explicit CSR.
//
// One V(1,1)-cycle, damped Jacobi smoothing, applied to b_0 = r (the CG residual):
//   x_l = omega_l D_l^-1 b_l ; r_l = b_l - A_l x_l ; b_{l+1} = R_l r_l ; ... ; x_L = A_L^-1 b_L (dense)
//   x_l += P_l x_{l+1} ; x_l += omega_l D_l^-1 (b_l - A_l x_l)
// HBM-bound row-gather kernels, one thread per row.
#pragma once

#include "kernels.cuh"

namespace arap {

// ---- fine level (matrix-free): (A x)_i = sum_j w_ij (x_i - x_j) on free rows ---------------------------
// Neighbours are processed in chunks of 6 (the typical valence): all index/weight loads, then all
// gathers of the chunk are issued before the arithmetic, so ~20 loads are in flight per thread.
template <typename S>
__device__ __forceinline__ Vec3d fine_apply_row(int i, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                const S *__restrict__ weight, const Vec3d *__restrict__ x) {
    constexpr int CH = kSpmvChunk;
    const int k0 = rowptr[i], k1 = rowptr[i + 1];
    const Vec3d xi = x[i];
    Vec3d out = {0, 0, 0};
    for (int k = k0; k < k1; k += CH) {
        int j[CH];
        double w[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const bool valid = k + u < k1;
            j[u] = valid ? __ldg(&colidx[k + u]) : i;
            w[u] = valid ? (double)__ldg(&weight[k + u]) : 0.0;
        }
        Vec3d xj[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) xj[u] = x[j[u]];
        const double gate = gather_gate(xj);
#pragma unroll
        for (int u = 0; u < CH; ++u) w[u] += gate;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            out.x += w[u] * (xi.x - xj[u].x); out.y += w[u] * (xi.y - xj[u].y); out.z += w[u] * (xi.z - xj[u].z);
        }
    }
    return out;
}

// r0 = b - A x0 on free rows (0 elsewhere)
template <typename S>
__global__ void __launch_bounds__(kBlock) mg_fine_residual_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                  const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                                  const Vec3d *__restrict__ b, const Vec3d *__restrict__ x,
                                                                  Vec3d *__restrict__ r, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vec3d out = {0, 0, 0};
    if (rest4[i].w != S(0)) {
        const Vec3d ax = fine_apply_row<S>(i, rowptr, colidx, weight, x);
        const Vec3d bi = b[i];
        out.x = bi.x - ax.x; out.y = bi.y - ax.y; out.z = bi.z - ax.z;
    }
    r[i] = out;
}

// z = x + omega D^-1 (b - A x) on free rows; fused rho_new = b . z (b is the CG residual) -> beta.
template <typename S>
__global__ void __launch_bounds__(kBlock) mg_fine_postsmooth_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                    const S *__restrict__ weight, const Vec4T<S> *__restrict__ rest4,
                                                                    const double *__restrict__ inv_diag, double omega,
                                                                    const Vec3d *__restrict__ b, const Vec3d *__restrict__ x,
                                                                    Vec3d *__restrict__ z, double *__restrict__ partials,
                                                                    unsigned *__restrict__ counter, CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    double red[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Vec3d out = {0, 0, 0};
        if (rest4[i].w != S(0)) {
            const Vec3d ax = fine_apply_row<S>(i, rowptr, colidx, weight, x);
            const Vec3d bi = b[i], xi = x[i];
            const double s = omega * inv_diag[i];
            out.x = xi.x + s * (bi.x - ax.x); out.y = xi.y + s * (bi.y - ax.y); out.z = xi.z + s * (bi.z - ax.z);
            red[0] += bi.x * out.x; red[1] += bi.y * out.y; red[2] += bi.z * out.z;
        }
        z[i] = out;
    }
    double total[3];
    if (grid_sum_last_block<3>(red, partials, counter, total)) cg_finish_reduction<3>(cg, CG_STAGE_RHO, total);
}

// ---- generic CSR levels -------------------------------------------------------------------------------
// LANES (a power of two <= 32) consecutive threads share one row and reduce with shuffles, so that
// rows of ~10-25 entries (coarse operators, restriction) still spread over enough threads to fill the GPU.
template <int LANES>
__device__ __forceinline__ Vec3d csr_apply_row(int row, bool row_valid, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                               const double *__restrict__ val, const Vec3d *__restrict__ x) {
    Vec3d out = {0, 0, 0};
    const int sub = threadIdx.x & (LANES - 1);
    if (row_valid) {
        const int k0 = rowptr[row], k1 = rowptr[row + 1];
        for (int k = k0 + sub; k < k1; k += LANES) {
            const int j = __ldg(&colidx[k]);
            const double a = __ldg(&val[k]);
            const Vec3d xj = x[j];
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
                                                                 const double *__restrict__ val, const Vec3d *__restrict__ b,
                                                                 const Vec3d *__restrict__ x, Vec3d *__restrict__ r,
                                                                 const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const Vec3d ax = csr_apply_row<LANES>(i, i < n, rowptr, colidx, val, x);
    if (i < n && (threadIdx.x & (LANES - 1)) == 0) {
        const Vec3d bi = b[i];
        r[i] = Vec3d{bi.x - ax.x, bi.y - ax.y, bi.z - ax.z};
    }
}

// b_c = R r_f ; x_c = omega_c D_c^-1 b_c   (restriction fused with the coarse level's pre-smoothing from a zero guess)
template <int LANES>
__global__ void __launch_bounds__(kBlock) mg_restrict_presmooth_kernel(int nc, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                       const double *__restrict__ val, const Vec3d *__restrict__ r_fine,
                                                                       const double *__restrict__ inv_diag_c, double omega_c,
                                                                       Vec3d *__restrict__ b_c, Vec3d *__restrict__ x_c,
                                                                       const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const Vec3d bc = csr_apply_row<LANES>(i, i < nc, rowptr, colidx, val, r_fine);
    if (i < nc && (threadIdx.x & (LANES - 1)) == 0) {
        b_c[i] = bc;
        const double s = omega_c * inv_diag_c[i];
        x_c[i] = Vec3d{s * bc.x, s * bc.y, s * bc.z};
    }
}

// x += P x_c   (P rows hold ~1-4 entries: one thread per row)
__global__ void __launch_bounds__(kBlock) mg_prolong_add_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                const double *__restrict__ val, const Vec3d *__restrict__ x_c,
                                                                Vec3d *__restrict__ x, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (rowptr[i] == rowptr[i + 1]) return;
    const Vec3d c = csr_apply_row<1>(i, true, rowptr, colidx, val, x_c);
    Vec3d xi = x[i];
    xi.x += c.x; xi.y += c.y; xi.z += c.z;
    x[i] = xi;
}

// x_out = x + omega D^-1 (b - A x)
template <int LANES>
__global__ void __launch_bounds__(kBlock) mg_csr_postsmooth_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                   const double *__restrict__ val, const double *__restrict__ inv_diag,
                                                                   double omega, const Vec3d *__restrict__ b, const Vec3d *__restrict__ x,
                                                                   Vec3d *__restrict__ x_out, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const Vec3d ax = csr_apply_row<LANES>(i, i < n, rowptr, colidx, val, x);
    if (i < n && (threadIdx.x & (LANES - 1)) == 0) {
        const Vec3d bi = b[i], xi = x[i];
        const double s = omega * inv_diag[i];
        x_out[i] = Vec3d{xi.x + s * (bi.x - ax.x), xi.y + s * (bi.y - ax.y), xi.z + s * (bi.z - ax.z)};
    }
}

// coarsest level: x = A^-1 b with the dense inverse; one warp per row.
__global__ void __launch_bounds__(kBlock) mg_dense_solve_kernel(int n, const double *__restrict__ inv, const Vec3d *__restrict__ b,
                                                                Vec3d *__restrict__ x, const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int c = lane; c < n; c += 32) {
        const double a = inv[(size_t)row * n + c];
        const Vec3d bc = b[c];
        s0 += a * bc.x; s1 += a * bc.y; s2 += a * bc.z;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) x[row] = Vec3d{s0, s1, s2};
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
                                                              double *__restrict__ x, double *__restrict__ r, double *__restrict__ x0,
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
            const double s0 = omega0 * inv_diag[e / 3], s1 = omega0 * inv_diag[(e + 1) / 3];
            *reinterpret_cast<double2 *>(x0 + e) = make_double2(s0 * rv.x, s1 * rv.y);
            red[0] += rv.x * rv.x + rv.y * rv.y;
        } else {
            const double a0 = pick3(e % 3, al0, al1, al2);
            const double xv = x[e] + a0 * d[e], rv = r[e] - a0 * ad[e];
            x[e] = xv; r[e] = rv;
            x0[e] = omega0 * inv_diag[e / 3] * rv;
            red[0] += rv * rv;
        }
    }
    double total[1];
    if (grid_sum_last_block<1>(red, partials, counter, total)) cg_finish_reduction<1>(cg, CG_STAGE_UPDATE_MG, total);
}

// d = z + beta d
__global__ void __launch_bounds__(kBlock) cg_direction_mg_kernel(int n3, const double *__restrict__ z, double *__restrict__ d,
                                                                 const CgScalars *__restrict__ cg) {
    if (cg->converged) return;
    const int e = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (e >= n3) return;
    const double b0 = cg->beta[0], b1 = cg->beta[1], b2 = cg->beta[2];
    if (e + 1 < n3) {
        const double2 zv = *reinterpret_cast<const double2 *>(z + e);
        double2 dv = *reinterpret_cast<const double2 *>(d + e);
        dv.x = zv.x + pick3(e % 3, b0, b1, b2) * dv.x;
        dv.y = zv.y + pick3((e + 1) % 3, b0, b1, b2) * dv.y;
        *reinterpret_cast<double2 *>(d + e) = dv;
    } else {
        d[e] = z[e] + pick3(e % 3, b0, b1, b2) * d[e];
    }
}

}  // namespace arap
