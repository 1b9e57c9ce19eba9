// mg_setup.cpp -- host-side smoothed-aggregation setup (see mg_setup.h). Plain C++, no CUDA.
#include "mg_setup.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

namespace arap {

namespace {

// Run f(begin, end) over contiguous blocks of [0, n) on a few host threads. The block boundaries depend only on n and the
// thread count, and every block's work is independent, so results do not depend on scheduling.
int host_threads(int n) {
    static const int cap = [] {
        const char *env = getenv("ARAP_MG_THREADS");
        if (env && atoi(env) > 0) return atoi(env);
        unsigned hw = std::thread::hardware_concurrency();
        return (int)std::min<unsigned>(hw ? hw : 1u, 16u);
    }();
    return std::max(1, std::min(cap, n / 8192));
}

template <class F>
void parallel_blocks(int n, F f) {
    const int threads = host_threads(n);
    if (threads <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const int b = (int)((long long)n * t / threads), e = (int)((long long)n * (t + 1) / threads);
        pool.emplace_back([=, &f] { f(b, e, t); });
    }
    for (auto &th : pool) th.join();
}
int parallel_block_count(int n) { return host_threads(n); }

// Level 0: L = D - W on the free vertices, in full vertex index space (reference arap.h:310-334:
// diagonal sums ALL neighbours, off-diagonals only free neighbours; constrained rows do not exist).
template <typename S>
void build_level0(int V, const int *rowptr, const int *colidx, const S *w, const unsigned char *con, HostCsr &A) {
    A.n_rows = A.n_cols = V;
    A.rowptr.assign((size_t)V + 1, 0);
    for (int i = 0; i < V; ++i) {
        int cnt = 0;
        if (!con[i]) {
            cnt = 1;
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) if (!con[colidx[k]] && colidx[k] != i) ++cnt;
        }
        A.rowptr[i + 1] = A.rowptr[i] + cnt;
    }
    A.colidx.resize((size_t)A.rowptr[V]);
    A.val.resize((size_t)A.rowptr[V]);
    for (int i = 0; i < V; ++i) {
        if (con[i]) continue;
        double diag = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) diag += (double)w[k];
        int q = A.rowptr[i];
        bool placed = false;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colidx[k];
            if (j == i) continue;                      // degenerate self edge: already part of diag, cancels in L
            if (!placed && j > i) { A.colidx[q] = i; A.val[q] = diag; ++q; placed = true; }
            if (!con[j]) { A.colidx[q] = j; A.val[q] = -(double)w[k]; ++q; }
        }
        if (!placed) { A.colidx[q] = i; A.val[q] = diag; ++q; }
    }
}

void extract_inv_diag(const HostCsr &A, std::vector<double> &inv_diag) {
    inv_diag.assign((size_t)A.n_rows, 0.0);
    for (int i = 0; i < A.n_rows; ++i)
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
            if (A.colidx[k] == i && A.val[k] > 0.0) inv_diag[i] = 1.0 / A.val[k];
}

// Largest eigenvalue of D^-1 A by power iteration (Rayleigh quotient x'Ax / x'Dx), with a safety margin.
double estimate_rho(const HostCsr &A, const std::vector<double> &inv_diag) {
    const int n = A.n_rows;
    std::vector<double> x((size_t)n), y((size_t)n);
    for (int i = 0; i < n; ++i) x[i] = inv_diag[i] > 0 ? 1.0 + 0.37 * ((i * 2654435761u) % 97) / 97.0 * ((i & 1) ? 1 : -1) : 0.0;
    double rho = 1.0;
    const int nb = parallel_block_count(n);
    std::vector<double> pxax((size_t)nb), pxdx((size_t)nb), pnrm((size_t)nb);
    for (int it = 0; it < 12; ++it) {
        parallel_blocks(n, [&](int b, int e, int t) {
            double xax = 0, xdx = 0, nrm = 0;
            for (int i = b; i < e; ++i) {
                double s = 0;
                for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) s += A.val[k] * x[A.colidx[k]];
                xax += x[i] * s;
                if (inv_diag[i] > 0) xdx += x[i] * x[i] / inv_diag[i];
                y[i] = s * inv_diag[i];
                nrm += y[i] * y[i];
            }
            pxax[(size_t)t] = xax; pxdx[(size_t)t] = xdx; pnrm[(size_t)t] = nrm;
        });
        double xax = 0, xdx = 0, nrm = 0;
        for (int t = 0; t < nb; ++t) { xax += pxax[(size_t)t]; xdx += pxdx[(size_t)t]; nrm += pnrm[(size_t)t]; }
        if (xdx > 0) rho = xax / xdx;
        nrm = std::sqrt(nrm);
        if (!(nrm > 0)) break;
        parallel_blocks(n, [&](int b, int e, int) { for (int i = b; i < e; ++i) x[i] = y[i] / nrm; });
    }
    double gersh = 0;                                  // Gershgorin bound on D^-1 A
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) s += std::fabs(A.val[k]);
        gersh = std::max(gersh, s * inv_diag[i]);
    }
    rho *= 1.1;
    if (gersh > 0) rho = std::min(rho, gersh);
    return std::max(rho, 1.0);
}

// Greedy aggregation on the strength graph. agg[i] = aggregate id, or -1 if the row takes no part
// in the coarse level (empty row, or no strong neighbour: Jacobi alone solves such rows).
int aggregate(const HostCsr &A, const std::vector<double> &inv_diag, double theta, std::vector<int> &agg, const int *order,
              const int *block) {
    const int n = A.n_rows;
    std::vector<unsigned char> strong((size_t)A.nnz(), 0);
    std::vector<unsigned char> has_strong((size_t)n, 0);
    for (int i = 0; i < n; ++i) {
        if (!(inv_diag[i] > 0)) continue;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
            const int j = A.colidx[k];
            if (j == i || !(inv_diag[j] > 0)) continue;
            if (block && block[i] != block[j]) continue;       // aggregates stay inside one partition block
            // |a_ij| >= theta sqrt(a_ii a_jj)  <=>  a_ij^2 * inv_ii * inv_jj >= theta^2
            if (A.val[k] * A.val[k] * inv_diag[i] * inv_diag[j] >= theta * theta) { strong[k] = 1; has_strong[i] = 1; }
        }
    }
    const int UNSET = -2;
    agg.assign((size_t)n, UNSET);
    for (int i = 0; i < n; ++i) if (!has_strong[i]) agg[i] = -1;
    int n_agg = 0;
    // pass 1: a vertex whose strong neighbourhood is untouched becomes a root
    for (int t = 0; t < n; ++t) {
        const int i = order ? order[t] : t;
        if (agg[i] != UNSET) continue;
        bool free_nbhd = true;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1] && free_nbhd; ++k)
            if (strong[k] && agg[A.colidx[k]] >= 0) free_nbhd = false;
        if (!free_nbhd) continue;
        agg[i] = n_agg;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
            if (strong[k] && agg[A.colidx[k]] == UNSET) agg[A.colidx[k]] = n_agg;
        ++n_agg;
    }
    // pass 2: leftovers join the pass-1 aggregate they are most strongly connected to
    std::vector<int> joined((size_t)n, UNSET);
    for (int i = 0; i < n; ++i) {
        if (agg[i] != UNSET) continue;
        double best = -1;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
            const int j = A.colidx[k];
            if (strong[k] && agg[j] >= 0 && std::fabs(A.val[k]) > best) { best = std::fabs(A.val[k]); joined[i] = agg[j]; }
        }
    }
    for (int i = 0; i < n; ++i) if (agg[i] == UNSET && joined[i] != UNSET) agg[i] = joined[i];
    // pass 3: whatever is still unset forms new aggregates with its unset strong neighbours
    for (int t = 0; t < n; ++t) {
        const int i = order ? order[t] : t;
        if (agg[i] != UNSET) continue;
        agg[i] = n_agg;
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
            if (strong[k] && agg[A.colidx[k]] == UNSET) agg[A.colidx[k]] = n_agg;
        ++n_agg;
    }
    return n_agg;
}

void sort_rows(HostCsr &M) {
    parallel_blocks(M.n_rows, [&](int b, int e, int) {
        std::vector<std::pair<int, double>> tmp;
        for (int i = b; i < e; ++i) {
            const int lo = M.rowptr[i], hi = M.rowptr[i + 1];
            if (hi - lo <= 24) {                               // short rows: in-place insertion sort
                for (int a = lo + 1; a < hi; ++a) {
                    const int cj = M.colidx[a];
                    const double cv = M.val[a];
                    int q = a - 1;
                    while (q >= lo && M.colidx[q] > cj) { M.colidx[q + 1] = M.colidx[q]; M.val[q + 1] = M.val[q]; --q; }
                    M.colidx[q + 1] = cj;
                    M.val[q + 1] = cv;
                }
                continue;
            }
            tmp.resize((size_t)(hi - lo));
            for (int k = lo; k < hi; ++k) tmp[(size_t)(k - lo)] = {M.colidx[k], M.val[k]};
            std::sort(tmp.begin(), tmp.end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b2) { return a.first < b2.first; });
            for (int k = lo; k < hi; ++k) { M.colidx[k] = tmp[(size_t)(k - lo)].first; M.val[k] = tmp[(size_t)(k - lo)].second; }
        }
    });
}

// Row-blocked assembly of a CSR matrix: every host thread produces the rows of its block into private vectors
// (emit(i, cols, vals) appends row i), then the blocks are concatenated in order.
template <class RowFn>
void assemble_rows(int n_rows, int n_cols, HostCsr &C, RowFn row_fn) {
    C.n_rows = n_rows;
    C.n_cols = n_cols;
    C.rowptr.assign((size_t)n_rows + 1, 0);
    // (Measured: running this assembly on 8 host threads is SLOWER than on one -- the threads fight over page faults of
    // their freshly grown output vectors -- so it stays sequential; only the read-only sweeps above use the thread pool.)
    const int nb = 1;
    std::vector<std::vector<int>> bc((size_t)nb);
    std::vector<std::vector<double>> bv((size_t)nb);
    auto one_block = [&](int b, int e, int t) {
        std::vector<int> slot((size_t)n_cols, -1);
        std::vector<int> &cols = bc[(size_t)t];
        std::vector<double> &vals = bv[(size_t)t];
        cols.reserve((size_t)(e - b) * 8);
        vals.reserve((size_t)(e - b) * 8);
        for (int i = b; i < e; ++i) {
            const int start = (int)cols.size();
            row_fn(i, start, cols, vals, slot);
            for (int q = start; q < (int)cols.size(); ++q) slot[cols[q]] = -1;
            C.rowptr[(size_t)i + 1] = (int)cols.size() - start;          // row length for now
        }
    };
    one_block(0, n_rows, 0);
    for (int i = 0; i < n_rows; ++i) C.rowptr[(size_t)i + 1] += C.rowptr[(size_t)i];
    C.colidx.swap(bc[0]);
    C.val.swap(bv[0]);
    sort_rows(C);
}

// P = (I - omega D^-1 A) T, T piecewise constant over the aggregates.
void smoothed_prolongator(const HostCsr &A, const std::vector<double> &inv_diag, const std::vector<int> &agg, int n_agg,
                          double omega, HostCsr &P) {
    assemble_rows(A.n_rows, n_agg, P, [&](int i, int start, std::vector<int> &cols, std::vector<double> &vals, std::vector<int> &slot) {
        if (!(inv_diag[i] > 0)) return;
        if (agg[i] >= 0) { slot[agg[i]] = (int)cols.size(); cols.push_back(agg[i]); vals.push_back(1.0); }
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
            const int J = agg[A.colidx[k]];
            if (J < 0) continue;
            const double v = -omega * inv_diag[i] * A.val[k];
            if (slot[J] >= start) vals[(size_t)slot[J]] += v;
            else { slot[J] = (int)cols.size(); cols.push_back(J); vals.push_back(v); }
        }
    });
}

void transpose(const HostCsr &M, HostCsr &T) {
    T.n_rows = M.n_cols;
    T.n_cols = M.n_rows;
    T.rowptr.assign((size_t)T.n_rows + 1, 0);
    for (int c : M.colidx) T.rowptr[(size_t)c + 1]++;
    for (int i = 0; i < T.n_rows; ++i) T.rowptr[i + 1] += T.rowptr[i];
    T.colidx.resize(M.colidx.size());
    T.val.resize(M.val.size());
    std::vector<int> cur(T.rowptr.begin(), T.rowptr.end() - 1);
    for (int i = 0; i < M.n_rows; ++i)
        for (int k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) {
            const int q = cur[M.colidx[k]]++;
            T.colidx[q] = i;
            T.val[q] = M.val[k];
        }
}

// C = A * B (Gustavson, sparse accumulator), rows sorted on exit.
void spgemm(const HostCsr &A, const HostCsr &B, HostCsr &C) {
    assemble_rows(A.n_rows, B.n_cols, C, [&](int i, int start, std::vector<int> &cols, std::vector<double> &vals, std::vector<int> &slot) {
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
            const int j = A.colidx[k];
            const double a = A.val[k];
            for (int q = B.rowptr[j]; q < B.rowptr[j + 1]; ++q) {
                const int cc = B.colidx[q];
                if (slot[cc] >= start) vals[(size_t)slot[cc]] += a * B.val[q];
                else { slot[cc] = (int)cols.size(); cols.push_back(cc); vals.push_back(a * B.val[q]); }
            }
        }
    });
}

// Dense inverse of the (SPD up to a null space) coarsest operator by Gauss-Jordan with partial pivoting.
bool dense_inverse(const HostCsr &A, std::vector<double> &inv) {
    const int n = A.n_rows;
    std::vector<double> M((size_t)n * n, 0.0);
    double trace = 0;
    for (int i = 0; i < n; ++i)
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
            M[(size_t)i * n + A.colidx[k]] += A.val[k];
            if (A.colidx[k] == i) trace += A.val[k];
        }
    const double shift = n > 0 ? 1e-13 * trace / n : 0.0;   // keeps a pure-Neumann component invertible
    for (int i = 0; i < n; ++i) {
        if (M[(size_t)i * n + i] == 0.0) M[(size_t)i * n + i] = 1.0;   // empty row: identity
        else M[(size_t)i * n + i] += shift;
    }
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(M[(size_t)r * n + c]) > std::fabs(M[(size_t)piv * n + c])) piv = r;
        if (M[(size_t)piv * n + c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < n; ++k) {
                std::swap(M[(size_t)c * n + k], M[(size_t)piv * n + k]);
                std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]);
            }
        const double d = 1.0 / M[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) { M[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = M[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { M[(size_t)r * n + k] -= f * M[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
        }
    }
    return true;
}

}  // namespace

struct PhaseTimer {
    bool on = getenv("ARAP_MG_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char *what, int level) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mg setup] level %d %-12s %7.1f ms\n", level, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

template <typename S>
void mg_build_hierarchy(int V, const int *rowptr, const int *colidx, const S *weight, const unsigned char *con,
                        const MgSetupOptions &opt, MgHierarchyHost &out, const int *visit_order, const int *block) {
    PhaseTimer timer;
    std::vector<int> cur_block;
    if (block) cur_block.assign(block, block + V);
    out.levels.clear();
    out.coarse_inv.clear();
    HostCsr A;
    build_level0<S>(V, rowptr, colidx, weight, con, A);
    timer.lap("level0", 0);
    const double fine_nnz = std::max(1, A.nnz());
    double total_nnz = 0;
    while (true) {
        MgLevelHost lvl;
        lvl.A = std::move(A);
        lvl.block = cur_block;
        extract_inv_diag(lvl.A, lvl.inv_diag);
        total_nnz += lvl.A.nnz();
        int active = 0;
        for (double d : lvl.inv_diag) if (d > 0) ++active;
        const bool last = active <= opt.coarse_size || (int)out.levels.size() + 1 >= opt.max_levels;
        if (!last) {
            const int lv = (int)out.levels.size();
            lvl.omega = 4.0 / (3.0 * estimate_rho(lvl.A, lvl.inv_diag));
            timer.lap("rho", lv);
            std::vector<int> agg;
            const int n_agg = aggregate(lvl.A, lvl.inv_diag, opt.theta, agg, out.levels.empty() ? visit_order : nullptr,
                                        block ? lvl.block.data() : nullptr);
            timer.lap("aggregate", lv);
            if (n_agg > 0 && n_agg < 0.8 * active) {
                smoothed_prolongator(lvl.A, lvl.inv_diag, agg, n_agg, lvl.omega, lvl.P);
                timer.lap("prolongator", lv);
                transpose(lvl.P, lvl.R);
                timer.lap("transpose", lv);
                HostCsr AP;
                spgemm(lvl.A, lvl.P, AP);
                timer.lap("A*P", lv);
                spgemm(lvl.R, AP, A);
                timer.lap("R*(AP)", lv);
                if (block) {
                    cur_block.assign((size_t)n_agg, 0);
                    for (int i = 0; i < lvl.A.n_rows; ++i) if (agg[(size_t)i] >= 0) cur_block[(size_t)agg[(size_t)i]] = lvl.block[(size_t)i];
                }
                out.levels.push_back(std::move(lvl));
                continue;
            }
        }
        // coarsest level: keep A in a final pseudo-level without P
        lvl.omega = 4.0 / (3.0 * estimate_rho(lvl.A, lvl.inv_diag));
        out.n_coarse = lvl.A.n_rows;
        out.coarse_dense_on_device = false;
        if (lvl.A.n_rows <= opt.host_dense_max && lvl.A.n_rows <= opt.max_dense) {
            if (!dense_inverse(lvl.A, out.coarse_inv)) out.coarse_inv.clear();
        } else if (lvl.A.n_rows <= opt.max_dense) {
            out.coarse_dense_on_device = true;
        }
        timer.lap("dense inverse", (int)out.levels.size());
        out.levels.push_back(std::move(lvl));
        break;
    }
    out.operator_complexity = total_nnz / fine_nnz;
}

template void mg_build_hierarchy<float>(int, const int *, const int *, const float *, const unsigned char *,
                                        const MgSetupOptions &, MgHierarchyHost &, const int *, const int *);
template void mg_build_hierarchy<double>(int, const int *, const int *, const double *, const unsigned char *,
                                         const MgSetupOptions &, MgHierarchyHost &, const int *, const int *);

}  // namespace arap
