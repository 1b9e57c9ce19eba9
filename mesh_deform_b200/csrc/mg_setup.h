// mg_setup.h -- host-side setup of the smoothed-aggregation multigrid hierarchy that preconditions
// the global step's CG (the role SimplicialLDLT::compute plays in the reference, arap.h:336-337:
// a one-off analysis of L per constraint change; the per-iteration work is all on the device).
//
// Input: the one-ring CSR (_edgeWeights, arap.h:453) and the constrained-vertex mask. Level 0 lives
// in full vertex index space (constrained rows/columns simply do not exist in L), so no free-index
// compaction is needed anywhere. Output per level: A (CSR, fp64), 1/diag, Jacobi weight, P and R = P^T.
//
// Algorithm (Vanek, Mandel, Brezina 1996): strength graph |a_ij| >= theta sqrt(a_ii a_jj); greedy
// root + neighbours aggregation; tentative piecewise-constant prolongator T; P = (I - omega D^-1 A) T;
// Galerkin A_c = P^T A P; dense inverse on the coarsest level.
#pragma once

#include <cstdint>
#include <vector>

namespace arap {

struct HostCsr {
    int n_rows = 0, n_cols = 0;
    std::vector<int> rowptr, colidx;
    std::vector<double> val;
    int nnz() const { return (int)colidx.size(); }
};

struct MgLevelHost {
    HostCsr A;                       // level operator (level 0: L in full vertex index space, constrained rows empty)
    std::vector<double> inv_diag;    // 0 where the row is empty
    double omega = 2.0 / 3.0;        // damped-Jacobi weight
    HostCsr P;                       // n_l x n_{l+1}
    HostCsr R;                       // P^T
    std::vector<int> block;          // per row: the partition block (owner rank) it belongs to; empty without blocks
};

struct MgHierarchyHost {
    std::vector<MgLevelHost> levels; // levels[l] for l = 0 .. L-1 (each has a P to the next level)
    int n_coarse = 0;                // size of the coarsest level
    std::vector<double> coarse_inv;  // dense n_coarse x n_coarse inverse (row-major); empty if not computed here
    bool coarse_dense_on_device = false;   // the coarsest level is small enough for a dense inverse, but too large to invert on
                                           // the host in reasonable time: the engine inverts levels.back().A on the device
    double operator_complexity = 0;
};

struct MgSetupOptions {
    double theta = 0.08;
    int coarse_size = 256;
    int max_levels = 12;
    int max_dense = 2048;
    int host_dense_max = 2048;       // largest coarsest level inverted on the host (O(n^3) scalar code: 256 rows = 10 ms, 1234 = seconds)
};

// w may be float or double (the handle precision); it is widened to double.
// visit_order (optional, n_vertices entries): the order in which the greedy aggregation of the FINE level walks the
// vertices. A spatially coherent order (Morton) gives compact, regular aggregates -- fewer CG iterations -- whatever
// the memory order of the vertices is; coarse levels inherit it through the aggregate numbering.
// block (optional, n_vertices entries): a partition of the vertices (owner rank per vertex, partitioned mode). Aggregates
// never contain vertices of two blocks, on any level, so every coarse row has a well-defined owner (levels[l].block)
// and the restriction of an owned coarse row only reads fine rows within one ring of the owner's rows.
template <typename S>
void mg_build_hierarchy(int n_vertices, const int *rowptr, const int *colidx, const S *weight,
                        const unsigned char *is_constrained, const MgSetupOptions &opt, MgHierarchyHost &out,
                        const int *visit_order = nullptr, const int *block = nullptr);

}  // namespace arap
