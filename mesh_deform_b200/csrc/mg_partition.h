// mg_partition.h -- one multigrid hierarchy for a mesh that is partitioned over several GPUs.
//
// Block-Jacobi across ranks (every rank preconditioning its own block) costs O(sqrt(H/h)) CG iterations because nothing
// couples the blocks on the coarse levels. Here the hierarchy is GLOBAL: it is the same smoothed-aggregation hierarchy a
// single GPU would build for the whole mesh, except that aggregates never straddle two ranks (mg_setup.h, `block`), so
// every row of every level has one owner. Each rank keeps the rows it owns, numbered owned-first with the level's halo
// behind them, and the V-cycle exchanges halos level by level (engine.cu, vcycle_partitioned).
//
// Setup is replicated, not distributed: every rank builds the whole hierarchy from the global mesh on its host (the
// result is deterministic, so all ranks agree without talking) and then cuts out its share. That costs each rank the
// host time and memory of a single-GPU setup of the whole mesh; a distributed setup would have to exchange
// variable-length matrix rows between ranks and is left for later.
#pragma once

#include "halo_plan.h"
#include "mg_setup.h"

#include <string>
#include <vector>

namespace arap {

// Global one-ring CSR with cotan weights (reference arap.h:182-239) built on the host from the global mesh.
// rest_xyz: n_vertices x 3 doubles. Rows sorted by column, duplicates summed.
void build_global_csr(int n_vertices, int n_faces, const int *faces, const double *rest_xyz, std::vector<int> &rowptr,
                      std::vector<int> &colidx, std::vector<double> &weight);

// Morton (Z-curve) sequence of the vertices: the order in which the aggregation walks the fine level.
void morton_sequence(int n_vertices, const double *xyz, std::vector<int> &order);

struct MgLocalLevel {
    int n_own = 0, n_halo = 0;       // local numbering: owned rows (ascending global id), then the halo grouped by owner
    double omega = 2.0 / 3.0;
    HostCsr A;                       // owned rows, local columns (levels >= 1; level 0 is matrix-free in the engine)
    std::vector<double> inv_diag;    // owned rows
    HostCsr P;                       // owned rows -> next level's local numbering
    HostCsr R;                       // next level's owned rows -> this level's local numbering
    HaloPlan plan;                   // levels >= 1 (level 0 uses the engine's own plan)
    std::vector<int> global_id;      // local index -> row id in the global hierarchy (diagnostics and tests)
};

struct MgLocalHierarchy {
    std::vector<MgLocalLevel> levels;   // levels >= first_replicated (at least the coarsest) are replicated on every rank, global numbering
    int first_replicated = 0;
    int n_coarse = 0;
    std::vector<double> coarse_inv;     // dense inverse of the coarsest operator, or empty when ...
    bool coarse_dense_on_device = false;   // ... the engine is to invert coarse_A on the device (see MgHierarchyHost)
    HostCsr coarse_A;
    double operator_complexity = 0;
};

// Cut rank `rank`'s share out of a hierarchy built with blocks. global_of_local0: for every LOCAL fine index of the
// engine (owned first, then the one-ring halo) its global vertex id. Returns false (with a message) when the hierarchy
// cannot be used in partitioned form (no dense coarsest level, or a column outside the halo).
// replicate_rows: levels (other than the finest) with at most this many rows are kept WHOLE on every rank, in global
// numbering, like the coarsest level always is: the V-cycle then exchanges halos only on the large levels above them and
// sums one right-hand side (every rank fills the rows it owns) where it enters the replicated part. A partitioned level
// costs four halo exchanges per V-cycle -- pure latency on levels this small -- and replicated work is cheap there.
bool mg_slice_hierarchy(const MgHierarchyHost &H, int rank, int n_owned0, int n_local0, const int *global_of_local0,
                        MgLocalHierarchy &out, std::string &error, int replicate_rows = 0);

}  // namespace arap
