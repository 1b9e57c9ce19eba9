// tests/cpp/mg_host_shim.cpp -- TEST INFRASTRUCTURE ONLY: exposes the engine's host-side multigrid
// setup (mesh_deform_b200/csrc/mg_setup.cpp) to Python so the hierarchy can be validated on a
// machine without a GPU (the V-cycle itself only exists as CUDA kernels).
#include "../../mesh_deform_b200/csrc/mg_setup.h"
#include "../../mesh_deform_b200/csrc/mg_partition.h"

#include <string>

static arap::MgHierarchyHost g_h;
static arap::MgLocalHierarchy g_local;
static std::string g_error;

extern "C" {
int mgshim_build(int V, const int *rowptr, const int *colidx, const double *w, const unsigned char *con, double theta, int coarse) {
    arap::MgSetupOptions o;
    if (theta > 0) o.theta = theta;
    if (coarse > 0) o.coarse_size = coarse;
    arap::mg_build_hierarchy<double>(V, rowptr, colidx, w, con, o, g_h);
    return (int)g_h.levels.size();
}
int mgshim_build_ordered(int V, const int *rowptr, const int *colidx, const double *w, const unsigned char *con, double theta, int coarse,
                         const int *visit_order) {
    arap::MgSetupOptions o;
    if (theta > 0) o.theta = theta;
    if (coarse > 0) o.coarse_size = coarse;
    arap::mg_build_hierarchy<double>(V, rowptr, colidx, w, con, o, g_h, visit_order);
    return (int)g_h.levels.size();
}
double mgshim_complexity() { return g_h.operator_complexity; }
int mgshim_ncoarse() { return g_h.n_coarse; }
int mgshim_has_inverse() { return g_h.coarse_inv.empty() ? 0 : 1; }
void mgshim_coarse_inverse(double *out) { for (size_t i = 0; i < g_h.coarse_inv.size(); ++i) out[i] = g_h.coarse_inv[i]; }
double mgshim_omega(int l) { return g_h.levels[l].omega; }
// which: 0 = A, 1 = P, 2 = R
static const arap::HostCsr &pick(int l, int which) { return which == 0 ? g_h.levels[l].A : which == 1 ? g_h.levels[l].P : g_h.levels[l].R; }
void mgshim_dims(int l, int which, int *rows, int *cols, int *nnz) { const auto &m = pick(l, which); *rows = m.n_rows; *cols = m.n_cols; *nnz = m.nnz(); }
void mgshim_get(int l, int which, int *rowptr, int *colidx, double *val) {
    const auto &m = pick(l, which);
    for (size_t i = 0; i < m.rowptr.size(); ++i) rowptr[i] = m.rowptr[i];
    for (size_t i = 0; i < m.colidx.size(); ++i) { colidx[i] = m.colidx[i]; val[i] = m.val[i]; }
}
void mgshim_inv_diag(int l, double *out) { for (size_t i = 0; i < g_h.levels[l].inv_diag.size(); ++i) out[i] = g_h.levels[l].inv_diag[i]; }

// ---- partitioned mode: global hierarchy with blocks, and one rank's share of it (mg_partition.h)
int mgshim_build_blocks(int V, const int *rowptr, const int *colidx, const double *w, const unsigned char *con, double theta, int coarse,
                        const int *visit_order, const int *block) {
    arap::MgSetupOptions o;
    if (theta > 0) o.theta = theta;
    if (coarse > 0) o.coarse_size = coarse;
    arap::mg_build_hierarchy<double>(V, rowptr, colidx, w, con, o, g_h, visit_order, block);
    return (int)g_h.levels.size();
}
void mgshim_block(int l, int *out) { for (size_t i = 0; i < g_h.levels[l].block.size(); ++i) out[i] = g_h.levels[l].block[i]; }
int mgshim_slice(int rank, int n_owned0, int n_local0, const int *global_of_local0) {
    return arap::mg_slice_hierarchy(g_h, rank, n_owned0, n_local0, global_of_local0, g_local, g_error) ? 1 : 0;
}
int mgshim_slice_replicated(int rank, int n_owned0, int n_local0, const int *global_of_local0, int replicate_rows) {
    return arap::mg_slice_hierarchy(g_h, rank, n_owned0, n_local0, global_of_local0, g_local, g_error, replicate_rows) ? 1 : 0;
}
int mgshim_first_replicated() { return g_local.first_replicated; }
const char *mgshim_error() { return g_error.c_str(); }
static const arap::HostCsr &pick_local(int l, int which) { return which == 0 ? g_local.levels[l].A : which == 1 ? g_local.levels[l].P : g_local.levels[l].R; }
void mgshim_local_dims(int l, int which, int *rows, int *cols, int *nnz) { const auto &m = pick_local(l, which); *rows = m.n_rows; *cols = m.n_cols; *nnz = m.nnz(); }
void mgshim_local_get(int l, int which, int *rowptr, int *colidx, double *val) {
    const auto &m = pick_local(l, which);
    for (size_t i = 0; i < m.rowptr.size(); ++i) rowptr[i] = m.rowptr[i];
    for (size_t i = 0; i < m.colidx.size(); ++i) { colidx[i] = m.colidx[i]; val[i] = m.val[i]; }
}
void mgshim_local_level(int l, int *n_own, int *n_halo, int *n_nbr, int *n_send, double *omega) {
    const auto &lv = g_local.levels[l];
    *n_own = lv.n_own; *n_halo = lv.n_halo; *n_nbr = (int)lv.plan.neighbor_rank.size(); *n_send = lv.plan.n_send(); *omega = lv.omega;
}
void mgshim_local_plan(int l, int *nbr, int *send_off, int *send_idx, int *recv_off) {
    const auto &pl = g_local.levels[l].plan;
    for (size_t i = 0; i < pl.neighbor_rank.size(); ++i) nbr[i] = pl.neighbor_rank[i];
    for (size_t i = 0; i < pl.send_offset.size(); ++i) send_off[i] = pl.send_offset[i];
    for (size_t i = 0; i < pl.send_index.size(); ++i) send_idx[i] = pl.send_index[i];
    for (size_t i = 0; i < pl.recv_offset.size(); ++i) recv_off[i] = pl.recv_offset[i];
}
void mgshim_local_global_id(int l, int *out) { for (size_t i = 0; i < g_local.levels[l].global_id.size(); ++i) out[i] = g_local.levels[l].global_id[i]; }
void mgshim_local_inv_diag(int l, double *out) { for (size_t i = 0; i < g_local.levels[l].inv_diag.size(); ++i) out[i] = g_local.levels[l].inv_diag[i]; }
void mgshim_csr(int V, int F, const int *faces, const double *xyz, int *rowptr, int *colidx, double *w, int *nnz) {
    std::vector<int> rp, ci;
    std::vector<double> ww;
    arap::build_global_csr(V, F, faces, xyz, rp, ci, ww);
    *nnz = (int)ci.size();
    if (!rowptr) return;
    for (size_t i = 0; i < rp.size(); ++i) rowptr[i] = rp[i];
    for (size_t i = 0; i < ci.size(); ++i) { colidx[i] = ci[i]; w[i] = ww[i]; }
}
void mgshim_morton(int V, const double *xyz, int *out) {
    std::vector<int> order;
    arap::morton_sequence(V, xyz, order);
    for (int i = 0; i < V; ++i) out[i] = order[(size_t)i];
}
}
