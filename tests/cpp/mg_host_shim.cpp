// tests/cpp/mg_host_shim.cpp -- TEST INFRASTRUCTURE ONLY: exposes the engine's host-side multigrid
// setup (mesh_deform_b200/csrc/mg_setup.cpp) to Python so the hierarchy can be validated on a
// machine without a GPU (the V-cycle itself only exists as CUDA kernels).
#include "../../mesh_deform_b200/csrc/mg_setup.h"

static arap::MgHierarchyHost g_h;

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
}
