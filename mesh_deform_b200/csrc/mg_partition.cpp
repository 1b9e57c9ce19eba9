// mg_partition.cpp -- see mg_partition.h. Plain C++, no CUDA.
#include "mg_partition.h"

#include "arap_math.cuh"

#include <algorithm>
#include <cstdint>
#include <map>

namespace arap {

void build_global_csr(int V, int F, const int *faces, const double *xyz, std::vector<int> &rowptr, std::vector<int> &colidx,
                      std::vector<double> &weight) {
    // every face gives each of its corners two directed entries (arap.h:220-232)
    std::vector<int> raw_ptr((size_t)V + 1, 0);
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < 3; ++c) raw_ptr[(size_t)faces[3 * (size_t)f + c] + 1] += 2;
    for (int v = 0; v < V; ++v) raw_ptr[(size_t)v + 1] += raw_ptr[(size_t)v];
    std::vector<int> raw_col((size_t)raw_ptr[(size_t)V]);
    std::vector<double> raw_val((size_t)raw_ptr[(size_t)V]);
    std::vector<int> cursor(raw_ptr.begin(), raw_ptr.end() - 1);
    for (int f = 0; f < F; ++f) {
        const int v[3] = {faces[3 * (size_t)f], faces[3 * (size_t)f + 1], faces[3 * (size_t)f + 2]};
        double half[3];
        cotan_half_weights<double>(xyz + 3 * (size_t)v[0], xyz + 3 * (size_t)v[1], xyz + 3 * (size_t)v[2], half);
        for (int e = 0; e < 3; ++e) {                       // edge e = (v[e], v[e+1])
            const int a = v[e], b = v[(e + 1) % 3];
            raw_col[(size_t)cursor[(size_t)a]] = b; raw_val[(size_t)cursor[(size_t)a]++] = half[e];
            raw_col[(size_t)cursor[(size_t)b]] = a; raw_val[(size_t)cursor[(size_t)b]++] = half[e];
        }
    }
    rowptr.assign((size_t)V + 1, 0);
    colidx.clear();
    weight.clear();
    colidx.reserve(raw_col.size() / 2 + 16);
    weight.reserve(raw_col.size() / 2 + 16);
    std::vector<std::pair<int, double>> row;
    for (int v = 0; v < V; ++v) {
        row.clear();
        for (int k = raw_ptr[(size_t)v]; k < raw_ptr[(size_t)v + 1]; ++k) row.push_back({raw_col[(size_t)k], raw_val[(size_t)k]});
        std::stable_sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) { return x.first < y.first; });
        for (size_t k = 0; k < row.size(); ++k) {
            if (k > 0 && row[k].first == row[k - 1].first) weight.back() += row[k].second;
            else { colidx.push_back(row[k].first); weight.push_back(row[k].second); }
        }
        rowptr[(size_t)v + 1] = (int)colidx.size();
    }
}

void morton_sequence(int V, const double *xyz, std::vector<int> &order) {
    order.resize((size_t)V);
    for (int i = 0; i < V; ++i) order[(size_t)i] = i;
    if (V < 2) return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int v = 0; v < V; ++v)
        for (int d = 0; d < 3; ++d) { const double c = xyz[3 * (size_t)v + d]; if (c < lo[d]) lo[d] = c; if (c > hi[d]) hi[d] = c; }
    double extent = 0;
    for (int d = 0; d < 3; ++d) extent = std::max(extent, hi[d] - lo[d]);
    const double scale = extent > 0 ? 2097151.0 / extent : 0.0;                 // 21 bits per axis
    auto spread = [](uint64_t x) -> uint64_t {
        x &= 0x1fffffULL;
        x = (x | x << 32) & 0x1f00000000ffffULL;
        x = (x | x << 16) & 0x1f0000ff0000ffULL;
        x = (x | x << 8) & 0x100f00f00f00f00fULL;
        x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
        x = (x | x << 2) & 0x1249249249249249ULL;
        return x;
    };
    std::vector<std::pair<uint64_t, int>> keyed((size_t)V);
    for (int v = 0; v < V; ++v) {
        uint64_t key = 0;
        for (int d = 0; d < 3; ++d) key |= spread((uint64_t)((xyz[3 * (size_t)v + d] - lo[d]) * scale)) << d;
        keyed[(size_t)v] = {key, v};
    }
    std::sort(keyed.begin(), keyed.end());
    for (int i = 0; i < V; ++i) order[(size_t)i] = keyed[(size_t)i].second;
}

namespace {

// rows[r] = global row id (or -1 for an empty row); columns are renumbered through colmap (must be >= 0)
bool extract_rows(const HostCsr &M, const std::vector<int> &rows, const std::vector<int> &colmap, int n_cols, HostCsr &out) {
    out.n_rows = (int)rows.size();
    out.n_cols = n_cols;
    out.rowptr.assign(rows.size() + 1, 0);
    size_t total = 0;
    for (size_t r = 0; r < rows.size(); ++r) {
        if (rows[r] >= 0) total += (size_t)(M.rowptr[(size_t)rows[r] + 1] - M.rowptr[(size_t)rows[r]]);
        out.rowptr[r + 1] = (int)total;
    }
    out.colidx.resize(total);
    out.val.resize(total);
    size_t q = 0;
    for (size_t r = 0; r < rows.size(); ++r) {
        if (rows[r] < 0) continue;
        for (int k = M.rowptr[(size_t)rows[r]]; k < M.rowptr[(size_t)rows[r] + 1]; ++k, ++q) {
            const int c = colmap[(size_t)M.colidx[(size_t)k]];
            if (c < 0) return false;
            out.colidx[q] = c;
            out.val[q] = M.val[(size_t)k];
        }
    }
    return true;
}

}  // namespace

bool mg_slice_hierarchy(const MgHierarchyHost &H, int rank, int n_owned0, int n_local0, const int *global_of_local0,
                        MgLocalHierarchy &out, std::string &error, int replicate_rows) {
    const int L = (int)H.levels.size();
    if (L < 2) { error = "global multigrid: the hierarchy has a single level"; return false; }
    if (H.coarse_inv.empty() && !H.coarse_dense_on_device) { error = "global multigrid: no dense coarsest level"; return false; }
    for (int l = 0; l < L; ++l)
        if ((int)H.levels[(size_t)l].block.size() != H.levels[(size_t)l].A.n_rows) { error = "global multigrid: hierarchy was built without blocks"; return false; }
    out = MgLocalHierarchy();
    out.levels.resize((size_t)L);
    out.n_coarse = H.n_coarse;
    out.coarse_inv = H.coarse_inv;
    out.coarse_dense_on_device = H.coarse_dense_on_device;
    if (H.coarse_dense_on_device) out.coarse_A = H.levels.back().A;
    out.operator_complexity = H.operator_complexity;
    int Lr = L - 1;                                         // first replicated level (>= 1)
    while (Lr > 1 && H.levels[(size_t)Lr - 1].A.n_rows <= replicate_rows) --Lr;
    out.first_replicated = Lr;

    std::vector<std::vector<int>> g2l((size_t)L), own((size_t)L);
    // ---- level 0: the engine's numbering
    {
        const int n0 = H.levels[0].A.n_rows;
        g2l[0].assign((size_t)n0, -1);
        for (int i = 0; i < n_local0; ++i) {
            const int g = global_of_local0[i];
            if (g < 0 || g >= n0) { error = "global multigrid: local-to-global map out of range"; return false; }
            g2l[0][(size_t)g] = i;
        }
        own[0].assign(global_of_local0, global_of_local0 + n_owned0);
        for (int i = 0; i < n_owned0; ++i)
            if (H.levels[0].block[(size_t)own[0][(size_t)i]] != rank) { error = "global multigrid: owner array disagrees with the local mesh"; return false; }
        out.levels[0].n_own = n_owned0;
        out.levels[0].n_halo = n_local0 - n_owned0;
        out.levels[0].omega = H.levels[0].omega;
        out.levels[0].global_id.assign(global_of_local0, global_of_local0 + n_local0);
    }
    // ---- levels 1 .. L-1
    for (int l = 1; l < L; ++l) {
        const MgLevelHost &hl = H.levels[(size_t)l];
        const MgLevelHost &hf = H.levels[(size_t)l - 1];
        MgLocalLevel &ol = out.levels[(size_t)l];
        const int n = hl.A.n_rows;
        const std::vector<int> &block = hl.block;
        ol.omega = hl.omega;
        if (l >= Lr) {                                      // replicated: every row, global numbering
            g2l[(size_t)l].resize((size_t)n);
            for (int c = 0; c < n; ++c) g2l[(size_t)l][(size_t)c] = c;
            for (int c = 0; c < n; ++c) if (block[(size_t)c] == rank) own[(size_t)l].push_back(c);
            ol.n_own = n;
            ol.n_halo = 0;
            ol.inv_diag = hl.inv_diag;
            ol.global_id = g2l[(size_t)l];
            if (l < L - 1) ol.A = hl.A;
            continue;
        }
        for (int c = 0; c < n; ++c) if (block[(size_t)c] == rank) own[(size_t)l].push_back(c);
        std::vector<unsigned char> need((size_t)n, 0);
        for (int c : own[(size_t)l])
            for (int k = hl.A.rowptr[(size_t)c]; k < hl.A.rowptr[(size_t)c + 1]; ++k) need[(size_t)hl.A.colidx[(size_t)k]] = 1;
        for (int i : own[(size_t)l - 1])
            for (int k = hf.P.rowptr[(size_t)i]; k < hf.P.rowptr[(size_t)i + 1]; ++k) need[(size_t)hf.P.colidx[(size_t)k]] = 1;
        std::vector<int> halo;
        for (int c = 0; c < n; ++c) if (need[(size_t)c] && block[(size_t)c] != rank) halo.push_back(c);
        std::stable_sort(halo.begin(), halo.end(), [&](int a, int b) { return block[(size_t)a] < block[(size_t)b]; });
        g2l[(size_t)l].assign((size_t)n, -1);
        for (size_t k = 0; k < own[(size_t)l].size(); ++k) g2l[(size_t)l][(size_t)own[(size_t)l][k]] = (int)k;
        for (size_t k = 0; k < halo.size(); ++k) g2l[(size_t)l][(size_t)halo[k]] = (int)(own[(size_t)l].size() + k);
        ol.n_own = (int)own[(size_t)l].size();
        ol.n_halo = (int)halo.size();
        ol.global_id = own[(size_t)l];
        ol.global_id.insert(ol.global_id.end(), halo.begin(), halo.end());
        // what the other ranks need from me: columns I own in THEIR rows of A_l and of P_{l-1}
        std::map<int, std::vector<int>> send;               // neighbour rank -> my global ids
        for (int c = 0; c < n; ++c) {
            const int q = block[(size_t)c];
            if (q == rank) continue;
            for (int k = hl.A.rowptr[(size_t)c]; k < hl.A.rowptr[(size_t)c + 1]; ++k)
                if (block[(size_t)hl.A.colidx[(size_t)k]] == rank) send[q].push_back(hl.A.colidx[(size_t)k]);
        }
        for (int i = 0; i < hf.A.n_rows; ++i) {
            const int q = hf.block[(size_t)i];
            if (q == rank) continue;
            for (int k = hf.P.rowptr[(size_t)i]; k < hf.P.rowptr[(size_t)i + 1]; ++k)
                if (block[(size_t)hf.P.colidx[(size_t)k]] == rank) send[q].push_back(hf.P.colidx[(size_t)k]);
        }
        std::map<int, int> recv_count;
        for (int c : halo) recv_count[block[(size_t)c]] += 1;
        std::vector<int> nbrs;
        for (auto &kv : send) nbrs.push_back(kv.first);
        for (auto &kv : recv_count) nbrs.push_back(kv.first);
        std::sort(nbrs.begin(), nbrs.end());
        nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
        HaloPlan &pl = ol.plan;
        pl = HaloPlan();
        pl.n_owned = ol.n_own;
        pl.neighbor_rank = nbrs;
        pl.send_offset.assign(1, 0);
        pl.recv_offset.assign(1, 0);
        for (int q : nbrs) {
            std::vector<int> &s = send[q];
            std::sort(s.begin(), s.end());
            s.erase(std::unique(s.begin(), s.end()), s.end());
            for (int g : s) pl.send_index.push_back(g2l[(size_t)l][(size_t)g]);
            pl.send_offset.push_back((int)pl.send_index.size());
            pl.recv_offset.push_back(pl.recv_offset.back() + (recv_count.count(q) ? recv_count[q] : 0));
        }
        // A_l on the owned rows
        if (!extract_rows(hl.A, own[(size_t)l], g2l[(size_t)l], ol.n_own + ol.n_halo, ol.A)) { error = "global multigrid: operator column outside the halo"; return false; }
        ol.inv_diag.resize((size_t)ol.n_own);
        for (int k = 0; k < ol.n_own; ++k) ol.inv_diag[(size_t)k] = hl.inv_diag[(size_t)own[(size_t)l][(size_t)k]];
    }
    // ---- transfer operators
    for (int l = 0; l + 1 < L; ++l) {
        const MgLevelHost &hl = H.levels[(size_t)l];
        MgLocalLevel &ol = out.levels[(size_t)l];
        MgLocalLevel &oc = out.levels[(size_t)l + 1];
        if (l >= Lr) {                                      // between two replicated levels: the whole operators
            ol.P = hl.P;
            ol.R = hl.R;
            continue;
        }
        if (!extract_rows(hl.P, own[(size_t)l], g2l[(size_t)l + 1], oc.n_own + oc.n_halo, ol.P)) { error = "global multigrid: prolongation column outside the coarse halo"; return false; }
        std::vector<int> rrows;
        if (l + 1 == Lr) {                                  // into the replicated part: all rows, only mine filled (summed over the ranks)
            rrows.assign((size_t)oc.n_own, -1);
            for (int c : own[(size_t)l + 1]) rrows[(size_t)c] = c;
        } else {
            rrows = own[(size_t)l + 1];
        }
        if (!extract_rows(hl.R, rrows, g2l[(size_t)l], ol.n_own + ol.n_halo, ol.R)) { error = "global multigrid: restriction column outside the halo"; return false; }
    }
    return true;
}

}  // namespace arap
