// halo_plan.h -- who sends which owned entries to whom (plain host struct shared by the CUDA engine and the host-side
// multigrid slicing). See partition.cuh for the transports that execute a plan.
#pragma once

#include <vector>

namespace arap {

struct HaloPlan {
    int n_owned = 0;
    std::vector<int> neighbor_rank;   // ranks this rank exchanges with
    std::vector<int> send_offset;     // [n_neighbors + 1] into send_index
    std::vector<int> send_index;      // owned local indices, grouped by neighbour, in the neighbour's halo order
    std::vector<int> recv_offset;     // [n_neighbors + 1]: halo from neighbour k sits at local n_owned + recv_offset[k] ...
    int n_send() const { return send_offset.empty() ? 0 : send_offset.back(); }
    int n_halo() const { return recv_offset.empty() ? 0 : recv_offset.back(); }
};

}  // namespace arap
