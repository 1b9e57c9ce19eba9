// peer_transport.cuh -- halo exchange and small all-reduces by direct stores into the peers' memory (NVLink / NVSwitch
// peer access between GPUs of one node; plain device pointers between partitions that share a GPU).
//
// Why: a CG iteration of the partitioned solver makes ~27 tiny exchanges and reductions (one per gathered vector and
// multigrid level). As NCCL operations each costs ~15 us on the device, which is most of the 2-GPU iteration time. Here
// an exchange is ONE kernel on the sending side and ONE on the receiving side, with no library call in between:
//
//   put kernel    gathers the owned entries the neighbours need and stores them straight into each neighbour's MAILBOX
//                 (a region of the neighbour's arena reserved for this call site and this sender); the last CTA then
//                 publishes the site's epoch number in the neighbour's flag word (st.release.sys after a system fence).
//   wait kernel   (one CTA) spins on the flag words of all neighbours (ld.acquire.sys) until they show this epoch, then
//                 copies the mailbox -- laid out in the halo's own order -- into the halo slots of the array.
//
// An all-reduce is the same thing with every rank as neighbour: each rank stores its n partial sums into its slot of
// every rank's mailbox, the wait kernel adds the slots up in rank order (so every rank gets identical bits).
// Everything is stream-ordered kernel launches with static arguments (epochs are counted on the device), so a whole CG
// iteration, exchanges included, sits in one CUDA graph.
//
// Why a mailbox may be overwritten by the next epoch without an acknowledgement: every call site has its own mailbox,
// and between two uses of the same site each rank passes at least one other site where it WAITS for the same neighbour's
// later put -- and that neighbour issued this later put after (in stream order) it had unpacked the earlier mailbox.
// (Every site is symmetric: whoever receives from a rank also sends to it, if only the flag.)
//
// Setup (configure): every rank lays out its arena = [flag words][reduce mailboxes][exchange mailboxes], tells the others
// where each sender's segment starts (all-gather of a small table through the bootstrap transport: NCCL between
// processes, shared host memory between threads) and maps the peers' arenas (CUDA IPC between processes).
#pragma once

#include "partition.cuh"

#include <unistd.h>

namespace arap {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxSites = 64;
constexpr unsigned long long kPeerTimeoutNs = 30ULL * 1000000000ULL;   // a lost peer must not hang the GPU for ever

struct PeerExchangeDev {
    int n_nbr, words;                                   // 8-byte words per element
    int send_offset[kPeerMaxRanks + 1];
    unsigned long long *remote_payload[kPeerMaxRanks];  // where my segment starts in neighbour k's mailbox
    unsigned long long *remote_flag[kPeerMaxRanks];     // my flag word in neighbour k's arena
    const unsigned long long *local_payload;            // this site's mailbox in my arena (halo order)
    const unsigned long long *local_flag[kPeerMaxRanks];
    int n_recv_words;
    unsigned long long send_epoch;
    unsigned int done;
    unsigned int left;                                  // one-kernel exchange: CTAs that have read send_epoch_base
    unsigned long long send_epoch_base;                 // one-kernel exchange: epoch of the previous exchange at this site
};

struct PeerReduceDev {
    int world, rank, n;
    void *remote_slot[kPeerMaxRanks];                   // my slot in rank k's mailbox
    unsigned long long *remote_flag[kPeerMaxRanks];
    const void *local_slots;                            // `world` slots of n values in my arena
    const unsigned long long *local_flag[kPeerMaxRanks];
    unsigned long long send_epoch;
    unsigned int done;                                  // CTAs of the put kernel that have finished their stores
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool peer_spin_until(const unsigned long long *flag, unsigned long long epoch, int *error) {
    if (ld_acquire_sys(flag) >= epoch) return true;
    const unsigned long long t0 = global_timer_ns();
    for (unsigned spins = 0; ld_acquire_sys(flag) < epoch; ++spins) {
        if ((spins & 1023u) == 1023u && global_timer_ns() - t0 > kPeerTimeoutNs) { *error = 1; return false; }
        __nanosleep(32);
    }
    return true;
}

// sender: array[send_index[e]] -> the neighbours' mailboxes, then the epoch flag
__global__ void __launch_bounds__(256) peer_put_kernel(PeerExchangeDev *site, const int *__restrict__ send_index,
                                                       const unsigned long long *__restrict__ array) {
    const int words = site->words, n_nbr = site->n_nbr;
    const int total = site->send_offset[n_nbr] * words;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int e = t / words, w = t - e * words;
        int k = 0;
        while (e >= site->send_offset[k + 1]) ++k;
        site->remote_payload[k][(size_t)(e - site->send_offset[k]) * words + w] = array[(size_t)send_index[e] * words + w];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&site->done, 1u);
        if (prev == gridDim.x - 1) {
            site->done = 0;
            const unsigned long long epoch = ++site->send_epoch;
            __threadfence_system();
            for (int k = 0; k < n_nbr; ++k) st_release_sys(site->remote_flag[k], epoch);
        }
    }
}

// receiver: wait for every neighbour's flag, then mailbox -> halo slots (16 bytes per thread and step; a few CTAs, each
// waiting for itself). The epoch to wait for is this rank's OWN send count at the site: every exchange is a put
// followed by a wait on every rank, so after my put kernel (earlier in the stream) send_epoch is exactly the number
// of the exchange the neighbours are publishing.
template <bool WIDE>      // WIDE: the halo slots start 16-byte aligned (always, except for 24-byte elements behind an odd owned count)
__global__ void __launch_bounds__(256) peer_wait_unpack_kernel(const PeerExchangeDev *site, unsigned long long *__restrict__ halo, int *error) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    const unsigned long long epoch = site->send_epoch;
    if ((int)threadIdx.x < site->n_nbr && !peer_spin_until(site->local_flag[threadIdx.x], epoch, error)) ok = 0;
    __syncthreads();
    if (!ok) return;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const int nw = site->n_recv_words;
    if (WIDE) {
        const uint4 *src = reinterpret_cast<const uint4 *>(site->local_payload);
        uint4 *dst = reinterpret_cast<uint4 *>(halo);
        for (int t = tid; t < nw / 2; t += stride) dst[t] = __ldcv(src + t);                    // written by a peer: never through L1
        if ((nw & 1) && tid == 0) halo[nw - 1] = __ldcv(site->local_payload + nw - 1);
    } else {
        for (int t = tid; t < nw; t += stride) halo[t] = __ldcv(site->local_payload + t);
    }
}

// put + wait + unpack in ONE kernel (the default): every CTA first stores its share of the outgoing entries, the last CTA to
// finish publishes the epoch, then every CTA waits for the neighbours' flags itself and copies its share of the mailbox.
// Nothing in the second half depends on the other CTAs of this kernel, so no grid-wide barrier is needed; and every rank
// puts before it waits, so the ranks cannot wait for each other in a cycle. Half the launches of the two-kernel form.
template <bool WIDE>
__global__ void __launch_bounds__(256) peer_exchange_kernel(PeerExchangeDev *site, const int *__restrict__ send_index,
                                                            const unsigned long long *__restrict__ array, unsigned long long *__restrict__ halo,
                                                            int *error) {
    __shared__ unsigned long long s_epoch;
    __shared__ int ok;
    const int words = site->words, n_nbr = site->n_nbr;
    const int total = site->send_offset[n_nbr] * words;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int e = t / words, w = t - e * words;
        int k = 0;
        while (e >= site->send_offset[k + 1]) ++k;
        site->remote_payload[k][(size_t)(e - site->send_offset[k]) * words + w] = array[(size_t)send_index[e] * words + w];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        ok = 1;
        // the epoch of THIS exchange: the kernels of one stream run one after the other, so send_epoch + 1 is the same for
        // every CTA of this launch until the last one to finish its stores increments it (and only then can the neighbours'
        // matching puts be told apart from the previous exchange's)
        const unsigned long long epoch = *(volatile unsigned long long *)&site->send_epoch_base + 1ULL;
        s_epoch = epoch;
        const unsigned int prev = atomicAdd(&site->done, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            for (int k = 0; k < n_nbr; ++k) st_release_sys(site->remote_flag[k], epoch);
        }
        const unsigned int left = atomicAdd(&site->left, 1u);
        if (left == gridDim.x - 1) {            // the last CTA to have READ the base advances it for the next launch
            site->left = 0;
            site->done = 0;
            site->send_epoch = epoch;
            __threadfence();
            *(volatile unsigned long long *)&site->send_epoch_base = epoch;
        }
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    if ((int)threadIdx.x < n_nbr && !peer_spin_until(site->local_flag[threadIdx.x], epoch, error)) ok = 0;
    __syncthreads();
    if (!ok) return;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const int nw = site->n_recv_words;
    if (WIDE) {
        const uint4 *src = reinterpret_cast<const uint4 *>(site->local_payload);
        uint4 *dst = reinterpret_cast<uint4 *>(halo);
        for (int t = tid; t < nw / 2; t += stride) dst[t] = __ldcv(src + t);
        if ((nw & 1) && tid == 0) halo[nw - 1] = __ldcv(site->local_payload + nw - 1);
    } else {
        for (int t = tid; t < nw; t += stride) halo[t] = __ldcv(site->local_payload + t);
    }
}

// Both reduce kernels are grid-stride over the values: the CG's 8 scalars take one CTA, the right-hand side of a replicated
// multigrid level (up to a few hundred thousand rows) a few dozen.
template <typename T>
__global__ void __launch_bounds__(256) peer_reduce_put_kernel(PeerReduceDev *site, const T *__restrict__ values) {
    const int n = site->n, world = site->world;
    const long long total = (long long)n * world;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / n), c = (int)(t - (long long)k * n);
        ((T *)site->remote_slot[k])[c] = values[c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&site->done, 1u);
        if (prev == gridDim.x - 1) {
            site->done = 0;
            const unsigned long long epoch = ++site->send_epoch;
            __threadfence_system();
            for (int k = 0; k < world; ++k) st_release_sys(site->remote_flag[k], epoch);
        }
    }
}

// values <- sum over ranks, added in rank order (identical bits on every rank)
template <typename T>
__global__ void __launch_bounds__(256) peer_reduce_wait_kernel(PeerReduceDev *site, T *__restrict__ values, int *error) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    const unsigned long long epoch = site->send_epoch;          // my own put for this reduction came first (see above)
    const int n = site->n, world = site->world;
    if ((int)threadIdx.x < world && !peer_spin_until(site->local_flag[threadIdx.x], epoch, error)) ok = 0;
    __syncthreads();
    if (ok) {
        const volatile T *slots = (const volatile T *)site->local_slots;
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
            T sum = slots[c];
            for (int k = 1; k < world; ++k) sum += slots[(size_t)k * n + c];
            values[c] = sum;
        }
    }
}

// small all-reduce (the CG's scalars) in ONE single-CTA kernel: put to every rank, publish, wait for every rank, add up
template <typename T>
__global__ void __launch_bounds__(256) peer_reduce_small_kernel(PeerReduceDev *site, T *__restrict__ values, int *error) {
    __shared__ int ok;
    __shared__ unsigned long long s_epoch;
    const int n = site->n, world = site->world;
    for (int t = threadIdx.x; t < n * world; t += blockDim.x) {
        const int k = t / n, c = t - k * n;
        ((T *)site->remote_slot[k])[c] = values[c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        ok = 1;
        const unsigned long long epoch = ++site->send_epoch;
        s_epoch = epoch;
        __threadfence_system();
        for (int k = 0; k < world; ++k) st_release_sys(site->remote_flag[k], epoch);
    }
    __syncthreads();
    if ((int)threadIdx.x < world && !peer_spin_until(site->local_flag[threadIdx.x], s_epoch, error)) ok = 0;
    __syncthreads();
    if (ok) {
        const volatile T *slots = (const volatile T *)site->local_slots;
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            T sum = slots[c];
            for (int k = 1; k < world; ++k) sum += slots[(size_t)k * n + c];
            values[c] = sum;
        }
    }
}

class PeerTransport : public Transport {
public:
    std::unique_ptr<Transport> boot;       // NCCL (processes) or in-process (threads): setup-time all-gathers only
    bool same_process = false;
    int rank = 0, world = 1;

    ~PeerTransport() override { release(); }

    int init(std::unique_ptr<Transport> bootstrap, bool same_proc, int rank_, int world_) {
        boot = std::move(bootstrap);
        same_process = same_proc;
        rank = rank_;
        world = world_;
        if (world > kPeerMaxRanks) { error = "peer transport: too many ranks"; return -1; }
        if (cudaMallocHost(&error_flag, sizeof(int)) != cudaSuccess) { error = "peer transport: cudaMallocHost failed"; return -1; }
        *error_flag = 0;
        return 0;
    }

    bool capturable() const override { return true; }
    // Large reductions (the right-hand side of the replicated multigrid levels, megabytes) go through the bootstrap transport's
    // all-reduce when that is NCCL: every rank storing its whole vector into every other rank's mailbox would move world x the
    // data a ring / tree all-reduce moves. NCCL opens its connections on first use, hence the warm-up before any graph capture.
    bool needs_warm_up() const override { return !same_process; }
    int poll_error() override {
        if (error_flag && *error_flag) { error = "peer transport: timed out waiting for a neighbour's data"; return -1; }
        return 0;
    }
    int allgather_host(cudaStream_t stream, const void *in, size_t bytes, void *out) override { return boot->allgather_host(stream, in, bytes, out); }
    int barrier(cudaStream_t stream) override {
        char token = 0;
        std::vector<char> tokens((size_t)world);
        if (boot->allgather_host(stream, &token, 1, tokens.data())) { error = boot->error; return -1; }
        return 0;
    }

    int configure(cudaStream_t stream, const std::vector<SiteSpec> &sites) override {
        if ((int)sites.size() > kPeerMaxSites) { error = "peer transport: too many call sites"; return -1; }
        if (cudaStreamSynchronize(stream) != cudaSuccess) { error = "peer transport: sync failed"; return -1; }
        // ---- layout of MY arena
        const size_t flag_bytes = sizeof(unsigned long long) * kPeerMaxSites * kPeerMaxRanks;
        size_t off = flag_bytes;
        std::vector<size_t> site_off(sites.size(), 0);
        for (size_t s = 0; s < sites.size(); ++s) {                      // reduce mailboxes first: same offsets on every rank
            const SiteSpec &sp = sites[s];
            if (sp.kind != SiteSpec::REDUCE_F64 && sp.kind != SiteSpec::REDUCE_F32) continue;
            site_off[s] = off;
            off += align16((size_t)world * sp.n * (sp.kind == SiteSpec::REDUCE_F64 ? 8 : 4));
        }
        for (size_t s = 0; s < sites.size(); ++s) {
            const SiteSpec &sp = sites[s];
            if (sp.kind != SiteSpec::EXCHANGE) continue;
            if ((int)sp.plan->neighbor_rank.size() > kPeerMaxRanks) { error = "peer transport: too many neighbours"; return -1; }
            site_off[s] = off;
            off += align16((size_t)sp.plan->n_halo() * sp.elem_bytes);
        }
        const size_t need = off;
        // ---- (re)allocate and publish: pointer / IPC handle + where each sender's segment starts
        struct Blob {
            unsigned long long arena_ptr;
            long long pid;
            cudaIpcMemHandle_t handle;
            int device;
            unsigned long long segment[kPeerMaxSites][kPeerMaxRanks];    // byte offset in my arena, ~0 = not a neighbour at this site
        };
        const bool fresh = need > arena_bytes;
        if (fresh) {
            close_peers();
            if (arena) cudaFree(arena);
            arena = nullptr;
            arena_bytes = need + need / 2 + 4096;
            if (cudaMalloc(&arena, arena_bytes) != cudaSuccess) { error = "peer transport: arena allocation failed"; arena_bytes = 0; return -1; }
        }
        if (cudaMemset(arena, 0, flag_bytes) != cudaSuccess) { error = "peer transport: memset failed"; return -1; }
        std::vector<Blob> blobs((size_t)world);
        Blob mine;
        memset(&mine, 0, sizeof(mine));
        mine.arena_ptr = (unsigned long long)(uintptr_t)arena;
        mine.pid = (long long)getpid();
        cudaGetDevice(&mine.device);
        if (!same_process && cudaIpcGetMemHandle(&mine.handle, arena) != cudaSuccess) { error = "peer transport: cudaIpcGetMemHandle failed"; return -1; }
        for (int s = 0; s < kPeerMaxSites; ++s) for (int q = 0; q < kPeerMaxRanks; ++q) mine.segment[s][q] = ~0ULL;
        for (size_t s = 0; s < sites.size(); ++s) {
            const SiteSpec &sp = sites[s];
            if (sp.kind != SiteSpec::EXCHANGE) continue;
            for (size_t k = 0; k < sp.plan->neighbor_rank.size(); ++k)
                mine.segment[s][sp.plan->neighbor_rank[k]] = site_off[s] + (size_t)sp.plan->recv_offset[k] * sp.elem_bytes;
        }
        if (boot->allgather_host(stream, &mine, sizeof(Blob), blobs.data())) { error = boot->error; return -1; }
        // ---- map the peers' arenas (every configure: a peer may have re-allocated)
        close_peers();
        peer_base.assign((size_t)world, nullptr);
        for (int q = 0; q < world; ++q) {
            if (q == rank) { peer_base[(size_t)q] = arena; continue; }
            if (same_process) { peer_base[(size_t)q] = (char *)(uintptr_t)blobs[(size_t)q].arena_ptr; continue; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, blobs[(size_t)q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                error = std::string("peer transport: cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(cudaGetLastError());
                return -1;
            }
            peer_base[(size_t)q] = (char *)p;
            opened.push_back(p);
        }
        // ---- per-site device descriptors
        std::vector<PeerExchangeDev> ex(sites.size());
        std::vector<PeerReduceDev> rd(sites.size());
        memset(ex.data(), 0, sizeof(PeerExchangeDev) * ex.size());
        memset(rd.data(), 0, sizeof(PeerReduceDev) * rd.size());
        auto flag_of = [&](int owner_rank, size_t site, int writer) {
            return (unsigned long long *)(peer_base[(size_t)owner_rank]) + site * kPeerMaxRanks + (size_t)writer;
        };
        for (size_t s = 0; s < sites.size(); ++s) {
            const SiteSpec &sp = sites[s];
            if (sp.kind == SiteSpec::EXCHANGE) {
                PeerExchangeDev &d = ex[s];
                d.n_nbr = (int)sp.plan->neighbor_rank.size();
                d.words = sp.elem_bytes / 8;
                for (int k = 0; k <= d.n_nbr; ++k) d.send_offset[k] = sp.plan->send_offset[(size_t)k];
                for (int k = 0; k < d.n_nbr; ++k) {
                    const int q = sp.plan->neighbor_rank[(size_t)k];
                    const unsigned long long seg = blobs[(size_t)q].segment[s][rank];
                    if (seg == ~0ULL) { error = "peer transport: a neighbour does not list this rank at the same call site"; return -1; }
                    d.remote_payload[k] = (unsigned long long *)(peer_base[(size_t)q] + seg);
                    d.remote_flag[k] = flag_of(q, s, rank);
                    d.local_flag[k] = flag_of(rank, s, q);
                }
                d.local_payload = (const unsigned long long *)(arena + site_off[s]);
                d.n_recv_words = sp.plan->n_halo() * d.words;
            } else if (sp.kind == SiteSpec::REDUCE_F64 || sp.kind == SiteSpec::REDUCE_F32) {
                PeerReduceDev &d = rd[s];
                const size_t es = sp.kind == SiteSpec::REDUCE_F64 ? 8 : 4;
                d.world = world; d.rank = rank; d.n = sp.n;
                for (int q = 0; q < world; ++q) {
                    d.remote_slot[q] = peer_base[(size_t)q] + site_off[s] + (size_t)rank * sp.n * es;
                    d.remote_flag[q] = flag_of(q, s, rank);
                    d.local_flag[q] = flag_of(rank, s, q);
                }
                d.local_slots = arena + site_off[s];
            }
        }
        if (ex_dev) cudaFree(ex_dev);
        if (rd_dev) cudaFree(rd_dev);
        ex_dev = nullptr; rd_dev = nullptr;
        const size_t ns = sites.size() ? sites.size() : 1;
        if (cudaMalloc(&ex_dev, sizeof(PeerExchangeDev) * ns) != cudaSuccess || cudaMalloc(&rd_dev, sizeof(PeerReduceDev) * ns) != cudaSuccess) {
            error = "peer transport: descriptor allocation failed"; return -1;
        }
        if (!sites.empty() && (cudaMemcpy(ex_dev, ex.data(), sizeof(PeerExchangeDev) * ex.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
                               cudaMemcpy(rd_dev, rd.data(), sizeof(PeerReduceDev) * rd.size(), cudaMemcpyHostToDevice) != cudaSuccess)) {
            error = "peer transport: descriptor upload failed"; return -1;
        }
        kinds.resize(sites.size());
        for (size_t s = 0; s < sites.size(); ++s) kinds[s] = sites[s].kind;
        // nobody may start putting before everybody has zeroed its flags and mapped the arenas
        char token = 0;
        std::vector<char> tokens((size_t)world);
        if (boot->allgather_host(stream, &token, 1, tokens.data())) { error = boot->error; return -1; }
        return 0;
    }

    int exchange(cudaStream_t stream, int site, const HaloPlan &plan, const int *send_index_dev, char *, char *array, size_t elem_bytes) override {
        if (site < 0 || site >= (int)kinds.size() || kinds[(size_t)site] != SiteSpec::EXCHANGE) { error = "peer transport: call site was not configured"; return -1; }
        if (plan.neighbor_rank.empty()) return 0;
        const int total = plan.n_send() * (int)(elem_bytes / 8);
        int grid = (total + 255) / 256;
        grid = grid < 1 ? 1 : (grid > 128 ? 128 : grid);
        unsigned long long *halo = (unsigned long long *)(array + (size_t)plan.n_owned * elem_bytes);
        static const bool two_kernels = getenv("ARAP_PEER_TWO_KERNELS") != nullptr && atoi(getenv("ARAP_PEER_TWO_KERNELS")) != 0;
        if (!two_kernels) {
            // grid <= 64 CTAs: all of them are resident at once on any GPU this runs on, which the in-kernel wait relies on
            if (grid > 64) grid = 64;
            if (((uintptr_t)halo & 15) == 0) peer_exchange_kernel<true><<<grid, 256, 0, stream>>>(ex_dev + site, send_index_dev, (const unsigned long long *)array, halo, error_flag);
            else peer_exchange_kernel<false><<<grid, 256, 0, stream>>>(ex_dev + site, send_index_dev, (const unsigned long long *)array, halo, error_flag);
            return cudaGetLastError() == cudaSuccess ? 0 : (error = "peer transport: launch failed", -1);
        }
        peer_put_kernel<<<grid, 256, 0, stream>>>(ex_dev + site, send_index_dev, (const unsigned long long *)array);
        const size_t recv_bytes = (size_t)plan.n_halo() * elem_bytes;
        int wgrid = (int)((recv_bytes + 16383) / 16384);
        wgrid = wgrid < 1 ? 1 : (wgrid > 16 ? 16 : wgrid);
        if (((uintptr_t)halo & 15) == 0) peer_wait_unpack_kernel<true><<<wgrid, 256, 0, stream>>>(ex_dev + site, halo, error_flag);
        else peer_wait_unpack_kernel<false><<<wgrid, 256, 0, stream>>>(ex_dev + site, halo, error_flag);
        return cudaGetLastError() == cudaSuccess ? 0 : (error = "peer transport: launch failed", -1);
    }
    int allreduce_sum(cudaStream_t stream, int site, double *dev, int n) override { return reduce<double>(stream, site, dev, n, SiteSpec::REDUCE_F64); }
    int allreduce_sum_f32(cudaStream_t stream, int site, float *dev, int n) override { return reduce<float>(stream, site, dev, n, SiteSpec::REDUCE_F32); }

private:
    char *arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<char *> peer_base;
    std::vector<void *> opened;
    PeerExchangeDev *ex_dev = nullptr;
    PeerReduceDev *rd_dev = nullptr;
    std::vector<int> kinds;
    int *error_flag = nullptr;

    static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
    template <typename T>
    int reduce(cudaStream_t stream, int site, T *dev, int n, int kind) {
        if (site < 0 || site >= (int)kinds.size() || kinds[(size_t)site] != kind) { error = "peer transport: reduce site was not configured"; return -1; }
        if ((long long)n * world <= 2048) {
            peer_reduce_small_kernel<T><<<1, 256, 0, stream>>>(rd_dev + site, dev, error_flag);
            return cudaGetLastError() == cudaSuccess ? 0 : (error = "peer transport: launch failed", -1);
        }
        if (!same_process && (size_t)n * sizeof(T) >= (size_t)256 * 1024) {
            const int rc = sizeof(T) == 8 ? boot->allreduce_sum(stream, site, (double *)dev, n) : boot->allreduce_sum_f32(stream, site, (float *)dev, n);
            if (rc) error = boot->error;
            return rc;
        }
        int pgrid = (int)(((long long)n * world + 4095) / 4096), wgrid = (n + 4095) / 4096;
        pgrid = pgrid < 1 ? 1 : (pgrid > 128 ? 128 : pgrid);
        wgrid = wgrid < 1 ? 1 : (wgrid > 64 ? 64 : wgrid);
        peer_reduce_put_kernel<T><<<pgrid, 256, 0, stream>>>(rd_dev + site, dev);
        peer_reduce_wait_kernel<T><<<wgrid, 256, 0, stream>>>(rd_dev + site, dev, error_flag);
        return cudaGetLastError() == cudaSuccess ? 0 : (error = "peer transport: launch failed", -1);
    }
    void close_peers() {
        for (void *p : opened) cudaIpcCloseMemHandle(p);
        opened.clear();
    }
    void release() {
        close_peers();
        if (ex_dev) cudaFree(ex_dev);
        if (rd_dev) cudaFree(rd_dev);
        if (arena) cudaFree(arena);
        if (error_flag) cudaFreeHost(error_flag);
        ex_dev = nullptr; rd_dev = nullptr; arena = nullptr; error_flag = nullptr;
    }
};

}  // namespace arap
