// partition.cuh -- one mesh partitioned over several GPUs (BASELINE.json configs[4]).
//
// Each rank owns a contiguous block of local vertex indices [0, n_owned) followed by its halo: the one-ring
// neighbours of owned vertices that another rank owns, grouped by owner. All row-parallel kernels run on the owned
// rows only. Three things cross ranks, all through the Transport below:
//   * halo exchange of p' (after every global step), of R as quaternions (after every local step) and of the CG
//     search direction (every CG iteration): pack kernel -> point-to-point send/recv into the halo slots;
//   * all-reduce of the CG's 1-5 partial sums per reduction stage;
//   * nothing else: the multigrid preconditioner is applied per rank on its owned block (block-Jacobi across ranks).
// Transports: NCCL over NVLink (one process per GPU; the library is dlopen'ed so single-GPU use needs no NCCL),
// and an in-process transport (several partitions of one mesh on ONE GPU, one host thread per partition) that the
// single-GPU test tier uses to exercise exactly the same solver code path.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include "halo_plan.h"

#include <condition_variable>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace arap {

// sendbuf[k] = array[send_index[k]] in 8-byte words (element sizes are 16, 24 or 32 bytes)
__global__ void __launch_bounds__(256) halo_pack_kernel(int n_send, int words_per_elem, const int *__restrict__ send_index,
                                                        const unsigned long long *__restrict__ array, unsigned long long *__restrict__ sendbuf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_send * words_per_elem) return;
    const int k = t / words_per_elem, w = t - k * words_per_elem;
    sendbuf[t] = array[(size_t)send_index[k] * words_per_elem + w];
}

// One communication call site of the solver (a halo exchange of one array, or a small all-reduce). Sites are numbered
// the same on every rank; transports that keep per-site state (peer_transport.cuh) are told about them in configure().
struct SiteSpec {
    enum Kind { UNUSED = 0, EXCHANGE, REDUCE_F64, REDUCE_F32 } kind = UNUSED;
    const HaloPlan *plan = nullptr;   // EXCHANGE
    int elem_bytes = 0;               // EXCHANGE: 16, 24 or 32
    int n = 0;                        // REDUCE_*: number of values
};

class Transport {
public:
    virtual ~Transport() {}
    // array[n_owned + recv slots] <- the owners' values. send_index_dev: plan.send_index on the device; scratch: room for
    // plan.n_send() packed elements.
    virtual int exchange(cudaStream_t stream, int site, const HaloPlan &plan, const int *send_index_dev, char *scratch, char *array,
                         size_t elem_bytes) = 0;
    virtual int allreduce_sum(cudaStream_t stream, int site, double *dev, int n) = 0;
    virtual int allreduce_sum_f32(cudaStream_t stream, int site, float *dev, int n) = 0;
    // out = the `bytes` of every rank, in rank order (host buffers; setup-time only)
    virtual int allgather_host(cudaStream_t stream, const void *in, size_t bytes, void *out) = 0;
    virtual int configure(cudaStream_t, const std::vector<SiteSpec> &) { return 0; }
    // true if every call only enqueues work on the stream (no host synchronisation), i.e. a CG iteration with its
    // exchanges and reductions can be captured into a CUDA graph
    virtual bool capturable() const { return false; }
    virtual bool needs_warm_up() const { return false; }
    virtual int poll_error() { return 0; }
    // every rank has reached this point (host-level); only transports whose kernels wait on each other need one
    virtual int barrier(cudaStream_t) { return 0; }
    std::string error;
protected:
    static void pack(cudaStream_t stream, const HaloPlan &plan, const int *send_index_dev, const char *array, char *scratch, size_t elem_bytes) {
        const int n = plan.n_send(), words = (int)(elem_bytes / 8);
        if (n > 0)
            halo_pack_kernel<<<(n * words + 255) / 256, 256, 0, stream>>>(n, words, send_index_dev, (const unsigned long long *)array,
                                                                          (unsigned long long *)scratch);
    }
};

// ---- NCCL (dlopen'ed) --------------------------------------------------------------------------------------------------
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    typedef void *Comm;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *lib = nullptr;

    static NcclApi *get(std::string *err) {
        static NcclApi api;
        static bool tried = false;
        static std::string load_error;
        if (!tried) {
            tried = true;
            const char *env = getenv("ARAP_NCCL_LIB");
            const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
            for (const char *n : names) {
                if (!n) continue;
                api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (api.lib) break;
            }
            if (!api.lib) load_error = "cannot dlopen libnccl.so.2 (set ARAP_NCCL_LIB)";
            else {
#define ARAP_NCCL_SYM(field, name)                                                     \
    *(void **)(&api.field) = dlsym(api.lib, name);                                     \
    if (!api.field) load_error = std::string("libnccl lacks ") + name;
                ARAP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
                ARAP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
                ARAP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
                ARAP_NCCL_SYM(GroupStart, "ncclGroupStart")
                ARAP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
                ARAP_NCCL_SYM(Send, "ncclSend")
                ARAP_NCCL_SYM(Recv, "ncclRecv")
                ARAP_NCCL_SYM(AllReduce, "ncclAllReduce")
                ARAP_NCCL_SYM(AllGather, "ncclAllGather")
                ARAP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ARAP_NCCL_SYM
            }
        }
        if (!load_error.empty()) { if (err) *err = load_error; return nullptr; }
        return &api;
    }
};

class NcclTransport : public Transport {
public:
    NcclApi *api = nullptr;
    NcclApi::Comm comm = nullptr;
    int init(int rank, int world, const void *unique_id) {
        api = NcclApi::get(&error);
        if (!api) return -1;
        NcclApi::UniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        this->world = world;
        const int rc = api->CommInitRank(&comm, world, id, rank);
        if (rc != 0) { error = std::string("ncclCommInitRank: ") + api->GetErrorString(rc); return -1; }
        return 0;
    }
    ~NcclTransport() override { if (api && comm) api->CommDestroy(comm); }
    int check(int rc, const char *what) {
        if (rc == 0) return 0;
        error = std::string(what) + ": " + api->GetErrorString(rc);
        return -1;
    }
    int world = 1;
    bool capturable() const override { return getenv("ARAP_NCCL_NO_GRAPH") == nullptr; }
    bool needs_warm_up() const override { return true; }      // connections are opened on first use: not inside a capture
    int exchange(cudaStream_t stream, int, const HaloPlan &plan, const int *send_index_dev, char *scratch, char *array, size_t elem_bytes) override {
        if (plan.neighbor_rank.empty()) return 0;
        pack(stream, plan, send_index_dev, array, scratch, elem_bytes);
        if (check(api->GroupStart(), "ncclGroupStart")) return -1;
        for (size_t k = 0; k < plan.neighbor_rank.size(); ++k) {
            const size_t ns = (size_t)(plan.send_offset[k + 1] - plan.send_offset[k]) * elem_bytes;
            const size_t nr = (size_t)(plan.recv_offset[k + 1] - plan.recv_offset[k]) * elem_bytes;
            if (ns && check(api->Send(scratch + (size_t)plan.send_offset[k] * elem_bytes, ns, /*ncclChar*/ 0, plan.neighbor_rank[k], comm, stream), "ncclSend")) return -1;
            if (nr && check(api->Recv(array + ((size_t)plan.n_owned + plan.recv_offset[k]) * elem_bytes, nr, 0, plan.neighbor_rank[k], comm, stream), "ncclRecv")) return -1;
        }
        return check(api->GroupEnd(), "ncclGroupEnd");
    }
    int allreduce_sum(cudaStream_t stream, int, double *dev, int n) override {
        return check(api->AllReduce(dev, dev, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, stream), "ncclAllReduce");
    }
    int allreduce_sum_f32(cudaStream_t stream, int, float *dev, int n) override {
        return check(api->AllReduce(dev, dev, (size_t)n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, stream), "ncclAllReduce");
    }
    int allgather_host(cudaStream_t stream, const void *in, size_t bytes, void *out) override {
        char *dev = nullptr;
        if (cudaMalloc(&dev, bytes * ((size_t)world + 1)) != cudaSuccess) { error = "allgather: cudaMalloc failed"; return -1; }
        int rc = 0;
        if (cudaMemcpyAsync(dev, in, bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) rc = -1;
        if (!rc) rc = check(api->AllGather(dev, dev + bytes, bytes, /*ncclChar*/ 0, comm, stream), "ncclAllGather");
        if (!rc && (cudaMemcpyAsync(out, dev + bytes, bytes * (size_t)world, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                    cudaStreamSynchronize(stream) != cudaSuccess)) { error = "allgather: copy failed"; rc = -1; }
        cudaFree(dev);
        return rc;
    }
};

// ---- in-process transport: P partitions on one GPU, one host thread each ------------------------------------------------
struct LocalGroup {
    int world = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    long generation = 0;
    std::vector<const char *> sendbuf;
    std::vector<const HaloPlan *> plan;
    std::vector<std::vector<double>> values;
    std::vector<std::vector<unsigned char>> blobs;

    void barrier() {
        std::unique_lock<std::mutex> lock(mu);
        const long gen = generation;
        if (++arrived == world) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lock, [&] { return generation != gen; });
    }
    static std::shared_ptr<LocalGroup> get(int key, int world) {
        static std::mutex reg_mu;
        static std::map<int, std::weak_ptr<LocalGroup>> registry;
        std::lock_guard<std::mutex> lock(reg_mu);
        std::shared_ptr<LocalGroup> g = registry[key].lock();
        if (!g) {
            g = std::make_shared<LocalGroup>();
            g->world = world;
            g->sendbuf.assign((size_t)world, nullptr);
            g->plan.assign((size_t)world, nullptr);
            g->values.assign((size_t)world, std::vector<double>());
            g->blobs.assign((size_t)world, std::vector<unsigned char>());
            registry[key] = g;
        }
        return g;
    }
};

class LocalTransport : public Transport {
public:
    std::shared_ptr<LocalGroup> group;
    int rank = 0;
    int init(int rank_, int world, int key) {
        rank = rank_;
        group = LocalGroup::get(key, world);
        if (group->world != world) { error = "in-process group: world size mismatch"; return -1; }
        return 0;
    }
    int exchange(cudaStream_t stream, int, const HaloPlan &plan, const int *send_index_dev, char *scratch, char *array, size_t elem_bytes) override {
        pack(stream, plan, send_index_dev, array, scratch, elem_bytes);
        const char *sendbuf = scratch;
        if (cudaStreamSynchronize(stream) != cudaSuccess) { error = "sync before exchange"; return -1; }
        group->sendbuf[(size_t)rank] = sendbuf;
        group->plan[(size_t)rank] = &plan;
        group->barrier();
        for (size_t k = 0; k < plan.neighbor_rank.size(); ++k) {
            const int peer = plan.neighbor_rank[k];
            const HaloPlan *pp = group->plan[(size_t)peer];
            size_t slot = pp->neighbor_rank.size();
            for (size_t q = 0; q < pp->neighbor_rank.size(); ++q) if (pp->neighbor_rank[q] == rank) slot = q;
            if (slot == pp->neighbor_rank.size()) { error = "in-process exchange: asymmetric neighbour lists"; return -1; }
            const size_t n = (size_t)(plan.recv_offset[k + 1] - plan.recv_offset[k]);
            if ((size_t)(pp->send_offset[slot + 1] - pp->send_offset[slot]) != n) { error = "in-process exchange: send/recv counts differ"; return -1; }
            if (n && cudaMemcpyAsync(array + ((size_t)plan.n_owned + plan.recv_offset[k]) * elem_bytes,
                                     group->sendbuf[(size_t)peer] + (size_t)pp->send_offset[slot] * elem_bytes, n * elem_bytes,
                                     cudaMemcpyDeviceToDevice, stream) != cudaSuccess) { error = "in-process exchange: copy failed"; return -1; }
        }
        if (cudaStreamSynchronize(stream) != cudaSuccess) { error = "sync after exchange"; return -1; }
        group->barrier();      // nobody repacks its send buffer before every reader is done
        return 0;
    }
    template <typename T>
    int allreduce_host(cudaStream_t stream, T *dev, int n) {
        std::vector<double> mine((size_t)n);
        std::vector<T> raw((size_t)n);
        if (cudaMemcpyAsync(raw.data(), dev, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess) { error = "in-process allreduce: D2H failed"; return -1; }
        for (int c = 0; c < n; ++c) mine[(size_t)c] = (double)raw[(size_t)c];
        group->values[(size_t)rank] = mine;
        group->barrier();
        std::vector<double> sum((size_t)n, 0.0);
        for (int r = 0; r < group->world; ++r)                 // fixed rank order: every partition gets identical bits
            for (int c = 0; c < n; ++c) sum[(size_t)c] += group->values[(size_t)r][(size_t)c];
        for (int c = 0; c < n; ++c) raw[(size_t)c] = (T)sum[(size_t)c];
        if (cudaMemcpyAsync(dev, raw.data(), sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess) { error = "in-process allreduce: H2D failed"; return -1; }
        group->barrier();
        return 0;
    }
    int allreduce_sum(cudaStream_t stream, int, double *dev, int n) override { return allreduce_host<double>(stream, dev, n); }
    int allreduce_sum_f32(cudaStream_t stream, int, float *dev, int n) override { return allreduce_host<float>(stream, dev, n); }
    int allgather_host(cudaStream_t, const void *in, size_t bytes, void *out) override {
        group->blobs[(size_t)rank].assign((const unsigned char *)in, (const unsigned char *)in + bytes);
        group->barrier();
        for (int r = 0; r < group->world; ++r) {
            if (group->blobs[(size_t)r].size() != bytes) { error = "in-process allgather: size mismatch"; return -1; }
            memcpy((char *)out + bytes * (size_t)r, group->blobs[(size_t)r].data(), bytes);
        }
        group->barrier();
        return 0;
    }
};

}  // namespace arap
