// engine.cu -- host side of libarap_b200.so: device-memory ownership, kernel sequencing, and the
// extern "C" entry points declared in include/arap_b200.h. Thin by design: every number is computed
// by the kernels in kernels.cuh; the host only allocates, launches and copies.
//
// Mirrors the control flow of the reference's deform() (reference inc/deform/arap.h:101-138):
//   prepare()  = the `_dirty` block (:102-120),  iterate() = the loop (:122-129),
//   get_positions() = the write-back (:133-135).
#include "../../include/arap_b200.h"
#include "kernels.cuh"
#include "mg_kernels.cuh"
#include "tile_kernels.cuh"
#include "mg_setup_device.cuh"
#include "mg_setup.h"
#include "mg_partition.h"
#include "partition.cuh"
#include "peer_transport.cuh"
#include "batch_gemm_tc.cuh"
#include "../../inc/deform/detail/se3_spline.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

namespace arap {

// Default stopping rule of the multigrid solver (include/arap_b200.h, position_tolerance); chosen from the sweep in
// profiles/r01_h_stopping_rule.txt and the margins of the GPU tests: positions stay ~100x inside the parity bar on regular AND badly
// conditioned meshes; the relative ENERGY is the tighter bar on gently bent fine grids (3e-8 left a 200 x 160 grid at 7.8e-7 of the
// allowed 1e-6 after 4 iterations, 1e-8 at 7.9e-8), which is what decided between the two.
static const double kDefaultPositionTolerance = 1e-8;

static const char *kKernelNames[ARAP_K_COUNT_MAX] = {
    "weights_count", "weights_fill", "row_sort_merge", "csr_compact", "scan",
    "init_state", "diagonal", "local_step", "rhs_residual", "cg_spmv",
    "cg_update", "cg_direction", "apply_update", "energy", "misc",
    "mg_fine_residual", "mg_fine_postsmooth", "mg_csr_residual", "mg_restrict_presmooth", "mg_prolong_add",
    "mg_csr_postsmooth", "mg_dense_solve", "cg_update_mg", "cg_direction_mg", "cg_dot", "halo_exchange", "cg_finalize", "local_step_redo", "mg_tail",
    "allreduce_scalars", "allreduce_level"};

static thread_local std::string g_create_error;

struct Status {
    int code = ARAP_OK;
    std::string msg;
};

#define ARAP_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (expr);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            return fail(ARAP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));       \
        }                                                                                            \
    } while (0)

static inline int grid_for(size_t n) { return n ? (int)((n + kBlock - 1) / kBlock) : 1; }   // kernels bound-check; never a 0-CTA launch
// Kernels that end in a grid-wide reduction run as persistent grid-stride kernels: exactly as many CTAs
// as fit on the GPU at once (occupancy API, per kernel), so the number of per-block partials -- and the tail
// in which the last CTA sums them -- stays small. Measured (profiles/r01_d_variants.txt): 10 % per CG iteration.
#ifndef ARAP_PERSISTENT_CTAS_PER_SM
#define ARAP_PERSISTENT_CTAS_PER_SM 0      /* 0 = ask the occupancy API per kernel */
#endif

template <typename T>
struct DeviceBuffer {
    T *ptr = nullptr;
    size_t count = 0;
    ~DeviceBuffer() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
    cudaError_t ensure(size_t n) {
        if (n <= count && ptr) return cudaSuccess;
        release();
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc(&ptr, n * sizeof(T));
        if (e == cudaSuccess) count = n;
        return e;
    }
};

class EngineBase {
public:
    virtual ~EngineBase() {}
    virtual int set_constraints(int n, const int *idx, const void *xyz, int scalar_bytes) = 0;
    virtual int set_rigid_constraints(int n, int batch, int member_stride, const int *idx, const void *rest, int scalar_bytes,
                                      const double *transforms16) = 0;
    virtual int prepare(const void *rest_xyz, int scalar_bytes) = 0;
    virtual int iterate(int n, bool defer_sync = false) = 0;
    virtual int finish_iterate() = 0;
    virtual int get_positions(void *out, int scalar_bytes) = 0;
    virtual int deform_async(void *out, int scalar_bytes, int n) = 0;
    virtual int deform_wait() = 0;
    virtual bool deform_wait_pending() = 0;
    virtual int get_csr_nnz(int *nnz) = 0;
    virtual int get_csr(int *rowptr, int *colidx, void *weights) = 0;
    virtual int get_free_map(int *free_idx, int *n_free) = 0;
    virtual int get_rotations(void *rot9) = 0;
    virtual int get_rhs(double *out) = 0;
    virtual int get_render_buffers(float *positions, float *normals, int location) = 0;
    virtual int energy(double *e) = 0;
    virtual int attach_partition(const arap_partition_plan *p, int rank, int world, int kind, const void *id, int id_bytes) = 0;
    virtual int set_global_mesh(const arap_global_mesh *g) = 0;
    virtual int comm_benchmark(int rounds, double *us_exchange, double *us_allreduce) = 0;
    // arap_batch_*: the mesh is `members` disjoint copies of a `member_vertices`-vertex mesh, member-major
    void set_batch_layout(int members, int member_vertices_) { batch_members = members; member_vertices = member_vertices_; }
    int batch_members = 1, member_vertices = 0;

    int fail(int code, const std::string &msg) {
        last_error = msg;
        return code;
    }
    void activate() { cudaSetDevice(device); }   // every C-ABI entry runs on the handle's device

    // ---- profiling -------------------------------------------------------------------------------
    struct TimedLaunch { int id; cudaEvent_t start, stop; };
    void begin_launch(int id) {
        if (capturing) { graph_counts[id] += 1; return; }     // replayed per graph launch, see launch_cg_graph()
        profile.launches[id] += 1;
        if (!profile_events) return;
        TimedLaunch t;
        t.id = id;
        cudaEventCreate(&t.start);
        cudaEventCreate(&t.stop);
        cudaEventRecord(t.start, stream);
        pending.push_back(t);
    }
    void end_launch() {
        if (capturing || !profile_events) return;
        cudaEventRecord(pending.back().stop, stream);
        if (pending.size() >= 4096) collect_profile();
    }
    void collect_profile() {
        if (pending.empty()) return;
        cudaStreamSynchronize(stream);
        for (auto &t : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) profile.milliseconds[t.id] += ms;
            cudaEventDestroy(t.start);
            cudaEventDestroy(t.stop);
        }
        pending.clear();
    }

    // Launch of an iteration kernel (every one of them starts with pdl_wait(), device_utils.cuh) with programmatic stream
    // serialisation: inside a captured graph this becomes a programmatic edge, and the kernel's CTAs are scheduled while the
    // previous kernel drains. Not used while per-launch events are being recorded (they would sit between the two kernels)
    // and not for the first kernel after something that is not one of these kernels (pdl_next_plain).
    template <typename... Exp, typename... Act>
    void launch_pdl(void (*kernel)(Exp...), unsigned grid, unsigned block, size_t smem, Act &&... args) {
        cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(block, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        // pdl_mode 1: every iteration kernel; 2: only the small grids of the coarse multigrid levels (a persistent full-GPU
        // kernel whose CTAs trickle in while its predecessor drains ends up unevenly spread over the SMs)
        const bool pdl = pdl_mode > 0 && (pdl_mode == 1 || grid <= (unsigned)(2 * sm_count)) && !pdl_next_plain && !profile_events;
        cfg.attrs = pdl ? &attr : nullptr;
        cfg.numAttrs = pdl ? 1 : 0;
        pdl_next_plain = false;
        cudaLaunchKernelEx(&cfg, kernel, std::forward<Act>(args)...);
    }
    // Measured at 1M vertices (profiles/r02_experiments.txt): 2.08 ms per ARAP iteration without, 2.14 ms with the attribute on
    // every kernel -- off by default, ARAP_PDL=1|2 switches it on.
    int pdl_mode = getenv("ARAP_PDL") ? atoi(getenv("ARAP_PDL")) : 0;
    bool pdl_next_plain = true;

    int n_vertices = 0, n_faces = 0;
    int device = 0;
    int sm_count = 148;
    std::map<const void *, int> resident_ctas;     // per kernel: CTAs of kBlock threads resident per SM
    template <typename K>
    int reduce_grid(K kernel, size_t n, size_t dyn_smem = 0) {
        int per_sm = ARAP_PERSISTENT_CTAS_PER_SM;
        if (per_sm <= 0) {
            const void *key = (const void *)kernel;
            auto it = resident_ctas.find(key);
            if (it == resident_ctas.end()) {
                int occ = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, dyn_smem) != cudaSuccess || occ <= 0) occ = 4;
                it = resident_ctas.emplace(key, occ).first;
            }
            per_sm = it->second;
        }
        const int full = grid_for(n), cap = sm_count * per_sm;
        return full < cap ? (full > 0 ? full : 1) : cap;
    }
    cudaStream_t stream = nullptr;
    arap_options opt;
    bool dirty = true;
    bool prepared = false;
    std::string last_error;
    arap_profile profile;
    bool profile_events = false;
    bool capturing = false;
    int64_t graph_counts[ARAP_K_COUNT_MAX] = {0};
    std::vector<TimedLaunch> pending;
    arap_solver_stats stats;
    cudaEvent_t timer_start = nullptr, timer_stop = nullptr;
};

#define LAUNCH(id, kernel, grid, ...)                                                \
    do {                                                                             \
        begin_launch(id);                                                            \
        kernel<<<(grid), kBlock, 0, stream>>>(__VA_ARGS__);                          \
        end_launch();                                                                \
    } while (0)

#define LAUNCH_PDL(id, kernel, grid, ...)                                            \
    do {                                                                             \
        begin_launch(id);                                                            \
        launch_pdl(kernel, (unsigned)(grid), (unsigned)kBlock, 0, __VA_ARGS__);      \
        end_launch();                                                                \
    } while (0)

#define LAUNCH_PDL_SMEM(id, kernel, grid, smem, ...)                                  \
    do {                                                                             \
        begin_launch(id);                                                            \
        launch_pdl(kernel, (unsigned)(grid), (unsigned)kBlock, (size_t)(smem), __VA_ARGS__); \
        end_launch();                                                                \
    } while (0)

// Communication call sites of the partitioned solver (partition.cuh, SiteSpec): the same numbers on every rank.
enum {
    SITE_CUR4 = 0, SITE_QUAT, SITE_CG_D, SITE_MG0_X_PRE, SITE_MG0_R, SITE_MG0_X_POST, SITE_COARSE_B,
    SITE_RED_BASE = 8,          // + CgStage
    SITE_LEVEL_BASE = 16,       // + 4 * level + {0: x after pre-smoothing, 1: residual, 2: level result, 3: x after prolongation}
    SITE_COUNT = 64
};

// One multigrid level on the device. Level 0 keeps no explicit A (matrix-free on the one-ring CSR).
struct MgLevelDev {
    int n = 0;
    double omega = 2.0 / 3.0;
    DeviceBuffer<int> a_rowptr, a_colidx, p_rowptr, p_colidx, r_rowptr, r_colidx;
    DeviceBuffer<float> a_val, p_val, r_val, inv_diag;     // the V-cycle runs in fp32 (mg_kernels.cuh)
    DeviceBuffer<MgVec> b, x, x2, r;          // x: iterate before post-smoothing, x2: the level's result
    int a_lanes = 1, r_lanes = 1;             // threads per row for A and R kernels
    // partitioned mode with a global hierarchy: n = rows this rank owns, n_ext = owned + halo (vector length)
    int n_ext = 0;
    HaloPlan plan;
    DeviceBuffer<int> send_index;
};

template <typename T>
static cudaError_t upload_as_float(DeviceBuffer<float> &dst, const std::vector<T> &src, cudaStream_t stream, std::vector<float> &scratch) {
    scratch.assign(src.begin(), src.end());
    cudaError_t e = dst.ensure(scratch.size());
    if (e != cudaSuccess || scratch.empty()) return e;
    e = cudaMemcpyAsync(dst.ptr, scratch.data(), scratch.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);      // scratch is reused by the next upload
}

static inline int pick_lanes(size_t nnz, size_t rows) {
    // threads per row of the coarse-level CSR kernels: the smallest power of two >= (average row length / kRowsPerLane)
    static const double per_lane = getenv("ARAP_NNZ_PER_LANE") ? atof(getenv("ARAP_NNZ_PER_LANE")) : 3.0;
    const double avg = rows ? (double)nnz / (double)rows : 0.0;
    int lanes = 1;
    while (lanes < 32 && (double)lanes * per_lane < avg) lanes *= 2;
    return lanes;
}

#define ARAP_DISPATCH_LANES(lanes, CALL)                                   \
    switch (lanes) {                                                       \
        case 1: { constexpr int LN = 1; CALL; } break;                     \
        case 2: { constexpr int LN = 2; CALL; } break;                     \
        case 4: { constexpr int LN = 4; CALL; } break;                     \
        case 8: { constexpr int LN = 8; CALL; } break;                     \
        case 16: { constexpr int LN = 16; CALL; } break;                   \
        default: { constexpr int LN = 32; CALL; } break;                   \
    }

template <typename T>
static cudaError_t upload_vector(DeviceBuffer<T> &dst, const std::vector<T> &src, cudaStream_t stream) {
    cudaError_t e = dst.ensure(src.size());
    if (e != cudaSuccess || src.empty()) return e;
    return cudaMemcpyAsync(dst.ptr, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
}

template <typename S>
class Engine : public EngineBase {
public:
    // ---- device state (see kernels.cuh for the layout) -------------------------------------------
    DeviceBuffer<int> faces;                       // _faces
    DeviceBuffer<S> rest_xyz;                      // rest pose as uploaded (V x 3)
    DeviceBuffer<unsigned char> is_constrained;    // key set of _constrainedLocations
    DeviceBuffer<S> target_xyz;                    // values of _constrainedLocations (dense, V x 3)
    DeviceBuffer<int> rowptr, colidx;              // _edgeWeights
    DeviceBuffer<S> weight;
    DeviceBuffer<int> free_idx;                    // _freeIdxMap
    // internal (hot) order: perm[internal] = user index; the iteration kernels only ever see the hot CSR
    DeviceBuffer<int> perm, iperm, hot_rowptr, hot_colidx;
    DeviceBuffer<S> hot_weight;
    DeviceBuffer<float> hot_weight_f32;            // fp32 copy for the multigrid preconditioner
    DeviceBuffer<unsigned char> free_mask;         // 1 = free row (internal order)
    bool have_perm = false;
    // tiles of kTile consecutive rows with their halo lists and tile-local column indices (tile_kernels.cuh)
    DeviceBuffer<int> tile_halo, tile_halo_count, tile_scalars;      // tile_scalars: [0] largest halo, [1] a tile did not fit
    DeviceBuffer<unsigned short> tile_colidx;
    bool tiles_built = false, use_tiles = false, renumbered = false;
    int n_tiles = 0, tile_max_halo = 0;
    TileView tile_view() const { return TileView{tile_halo.ptr, tile_halo_count.ptr, tile_colidx.ptr, n_tiles}; }
    size_t tile_smem(size_t record_bytes, int arrays) const { return (size_t)(kTile + tile_max_halo) * record_bytes * (size_t)arrays; }
    std::vector<int> faces_host;                   // kept until the vertex order has been decided
    std::vector<int> mg_visit_order;               // Morton sequence of internal indices: aggregation order of the fine level
    DeviceBuffer<Vec4T<S>> rest4, cur4, quat;      // _p, _pprime, _rotations
    DeviceBuffer<double> inv_diag;
    DeviceBuffer<Vec3d> cg_r, cg_d, cg_ad, cg_x, cg_w;      // residual, direction, A d (s), solution, and w = A z (multigrid CG)
    DeviceBuffer<double> partials;
    DeviceBuffer<unsigned> counter;
    DeviceBuffer<CgScalars> cg;
    DeviceBuffer<double> energy_dev;
    DeviceBuffer<int> redo_list, redo_count;       // local step: vertices that need the Jacobi SVD fallback
    DeviceBuffer<unsigned> redo_done;
    // scratch for the CSR build
    DeviceBuffer<int> row_count, raw_rowptr, row_cursor, raw_col, unique_count, scan_tiles, flags;
    DeviceBuffer<S> raw_val;
    DeviceBuffer<unsigned> raw_tag;
    DeviceBuffer<unsigned char> staging;           // uploads / downloads in a foreign scalar type

    std::vector<std::unique_ptr<MgLevelDev>> mg;   // multigrid hierarchy (empty -> Jacobi preconditioner)
    DeviceBuffer<float> mg_coarse_inv;             // dense inverse of the coarsest operator, rows padded to mg_coarse_ld floats
    int mg_coarse_ld = 0;
    bool mg_dense = false;
    bool use_mg = false;
    std::vector<unsigned char> mg_mask;            // constrained(+halo) mask the hierarchy was built for
    int mg_nnz = -1;
    double mg_complexity = 0;
    bool mg_stale = false, mg_fresh = false;
    int mg_fresh_iterations = 0;                   // CG iterations of the first global step after a fresh setup
    // partitioned mode (partition.cuh): rows [0, n_rows) are owned, [n_rows, n_vertices) is the halo
    int n_rows = 0;
    HaloPlan plan;
    DeviceBuffer<int> send_index_dev;
    DeviceBuffer<unsigned char> halo_sendbuf;
    std::unique_ptr<Transport> transport;
    int my_rank = 0, world_size = 1;
    // the GLOBAL mesh (arap_partition_set_global_mesh): lets every rank build the same global multigrid hierarchy
    struct GlobalMesh {
        int n_vertices = 0, n_faces = 0;
        std::vector<int> faces, owner, local_to_global;
        std::vector<double> rest;
        std::vector<unsigned char> constrained;
    } global_mesh;
    bool have_global = false;
    bool mg_global = false;                        // the current hierarchy is this rank's share of the global one
    int mg_first_replicated = 0;                   // global hierarchy: levels >= this one are kept whole on every rank
    std::vector<unsigned char> mg_global_mask;     // global constrained mask the hierarchy was built for
    DeviceBuffer<unsigned char> mg_sendbuf;
    // the coarse tail of the V-cycle as one cluster kernel (mg_kernels.cuh, mg_tail_kernel): levels [tail_first, last]
    DeviceBuffer<float> mg_a_pack, mg_b_pack;      // batch GEMM operands as shared-memory stage images (batch_gemm_tc.cuh)
    bool mg_batch_tc = false;                      // batches: the GEMM runs on the tensor cores (tcgen05), else SIMT fp32
    bool mg_batch_dense = false;                   // batches: one member's dense inverse is the whole preconditioner (mg_batch_dense_kernel)
    double length_scale = 0;                       // bbox diagonal of the first rest pose (position-error stopping rule)
    MgTailArgs tail_args;
    int tail_first = 0;                            // 0 = no tail kernel
    int tail_cluster = 8;
    cudaGraph_t cg_graph = nullptr;                // one CG iteration (preconditioner included), replayed per iteration
    cudaGraphExec_t cg_graph_exec = nullptr;
    bool have_warm_rotations = false;              // quat[] holds the previous iteration's R_i

    bool comm_counting = false, comm_counted = false;   // cross-GPU operations of one CG iteration (arap_solver_stats)
    int comm_exchanges = 0, comm_allreduces = 0;
    long long comm_bytes = 0;
    void comm_count_begin() { comm_counting = true; comm_exchanges = comm_allreduces = 0; comm_bytes = 0; }
    void comm_count_end() {
        comm_counting = false;
        comm_counted = true;
        stats.comm_exchanges_per_cg_iteration = comm_exchanges;
        stats.comm_allreduces_per_cg_iteration = comm_allreduces;
        stats.comm_halo_bytes_per_cg_iteration = comm_bytes;
    }

    int nnz = 0;
    int n_free = 0;
    int n_constrained_calls = 0;
    CgScalars *cg_host = nullptr;                  // pinned mirror for convergence polls
    cudaEvent_t poll_event[2] = {nullptr, nullptr};

    void destroy_cg_graph() {
        if (cg_graph_exec) cudaGraphExecDestroy(cg_graph_exec);
        if (cg_graph) cudaGraphDestroy(cg_graph);
        cg_graph_exec = nullptr;
        cg_graph = nullptr;
    }

    ~Engine() override {
        collect_profile();
        destroy_cg_graph();
        destroy_step_graph();
        if (cg_host) cudaFreeHost(cg_host);
        for (auto &e : poll_event) if (e) cudaEventDestroy(e);
        if (timer_start) cudaEventDestroy(timer_start);
        if (timer_stop) cudaEventDestroy(timer_stop);
        for (int k = 0; k < 2; ++k) { if (snap_ready[k]) cudaEventDestroy(snap_ready[k]); if (copy_done[k]) cudaEventDestroy(copy_done[k]); }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }

    int init(const int *faces_host_in, int nf, int nv, const arap_options &o) {
        opt = o;
        n_vertices = nv;
        n_faces = nf;
        std::memset(&profile, 0, sizeof(profile));
        std::memset(&stats, 0, sizeof(stats));
        profile_events = o.profile != 0;
        int count = 0;
        ARAP_CUDA(cudaGetDeviceCount(&count));
        if (count <= 0) return fail(ARAP_ERR_CUDA, "no CUDA device");
        if (o.device >= 0) {
            ARAP_CUDA(cudaSetDevice(o.device));
            device = o.device;
        } else {
            ARAP_CUDA(cudaGetDevice(&device));
        }
        ARAP_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        ARAP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        ARAP_CUDA(cudaEventCreate(&timer_start));
        ARAP_CUDA(cudaEventCreate(&timer_stop));
        ARAP_CUDA(cudaEventCreateWithFlags(&poll_event[0], cudaEventDisableTiming));
        ARAP_CUDA(cudaEventCreateWithFlags(&poll_event[1], cudaEventDisableTiming));
        ARAP_CUDA(cudaMallocHost(&cg_host, 2 * sizeof(CgScalars)));
        faces_host.assign(faces_host_in, faces_host_in + 3 * (size_t)nf);
        ARAP_CUDA(faces.ensure(3 * (size_t)nf));
        ARAP_CUDA(cudaMemcpyAsync(faces.ptr, faces_host_in, sizeof(int) * 3 * (size_t)nf, cudaMemcpyHostToDevice, stream));
        ARAP_CUDA(is_constrained.ensure((size_t)nv));
        ARAP_CUDA(target_xyz.ensure(3 * (size_t)nv));
        ARAP_CUDA(cudaMemsetAsync(is_constrained.ptr, 0, (size_t)(nv > 0 ? nv : 1), stream));
        ARAP_CUDA(cudaMemsetAsync(target_xyz.ptr, 0, sizeof(S) * 3 * (size_t)(nv > 0 ? nv : 1), stream));
        ARAP_CUDA(partials.ensure((size_t)(grid_for((size_t)nv) + 1) * 8));
        ARAP_CUDA(counter.ensure(1));
        ARAP_CUDA(cudaMemsetAsync(counter.ptr, 0, sizeof(unsigned), stream));
        ARAP_CUDA(cg.ensure(1));
        ARAP_CUDA(cudaMemsetAsync(cg.ptr, 0, sizeof(CgScalars), stream));
        ARAP_CUDA(energy_dev.ensure(1));
        ARAP_CUDA(redo_list.ensure((size_t)nv + 1));
        ARAP_CUDA(redo_count.ensure(1));
        ARAP_CUDA(redo_done.ensure(1));
        ARAP_CUDA(cudaMemsetAsync(redo_count.ptr, 0, sizeof(int), stream));
        ARAP_CUDA(cudaMemsetAsync(redo_done.ptr, 0, sizeof(unsigned), stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }

    // exclusive scan of in[0..n) into out[0..n] on the stream
    int exclusive_scan(const int *in, int n, int *out) {
        const int tiles = (n + kScanTile - 1) / kScanTile;
        if (tiles == 0) {
            ARAP_CUDA(cudaMemsetAsync(out, 0, sizeof(int), stream));
            return ARAP_OK;
        }
        ARAP_CUDA(scan_tiles.ensure((size_t)tiles + 1));
        begin_launch(ARAP_K_SCAN);
        scan_tile_sums<<<tiles, kBlock, 0, stream>>>(in, n, scan_tiles.ptr);
        scan_spine<<<1, kBlock, 0, stream>>>(scan_tiles.ptr, tiles);
        scan_apply<<<tiles, kBlock, 0, stream>>>(in, n, scan_tiles.ptr, tiles, out);
        end_launch();
        ARAP_CUDA(cudaGetLastError());
        return ARAP_OK;
    }

    // upload a host xyz array of n scalars in `scalar_bytes` precision into a device S array
    int upload_cast(const void *host, size_t n, int scalar_bytes, S *dst) {
        if (scalar_bytes == (int)sizeof(S)) {
            ARAP_CUDA(cudaMemcpyAsync(dst, host, n * sizeof(S), cudaMemcpyHostToDevice, stream));
            return ARAP_OK;
        }
        ARAP_CUDA(staging.ensure(n * (size_t)scalar_bytes));
        ARAP_CUDA(cudaMemcpyAsync(staging.ptr, host, n * (size_t)scalar_bytes, cudaMemcpyHostToDevice, stream));
        begin_launch(ARAP_K_MISC);
        if (scalar_bytes == 4) cast_xyz_kernel<S, float><<<grid_for(n), kBlock, 0, stream>>>(n, (const float *)staging.ptr, dst);
        else cast_xyz_kernel<S, double><<<grid_for(n), kBlock, 0, stream>>>(n, (const double *)staging.ptr, dst);
        end_launch();
        ARAP_CUDA(cudaGetLastError());
        return ARAP_OK;
    }

    // setConstraint overwrites the map entry of a vertex (arap.h:83): when one call names a vertex several times the LAST
    // occurrence wins. The scatter kernels write one entry per thread, so duplicates are removed here (the kept entries stay
    // in call order). Returns false when `idx` has no duplicates (the common case: nothing is copied).
    static bool last_occurrences(int n, const int *idx, std::vector<int> &keep) {
        std::vector<std::pair<int, int>> order((size_t)n);
        for (int k = 0; k < n; ++k) order[(size_t)k] = {idx[k], k};
        std::sort(order.begin(), order.end());
        bool dup = false;
        for (int k = 0; k + 1 < n && !dup; ++k) dup = order[(size_t)k].first == order[(size_t)k + 1].first;
        if (!dup) return false;
        keep.clear();
        for (int k = 0; k < n; ++k)
            if (k + 1 == n || order[(size_t)k].first != order[(size_t)k + 1].first) keep.push_back(order[(size_t)k].second);
        std::sort(keep.begin(), keep.end());
        return true;
    }

    int set_constraints(int n, const int *idx, const void *xyz, int scalar_bytes) override {
        if (n < 0 || (n > 0 && (!idx || !xyz)) || (scalar_bytes != 4 && scalar_bytes != 8))
            return fail(ARAP_ERR_INVALID, "set_constraints: bad arguments");
        for (int k = 0; k < n; ++k)
            if (idx[k] < 0 || idx[k] >= n_vertices) return fail(ARAP_ERR_INVALID, "set_constraints: vertex index out of range");
        dirty = true;                                                    // arap.h:84
        if (n == 0) return ARAP_OK;
        std::vector<int> keep, idx_unique;
        std::vector<unsigned char> xyz_unique;
        if (last_occurrences(n, idx, keep)) {
            const size_t eb = 3 * (size_t)scalar_bytes;
            idx_unique.resize(keep.size());
            xyz_unique.resize(keep.size() * eb);
            for (size_t q = 0; q < keep.size(); ++q) {
                idx_unique[q] = idx[keep[q]];
                std::memcpy(&xyz_unique[q * eb], (const unsigned char *)xyz + (size_t)keep[q] * eb, eb);
            }
            n = (int)keep.size();
            idx = idx_unique.data();
            xyz = xyz_unique.data();
        }
        const size_t idx_bytes = sizeof(int) * (size_t)n, xyz_bytes = (size_t)scalar_bytes * 3 * (size_t)n;
        const size_t xyz_off = (idx_bytes + 15) & ~(size_t)15;
        ARAP_CUDA(staging.ensure(xyz_off + xyz_bytes));
        ARAP_CUDA(cudaMemcpyAsync(staging.ptr, idx, idx_bytes, cudaMemcpyHostToDevice, stream));
        ARAP_CUDA(cudaMemcpyAsync(staging.ptr + xyz_off, xyz, xyz_bytes, cudaMemcpyHostToDevice, stream));
        begin_launch(ARAP_K_MISC);
        if (scalar_bytes == 4)
            set_constraints_kernel<S, float><<<grid_for((size_t)n), kBlock, 0, stream>>>(
                n, (const int *)staging.ptr, (const float *)(staging.ptr + xyz_off), n_vertices, is_constrained.ptr, target_xyz.ptr);
        else
            set_constraints_kernel<S, double><<<grid_for((size_t)n), kBlock, 0, stream>>>(
                n, (const int *)staging.ptr, (const double *)(staging.ptr + xyz_off), n_vertices, is_constrained.ptr, target_xyz.ptr);
        end_launch();
        ARAP_CUDA(cudaGetLastError());
        ARAP_CUDA(cudaStreamSynchronize(stream));   // host buffers are caller-owned: finish the copies before returning
        return ARAP_OK;
    }

    int set_rigid_constraints(int n, int batch, int member_stride, const int *idx, const void *rest, int scalar_bytes,
                              const double *transforms16) override {
        if (n < 0 || batch <= 0 || (n > 0 && (!idx || !rest || !transforms16)) || (scalar_bytes != 4 && scalar_bytes != 8))
            return fail(ARAP_ERR_INVALID, "set_rigid_constraints: bad arguments");
        for (int k = 0; k < n; ++k)
            if (idx[k] < 0 || idx[k] >= member_stride || (long long)idx[k] + (long long)(batch - 1) * member_stride >= n_vertices)
                return fail(ARAP_ERR_INVALID, "set_rigid_constraints: vertex index out of range");
        dirty = true;                                                    // arap.h:84
        if (n == 0) return ARAP_OK;
        std::vector<int> keep, idx_unique;
        std::vector<unsigned char> rest_unique;
        if (last_occurrences(n, idx, keep)) {                            // the same handle listed twice: the last entry wins
            const size_t eb = 3 * (size_t)scalar_bytes;
            idx_unique.resize(keep.size());
            rest_unique.resize(keep.size() * eb);
            for (size_t q = 0; q < keep.size(); ++q) {
                idx_unique[q] = idx[keep[q]];
                std::memcpy(&rest_unique[q * eb], (const unsigned char *)rest + (size_t)keep[q] * eb, eb);
            }
            n = (int)keep.size();
            idx = idx_unique.data();
            rest = rest_unique.data();
        }
        std::vector<double> rows(12 * (size_t)batch);                    // top three rows of each 4x4
        for (int m = 0; m < batch; ++m) std::memcpy(&rows[12 * (size_t)m], transforms16 + 16 * (size_t)m, 12 * sizeof(double));
        const size_t idx_bytes = sizeof(int) * (size_t)n, xyz_bytes = (size_t)scalar_bytes * 3 * (size_t)n, tr_bytes = rows.size() * sizeof(double);
        const size_t xyz_off = (idx_bytes + 15) & ~(size_t)15, tr_off = (xyz_off + xyz_bytes + 15) & ~(size_t)15;
        ARAP_CUDA(staging.ensure(tr_off + tr_bytes));
        ARAP_CUDA(cudaMemcpyAsync(staging.ptr, idx, idx_bytes, cudaMemcpyHostToDevice, stream));
        ARAP_CUDA(cudaMemcpyAsync(staging.ptr + xyz_off, rest, xyz_bytes, cudaMemcpyHostToDevice, stream));
        ARAP_CUDA(cudaMemcpyAsync(staging.ptr + tr_off, rows.data(), tr_bytes, cudaMemcpyHostToDevice, stream));
        begin_launch(ARAP_K_MISC);
        const unsigned grid = grid_for((size_t)n * (size_t)batch);
        if (scalar_bytes == 4)
            set_rigid_constraints_kernel<S, float><<<grid, kBlock, 0, stream>>>(
                n, batch, member_stride, (const int *)staging.ptr, (const float *)(staging.ptr + xyz_off),
                (const double *)(staging.ptr + tr_off), n_vertices, is_constrained.ptr, target_xyz.ptr);
        else
            set_rigid_constraints_kernel<S, double><<<grid, kBlock, 0, stream>>>(
                n, batch, member_stride, (const int *)staging.ptr, (const double *)(staging.ptr + xyz_off),
                (const double *)(staging.ptr + tr_off), n_vertices, is_constrained.ptr, target_xyz.ptr);
        end_launch();
        ARAP_CUDA(cudaGetLastError());
        ARAP_CUDA(cudaStreamSynchronize(stream));   // host buffers (and `rows`) must outlive the copies
        return ARAP_OK;
    }

    int prepare(const void *rest_host, int scalar_bytes) override {
        if (!rest_host || (scalar_bytes != 4 && scalar_bytes != 8)) return fail(ARAP_ERR_INVALID, "prepare: bad arguments");
        const int V = n_vertices, F = n_faces;
        prepared = false;
        std::memset(&stats, 0, sizeof(stats));
        comm_counted = false;
        // ---- initializeMeshGeometry (arap.h:162-168)
        ARAP_CUDA(rest_xyz.ensure(3 * (size_t)V));
        { int rc = upload_cast(rest_host, 3 * (size_t)V, scalar_bytes, rest_xyz.ptr); if (rc) return rc; }
        // ---- computeCotanWeights (arap.h:182-239)
        ARAP_CUDA(row_count.ensure((size_t)V + 1));
        ARAP_CUDA(raw_rowptr.ensure((size_t)V + 1));
        ARAP_CUDA(row_cursor.ensure((size_t)V + 1));
        ARAP_CUDA(unique_count.ensure((size_t)V + 1));
        ARAP_CUDA(rowptr.ensure((size_t)V + 1));
        ARAP_CUDA(raw_col.ensure(6 * (size_t)F));
        ARAP_CUDA(raw_val.ensure(6 * (size_t)F));
        ARAP_CUDA(raw_tag.ensure(6 * (size_t)F));
        ARAP_CUDA(cudaMemsetAsync(row_count.ptr, 0, sizeof(int) * ((size_t)V + 1), stream));
        ARAP_CUDA(cudaMemsetAsync(row_cursor.ptr, 0, sizeof(int) * ((size_t)V + 1), stream));
        if (F > 0) LAUNCH(ARAP_K_WEIGHTS_COUNT, weights_count_kernel, grid_for((size_t)F), faces.ptr, F, row_count.ptr);
        { int rc = exclusive_scan(row_count.ptr, V, raw_rowptr.ptr); if (rc) return rc; }
        if (F > 0)
            LAUNCH(ARAP_K_WEIGHTS_FILL, weights_fill_kernel<S>, grid_for((size_t)F), faces.ptr, F, rest_xyz.ptr, raw_rowptr.ptr,
                   row_cursor.ptr, raw_col.ptr, raw_val.ptr, raw_tag.ptr);
        if (V > 0) {
            // rows of more than kLongRow raw triplets (very high valence) are listed in row_cursor -- free again after
            // weights_fill -- with the count in its last slot, and sorted by one CTA each
            ARAP_CUDA(cudaMemsetAsync(row_cursor.ptr + V, 0, sizeof(int), stream));
            LAUNCH(ARAP_K_ROW_SORT_MERGE, row_sort_merge_kernel<S>, grid_for((size_t)V), V, raw_rowptr.ptr, raw_col.ptr, raw_val.ptr,
                   raw_tag.ptr, unique_count.ptr, row_cursor.ptr, row_cursor.ptr + V);
            LAUNCH(ARAP_K_ROW_SORT_MERGE, row_sort_long_kernel<S>, sm_count, row_cursor.ptr, row_cursor.ptr + V, raw_rowptr.ptr, raw_col.ptr,
                   raw_val.ptr, raw_tag.ptr, unique_count.ptr);
        }
        { int rc = exclusive_scan(unique_count.ptr, V, rowptr.ptr); if (rc) return rc; }
        ARAP_CUDA(cudaMemcpyAsync(&nnz, rowptr.ptr + V, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(colidx.ensure((size_t)nnz));
        ARAP_CUDA(weight.ensure((size_t)nnz));
        if (V > 0)
            LAUNCH(ARAP_K_CSR_COMPACT, csr_compact_kernel<S>, grid_for((size_t)V), V, raw_rowptr.ptr, raw_col.ptr, raw_val.ptr,
                   rowptr.ptr, colidx.ptr, weight.ptr);
        // ---- initializeFreeVariableMapping (arap.h:261-272)
        ARAP_CUDA(flags.ensure((size_t)V + 1));
        ARAP_CUDA(free_idx.ensure((size_t)V + 1));
        if (V > 0) LAUNCH(ARAP_K_MISC, free_flag_kernel, grid_for((size_t)V), V, is_constrained.ptr, flags.ptr);
        { int rc = exclusive_scan(flags.ptr, V, row_count.ptr); if (rc) return rc; }   // row_count reused as the prefix array
        ARAP_CUDA(cudaMemcpyAsync(&n_free, row_count.ptr + V, sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (V > 0) LAUNCH(ARAP_K_MISC, free_map_kernel, grid_for((size_t)V), V, is_constrained.ptr, row_count.ptr, free_idx.ptr);
        // ---- internal vertex order (once per handle) and the hot CSR in that order (every prepare: the weights changed)
        n_rows = transport ? plan.n_owned : V;
        if (!have_perm) { int rc = build_permutation(rest_host, scalar_bytes); if (rc) return rc; }
        ARAP_CUDA(hot_rowptr.ensure((size_t)V + 1));
        ARAP_CUDA(hot_colidx.ensure((size_t)nnz + 4));      // + 4: TMA bulk copies round a tile's span up to 16 bytes
        ARAP_CUDA(hot_weight.ensure((size_t)nnz + 4));
        ARAP_CUDA(hot_weight_f32.ensure((size_t)nnz + 4));
        ARAP_CUDA(free_mask.ensure((size_t)V + 1));
        if (V > 0) LAUNCH(ARAP_K_MISC, perm_row_count_kernel, grid_for((size_t)V), V, perm.ptr, rowptr.ptr, unique_count.ptr);
        { int rc = exclusive_scan(unique_count.ptr, V, hot_rowptr.ptr); if (rc) return rc; }
        if (V > 0)
            LAUNCH(ARAP_K_MISC, perm_csr_fill_kernel<S>, grid_for((size_t)V), V, perm.ptr, iperm.ptr, rowptr.ptr, colidx.ptr, weight.ptr,
                   hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, hot_weight_f32.ptr);
        // ---- tiles of the hot CSR (topology and order are fixed per handle: built once)
        if (!tiles_built) { int rc = build_tiles(); if (rc) return rc; }
        // ---- initializeMeshGeometry / Rotations / Constraints into the solver layout (arap.h:162-168,246-249,277-281)
        ARAP_CUDA(rest4.ensure((size_t)V));
        ARAP_CUDA(cur4.ensure((size_t)V));
        ARAP_CUDA(quat.ensure((size_t)V));
        ARAP_CUDA(inv_diag.ensure((size_t)V));
        if (V > 0)
            LAUNCH(ARAP_K_INIT_STATE, init_state_kernel<S>, grid_for((size_t)V), V, perm.ptr, rest_xyz.ptr, is_constrained.ptr, target_xyz.ptr,
                   hot_rowptr.ptr, hot_weight.ptr, rest4.ptr, cur4.ptr, quat.ptr, inv_diag.ptr, free_mask.ptr);
        ARAP_CUDA(cudaGetLastError());
        ARAP_CUDA(cudaStreamSynchronize(stream));
        if (n_free == V && !transport) return ARAP_UNCONSTRAINED;        // arap.h:113-114: stays dirty, nothing solved
        // ---- setupLinearSystem (arap.h:292-340): matrix-free L, nothing to factor; allocate the CG vectors
        ARAP_CUDA(cg_r.ensure((size_t)V));
        ARAP_CUDA(cg_d.ensure((size_t)V));
        ARAP_CUDA(cg_ad.ensure((size_t)V));
        ARAP_CUDA(cg_x.ensure((size_t)V));
        ARAP_CUDA(cg_w.ensure((size_t)V));
        for (Vec3d *vec : {cg_r.ptr, cg_d.ptr, cg_ad.ptr, cg_x.ptr, cg_w.ptr})                                  // halo slots must start finite
            ARAP_CUDA(cudaMemsetAsync(vec, 0, sizeof(Vec3d) * (size_t)(V > 0 ? V : 1), stream));
        use_mg = (opt.solver != ARAP_SOLVER_PCG_JACOBI);
        if (use_mg) {
            // The hierarchy is only a preconditioner: the operator CG sees is always the exact, freshly weighted one.
            // The reference's dirty protocol re-weights the mesh on every constraint change (every frame in its demos,
            // arap.h:84,102-120), but as long as the SET of constrained vertices is the same the old hierarchy still
            // preconditions the new system well, so it is kept (saves the host setup, ~0.6 s at 1M vertices) until the
            // CG needs noticeably more iterations than right after the last fresh setup.
            std::vector<unsigned char> mask((size_t)V);      // user order is fine here: it is only compared with the previous one
            if (V > 0) ARAP_CUDA(cudaMemcpyAsync(mask.data(), is_constrained.ptr, (size_t)V, cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaStreamSynchronize(stream));
            // Partitioned mode with a global hierarchy: every rank must take the same decision, so only facts all ranks
            // share may enter it (the global constrained set, and mg_stale, which derives from all-reduced CG counts).
            const bool want_global = transport && have_global && getenv("ARAP_MG_BLOCK_JACOBI") == nullptr;
            const bool same_problem = want_global ? (mg_global && global_mesh.constrained == mg_global_mask) : (!mg_global && mask == mg_mask && nnz == mg_nnz);
            const bool reusable = !mg.empty() && !mg_stale && same_problem && getenv("ARAP_MG_ALWAYS_REBUILD") == nullptr;
            if (reusable) {
                stats.mg_levels = (int)mg.size();
                stats.mg_operator_complexity = mg_complexity;
                stats.setup_host_ms = 0.0;
                mg_fresh = false;
            } else {
                mg_global = false;
                if (want_global) {
                    int rc = setup_multigrid_global();
                    if (rc) return rc;
                    mg_global_mask = global_mesh.constrained;
                }
                if (!mg_global) { int rc = setup_multigrid(); if (rc) return rc; }
                mg_mask.swap(mask);
                mg_nnz = nnz;
                mg_complexity = stats.mg_operator_complexity;
                mg_stale = false;
                mg_fresh = true;
                mg_fresh_iterations = 0;
            }
        }
        CgScalars init;
        std::memset(&init, 0, sizeof(init));
        // default tolerance: with the multigrid preconditioner the residual tracks the error closely (1e-6 keeps
        // positions within ~2e-8 x bbox diagonal of a direct solve, measured); plain Jacobi needs a much smaller one.
        // Stopping rule. Jacobi-PCG: relative residual (default 1e-9). Multigrid: the estimated position error of the
        // iterate (kernels.cuh, CG_STAGE_RHO) against position_tolerance x bbox diagonal -- the parity bar is a position
        // bar, and the same relative residual means very different position errors on different meshes. An explicit
        // cg_tolerance keeps the residual rule (and switches the position rule off unless that is given explicitly too).
        const double tol = opt.cg_tolerance > 0 ? opt.cg_tolerance : (use_mg ? 1e-13 : 1e-9);
        init.tol2 = tol * tol;
        if (use_mg && (opt.position_tolerance > 0 || !(opt.cg_tolerance > 0))) {
            const double ptol = opt.position_tolerance > 0 ? opt.position_tolerance : kDefaultPositionTolerance;
            init.z8_tol = std::pow(ptol, 8.0);
            // a length SCALE: measured on the first prepare of the handle and kept (later rest poses are deformed states of
            // the same mesh; walking 3V host values again would cost the per-frame dirty cycle milliseconds at 1M vertices)
            if (!(length_scale > 0)) length_scale = rest_bbox_diagonal(rest_host, scalar_bytes);
            init.inv_len2 = length_scale > 0 ? 1.0 / (length_scale * length_scale) : 1.0;
        }
        init.distributed = transport ? 1 : 0;
        init.max_iterations = opt.max_cg_iterations > 0 ? opt.max_cg_iterations : 20000;
        cg_host[0] = init;
        ARAP_CUDA(cudaMemcpyAsync(cg.ptr, &cg_host[0], sizeof(CgScalars), cudaMemcpyHostToDevice, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        if (transport) { int rc = configure_transport(); if (rc) return rc; }
        destroy_step_graph();
        if (!transport) {
            int rc = build_cg_graph();
            if (rc && tail_first > 0) {           // a cooperative / cluster launch that cannot be captured here: one launch per phase instead
                cudaGetLastError();
                tail_first = 0;
                tail_grid = false;
                rc = build_cg_graph();
            }
            if (rc) return rc;
            { int rc2 = build_step_graph(); if (rc2) return rc2; }
        }
        else if (transport->capturable()) {
            // NCCL calls and the peer transport's kernels are stream operations and can sit inside the graph, which removes
            // ~70 host-side enqueues per CG iteration. NCCL opens connections on first use, which must not happen during
            // capture: touch every exchange plan and both all-reduce flavours once, eagerly, with the (still zero) vectors.
            if (transport->needs_warm_up()) { int rc = warm_up_transport(); if (rc) return rc; }
            if (build_cg_graph() != ARAP_OK) {                            // every rank fails alike: plain enqueueing instead
                destroy_cg_graph();
                cudaGetLastError();
            }
        } else destroy_cg_graph();                                        // in-process transport synchronises on the host
        // Partitions that share one GPU and wait on each other inside kernels (in-process peer transport) must not start
        // iterating while another one is still allocating: cudaFree waits for ALL device work, spinning kernels included.
        if (transport && transport->barrier(stream)) return fail(ARAP_ERR_CUDA, transport->error);
        stats.cg_graph = step_graph_exec ? 2 : (cg_graph_exec ? 1 : 0);
        stats.tile_max_halo = use_tiles ? tile_max_halo : 0;
        stats.launches_per_cg_iteration = 0;
        for (int k = 0; k < ARAP_K_COUNT_MAX; ++k) stats.launches_per_cg_iteration += (int)(step_graph_exec ? body_counts[k] : graph_counts_iter[k]);
        stats.renumbered = renumbered ? 1 : 0;
        pdl_next_plain = true;
        stats.mg_global = (use_mg && mg_global) ? 1 : 0;
        have_warm_rotations = false;                                     // initializeRotations (arap.h:246-249)
        dirty = false;                                                   // arap.h:119
        prepared = true;
        return ARAP_OK;
    }

    // Tile structure for the staged one-ring kernels (tile_kernels.cuh); falls back to the untiled kernels (use_tiles = false)
    // when a tile's halo does not fit, for tiny meshes, for batches (their members are far smaller than the halo capacity
    // would allow... a member's rows all see each other) and with ARAP_TILES=0.
    int build_tiles() {
        tiles_built = true;
        use_tiles = false;
        const int R = n_rows;
        // OFF by default: measured at 1M vertices (profiles/r02_experiments.txt) the staged fp64 kernels are 24-31 % SLOWER than the
        // direct gathers (CTA-wide stage / barrier / compute phases overlap worse than independent warps do), the fp32 ones equal.
        if (!(getenv("ARAP_TILES") && atoi(getenv("ARAP_TILES")) != 0)) return ARAP_OK;
        if (R < 2 * kTile || nnz <= 0) return ARAP_OK;
        n_tiles = (R + kTile - 1) / kTile;
        ARAP_CUDA(tile_halo.ensure((size_t)n_tiles * kTileHaloCap));
        ARAP_CUDA(tile_halo_count.ensure((size_t)n_tiles));
        ARAP_CUDA(tile_colidx.ensure((size_t)nnz));
        ARAP_CUDA(tile_scalars.ensure(2));
        ARAP_CUDA(cudaMemsetAsync(tile_scalars.ptr, 0, 2 * sizeof(int), stream));
        LAUNCH(ARAP_K_MISC, build_tiles_kernel, n_tiles, R, hot_rowptr.ptr, hot_colidx.ptr, tile_halo.ptr, tile_halo_count.ptr, tile_colidx.ptr,
               tile_scalars.ptr, tile_scalars.ptr + 1);
        int h[2] = {0, 0};
        ARAP_CUDA(cudaMemcpyAsync(h, tile_scalars.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        if (h[1] != 0) return ARAP_OK;                                  // a tile did not fit: untiled kernels
        tile_max_halo = (h[0] + 31) & ~31;
        // Dynamic shared memory of the staged kernels: (kTile + largest halo of THIS handle) records per staged array. The opt-in
        // limit is a property of the kernel function, shared by every handle of the process (several partitions of one mesh
        // may live in one process): always raise it to what the largest admissible halo needs, never to this handle's own size.
        const size_t cap_records = (size_t)kTile + kTileHaloCap;
        ARAP_CUDA(cudaFuncSetAttribute(local_step_tiled_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap_records * sizeof(Vec4T<S>) * 2)));
        ARAP_CUDA(cudaFuncSetAttribute(rhs_residual_tiled_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap_records * sizeof(Vec4T<S>) * 3)));
        ARAP_CUDA(cudaFuncSetAttribute(rhs_residual_tiled_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap_records * sizeof(Vec4T<S>) * 3)));
        ARAP_CUDA(cudaFuncSetAttribute(mg_fine_residual_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap_records * sizeof(MgVec))));
        ARAP_CUDA(cudaFuncSetAttribute(mg_fine_postsmooth_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap_records * sizeof(MgVec))));
        ARAP_CUDA(cudaFuncSetAttribute(cg_spmv_z_tiled_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap_records * sizeof(MgVec))));
        use_tiles = true;
        return ARAP_OK;
    }

    // ---- launches of the one-ring kernels: staged through shared memory when the tile structure exists ------------------
    void launch_local_step() {
        const int R = n_rows;
        if (use_tiles) {
            const size_t sm = tile_smem(sizeof(Vec4T<S>), 2);
            LAUNCH_PDL_SMEM(ARAP_K_LOCAL_STEP, local_step_tiled_kernel<S>, std::min(n_tiles, reduce_grid(local_step_tiled_kernel<S>, (size_t)R, sm)), sm, R, tile_view(),
                            (const int *)hot_rowptr.ptr, (const S *)hot_weight.ptr, (const Vec4T<S> *)rest4.ptr, (const Vec4T<S> *)cur4.ptr, quat.ptr,
                            redo_list.ptr, redo_count.ptr);
        } else {
            LAUNCH_PDL(ARAP_K_LOCAL_STEP, local_step_kernel<S>, grid_for((size_t)R), R, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, rest4.ptr, cur4.ptr, quat.ptr,
                       redo_list.ptr, redo_count.ptr);
        }
        LAUNCH_PDL(ARAP_K_LOCAL_STEP_REDO, local_step_redo_kernel<S>, sm_count * 2, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, rest4.ptr, cur4.ptr,
                   quat.ptr, redo_list.ptr, redo_count.ptr, redo_done.ptr);
    }
    template <bool MG>
    void launch_rhs_residual(double omega0, float4 *x0) {
        const int R = n_rows;
        if (use_tiles) {
            const size_t sm = tile_smem(sizeof(Vec4T<S>), 3);
            LAUNCH_PDL_SMEM(ARAP_K_RHS_RESIDUAL, (rhs_residual_tiled_kernel<S, MG>), std::min(n_tiles, reduce_grid(rhs_residual_tiled_kernel<S, MG>, (size_t)R, sm)), sm, R,
                            tile_view(), (const int *)hot_rowptr.ptr, (const S *)hot_weight.ptr, (const Vec4T<S> *)rest4.ptr, (const Vec4T<S> *)cur4.ptr,
                            (const Vec4T<S> *)quat.ptr, (const double *)inv_diag.ptr, omega0, cg_r.ptr, cg_d.ptr, cg_x.ptr, x0, partials.ptr, counter.ptr, cg.ptr);
        } else {
            LAUNCH_PDL(ARAP_K_RHS_RESIDUAL, (rhs_residual_kernel<S, MG>), reduce_grid(rhs_residual_kernel<S, MG>, (size_t)R), R, hot_rowptr.ptr, hot_colidx.ptr,
                       hot_weight.ptr, rest4.ptr, cur4.ptr, quat.ptr, inv_diag.ptr, omega0, cg_r.ptr, cg_d.ptr, cg_x.ptr, x0, partials.ptr, counter.ptr, cg.ptr);
        }
    }
    void launch_fine_residual() {
        const int R = n_rows;
        MgLevelDev &m0 = *mg[0];
        if (use_tiles) {
            const size_t sm = tile_smem(sizeof(MgVec), 1);
            LAUNCH_PDL_SMEM(ARAP_K_MG_FINE_RESIDUAL, mg_fine_residual_tiled_kernel, std::min(n_tiles, reduce_grid(mg_fine_residual_tiled_kernel, (size_t)R, sm)), sm, R,
                            tile_view(), (const int *)hot_rowptr.ptr, (const float *)hot_weight_f32.ptr, (const unsigned char *)free_mask.ptr,
                            (const Vec3d *)cg_r.ptr, (const MgVec *)m0.x.ptr, m0.r.ptr, (const CgScalars *)cg.ptr);
        } else {
            LAUNCH_PDL(ARAP_K_MG_FINE_RESIDUAL, mg_fine_residual_kernel, grid_for((size_t)R), R, hot_rowptr.ptr, hot_colidx.ptr, hot_weight_f32.ptr,
                       free_mask.ptr, cg_r.ptr, m0.x.ptr, m0.r.ptr, cg.ptr);
        }
    }
    void launch_fine_postsmooth() {
        const int R = n_rows;
        MgLevelDev &m0 = *mg[0];
        if (use_tiles) {
            const size_t sm = tile_smem(sizeof(MgVec), 1);
            LAUNCH_PDL_SMEM(ARAP_K_MG_FINE_POSTSMOOTH, mg_fine_postsmooth_tiled_kernel, std::min(n_tiles, reduce_grid(mg_fine_postsmooth_tiled_kernel, (size_t)R, sm)), sm,
                            R, tile_view(), (const int *)hot_rowptr.ptr, (const float *)hot_weight_f32.ptr, (const unsigned char *)free_mask.ptr,
                            (const double *)inv_diag.ptr, m0.omega, (const Vec3d *)cg_r.ptr, (const MgVec *)m0.x.ptr, m0.x2.ptr, partials.ptr, counter.ptr, cg.ptr);
        } else {
            LAUNCH_PDL(ARAP_K_MG_FINE_POSTSMOOTH, mg_fine_postsmooth_kernel, reduce_grid(mg_fine_postsmooth_kernel, (size_t)R), R, hot_rowptr.ptr,
                       hot_colidx.ptr, hot_weight_f32.ptr, free_mask.ptr, inv_diag.ptr, m0.omega, cg_r.ptr, m0.x.ptr, m0.x2.ptr, partials.ptr, counter.ptr, cg.ptr);
        }
    }
    void launch_spmv_z() {
        const int R = n_rows;
        MgLevelDev &m0 = *mg[0];
        if (use_tiles) {
            const size_t sm = tile_smem(sizeof(MgVec), 1);
            LAUNCH_PDL_SMEM(ARAP_K_CG_SPMV, cg_spmv_z_tiled_kernel<S>, std::min(n_tiles, reduce_grid(cg_spmv_z_tiled_kernel<S>, (size_t)R, sm)), sm, R, tile_view(),
                            (const int *)hot_rowptr.ptr, (const S *)hot_weight.ptr, (const unsigned char *)free_mask.ptr, (const float4 *)m0.x2.ptr, cg_w.ptr,
                            partials.ptr, counter.ptr, cg.ptr);
        } else {
            LAUNCH_PDL(ARAP_K_CG_SPMV, cg_spmv_z_kernel<S>, reduce_grid(cg_spmv_z_kernel<S>, (size_t)R), R, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr,
                       free_mask.ptr, (const float4 *)m0.x2.ptr, cg_w.ptr, partials.ptr, counter.ptr, cg.ptr);
        }
    }

    // bounding-box diagonal of the rest pose: the length scale of the position-error stopping rule. Partitioned mode with
    // the global mesh known: the GLOBAL box (every rank must scale alike); otherwise this rank's box (smaller: conservative).
    double rest_bbox_diagonal(const void *rest_host, int scalar_bytes) const {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        auto add = [&](double x, int d) { if (x < lo[d]) lo[d] = x; if (x > hi[d]) hi[d] = x; };
        if (transport && have_global) {
            for (size_t k = 0; k < global_mesh.rest.size(); ++k) add(global_mesh.rest[k], (int)(k % 3));
        } else {
            const size_t n3 = 3 * (size_t)n_vertices;
            if (scalar_bytes == 4) for (size_t k = 0; k < n3; ++k) add((double)((const float *)rest_host)[k], (int)(k % 3));
            else for (size_t k = 0; k < n3; ++k) add(((const double *)rest_host)[k], (int)(k % 3));
        }
        double s = 0;
        for (int d = 0; d < 3; ++d) if (hi[d] > lo[d]) s += (hi[d] - lo[d]) * (hi[d] - lo[d]);
        return std::sqrt(s);
    }

    // ---- internal vertex order --------------------------------------------------------------------------------------
    // Morton (Z-curve) order of the rest pose: vertices that are close in space -- a vertex and its one-ring -- get close
    // internal indices, so the gathers of a CTA mostly hit lines its own rows already pulled into L1. Computed once per
    // handle on the host (the topology is fixed, arap.h:95; the geometry of later dirty cycles stays close). In
    // partitioned mode only the owned block is reordered; the halo keeps the layout the exchange plan relies on.
    int build_permutation(const void *rest_host, int scalar_bytes) {
        const int V = n_vertices, owned = n_rows;
        std::vector<int> h_perm((size_t)V);
        for (int i = 0; i < V; ++i) h_perm[(size_t)i] = i;
        const char *env = getenv("ARAP_REORDER");
        if (!(env && atoi(env) == 0) && owned > 1 && batch_members <= 1) {     // batches keep their member-major order
            auto coord = [&](int v, int d) -> double {
                return scalar_bytes == 4 ? (double)((const float *)rest_host)[3 * (size_t)v + d] : ((const double *)rest_host)[3 * (size_t)v + d];
            };
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int v = 0; v < owned; ++v)
                for (int d = 0; d < 3; ++d) { const double c = coord(v, d); if (c < lo[d]) lo[d] = c; if (c > hi[d]) hi[d] = c; }
            double extent = 0;
            for (int d = 0; d < 3; ++d) extent = std::max(extent, hi[d] - lo[d]);
            const double scale = extent > 0 ? 2097151.0 / extent : 0.0;             // 21 bits per axis
            auto spread = [](uint64_t x) -> uint64_t {                                // insert two zero bits between the 21 low bits
                x &= 0x1fffffULL;
                x = (x | x << 32) & 0x1f00000000ffffULL;
                x = (x | x << 16) & 0x1f0000ff0000ffULL;
                x = (x | x << 8) & 0x100f00f00f00f00fULL;
                x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
                x = (x | x << 2) & 0x1249249249249249ULL;
                return x;
            };
            std::vector<std::pair<uint64_t, int>> keyed((size_t)owned);
            for (int v = 0; v < owned; ++v) {
                uint64_t key = 0;
                for (int d = 0; d < 3; ++d) key |= spread((uint64_t)((coord(v, d) - lo[d]) * scale)) << d;
                keyed[(size_t)v] = {key, v};
            }
            std::sort(keyed.begin(), keyed.end());
            // Patches, not a space-filling curve: consecutive chunks of the Morton sequence are compact blobs of vertices
            // (what one CTA should work on, so that most of its one-ring gathers hit lines it fetched itself), but INSIDE a
            // chunk the user's own order is kept: meshes are usually locally row-coherent, and a warp whose 32 lanes walk a
            // row gathers 32 contiguous neighbours -- a Z-curve inside the patch would scatter them.
            {
                const char *penv = getenv("ARAP_PATCH");
                const int patch = penv ? atoi(penv) : kBlock;
                if (patch > 1)
                    for (int b = 0; b < owned; b += patch) {
                        const int e = std::min(owned, b + patch);
                        std::sort(keyed.begin() + b, keyed.begin() + e,
                                  [](const std::pair<uint64_t, int> &x, const std::pair<uint64_t, int> &y) { return x.second < y.second; });
                    }
            }
            // Renumber the memory layout only if it buys locality: count the mesh edges whose end points are within a
            // few CTAs of each other, in the user's order and in Morton order. Generated / scanned-in-strips meshes are
            // often already well ordered (the headline icosphere is: renumbering it made the gather kernels 15 % slower).
            std::vector<int> rank_of((size_t)V);
            for (int i = 0; i < V; ++i) rank_of[(size_t)i] = i;
            for (int i = 0; i < owned; ++i) rank_of[(size_t)keyed[(size_t)i].second] = i;
            long long near_user = 0, near_morton = 0;
            const int window = 4 * kBlock;
            for (size_t f = 0; f + 2 < faces_host.size(); f += 3)
                for (int e = 0; e < 3; ++e) {
                    const int a = faces_host[f + e], b = faces_host[f + (e + 1) % 3];
                    if (a >= owned || b >= owned) continue;
                    if (std::abs(a - b) <= window) ++near_user;
                    if (std::abs(rank_of[(size_t)a] - rank_of[(size_t)b]) <= window) ++near_morton;
                }
            // With the tiled kernels (tile_kernels.cuh) what matters is how many mesh edges leave their tile of kTile consecutive
            // rows: every such edge puts a vertex on the tile's halo list. Count them in both orders.
            long long cut_user = 0, cut_morton = 0;
            for (size_t f = 0; f + 2 < faces_host.size(); f += 3)
                for (int e = 0; e < 3; ++e) {
                    const int a = faces_host[f + e], b = faces_host[f + (e + 1) % 3];
                    if (a >= owned || b >= owned) continue;
                    if (a / kTile != b / kTile) ++cut_user;
                    if (rank_of[(size_t)a] / kTile != rank_of[(size_t)b] / kTile) ++cut_morton;
                }
            const bool force = env && atoi(env) == 2;
            const bool tiling = getenv("ARAP_TILES") && atoi(getenv("ARAP_TILES")) != 0 && owned >= 2 * kTile;
            const bool renumber = force || (tiling ? (double)cut_morton < 0.8 * (double)cut_user : (double)near_morton > 1.15 * (double)near_user);
            mg_visit_order.resize((size_t)V);
            for (int i = 0; i < V; ++i) mg_visit_order[(size_t)i] = i;
            renumbered = renumber;
            if (renumber) {
                for (int i = 0; i < owned; ++i) h_perm[(size_t)i] = keyed[(size_t)i].second;       // internal order IS the Morton order
            } else {
                for (int i = 0; i < owned; ++i) mg_visit_order[(size_t)i] = keyed[(size_t)i].second; // identity layout, Morton visiting order
            }
        }
        std::vector<int>().swap(faces_host);
        ARAP_CUDA(upload_vector(perm, h_perm, stream));
        ARAP_CUDA(iperm.ensure((size_t)V));
        if (V > 0) LAUNCH(ARAP_K_MISC, invert_perm_kernel, grid_for((size_t)V), V, perm.ptr, iperm.ptr);
        if (transport && plan.n_send() > 0) {                  // the exchange plan speaks user-local indices
            std::vector<int> h_iperm((size_t)V);
            for (int i = 0; i < V; ++i) h_iperm[(size_t)h_perm[(size_t)i]] = i;
            std::vector<int> send((size_t)plan.n_send());
            for (size_t k = 0; k < send.size(); ++k) send[k] = h_iperm[(size_t)plan.send_index[k]];
            ARAP_CUDA(upload_vector(send_index_dev, send, stream));
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));              // h_perm dies at scope exit
        have_perm = true;
        return ARAP_OK;
    }

    // ---- partitioned mode ----------------------------------------------------------------------------------------
    int attach_partition(const arap_partition_plan *p, int rank, int world, int kind, const void *id, int id_bytes) override {
        if (!p || p->n_owned < 0 || p->n_owned > n_vertices || p->n_neighbors < 0 || world <= 0 || rank < 0 || rank >= world)
            return fail(ARAP_ERR_INVALID, "attach_partition: bad arguments");
        if (!p->send_offset || !p->recv_offset || (p->n_neighbors > 0 && !p->neighbor_rank))
            return fail(ARAP_ERR_INVALID, "attach_partition: null plan arrays");
        for (int k = 0; k < p->n_neighbors; ++k) {
            if (p->neighbor_rank[k] < 0 || p->neighbor_rank[k] >= world || p->neighbor_rank[k] == rank)
                return fail(ARAP_ERR_INVALID, "attach_partition: neighbour rank out of range");
            if (p->send_offset[k] < 0 || p->send_offset[k + 1] < p->send_offset[k] || p->recv_offset[k] < 0 || p->recv_offset[k + 1] < p->recv_offset[k])
                return fail(ARAP_ERR_INVALID, "attach_partition: offsets must start at >= 0 and ascend");
        }
        if (p->send_offset[0] != 0 || p->recv_offset[0] != 0) return fail(ARAP_ERR_INVALID, "attach_partition: offsets must start at 0");
        if (p->send_offset[p->n_neighbors] > 0 && !p->send_index) return fail(ARAP_ERR_INVALID, "attach_partition: null send_index");
        my_rank = rank;
        world_size = world;
        plan = HaloPlan();
        plan.n_owned = p->n_owned;
        plan.neighbor_rank.assign(p->neighbor_rank, p->neighbor_rank + p->n_neighbors);
        plan.send_offset.assign(p->send_offset, p->send_offset + p->n_neighbors + 1);
        plan.recv_offset.assign(p->recv_offset, p->recv_offset + p->n_neighbors + 1);
        plan.send_index.assign(p->send_index, p->send_index + plan.send_offset.back());
        if (plan.n_owned + plan.n_halo() != n_vertices) return fail(ARAP_ERR_INVALID, "attach_partition: n_owned + halo != local vertex count");
        for (int v : plan.send_index) if (v < 0 || v >= plan.n_owned) return fail(ARAP_ERR_INVALID, "attach_partition: send index is not an owned vertex");
        ARAP_CUDA(upload_vector(send_index_dev, plan.send_index, stream));
        ARAP_CUDA(halo_sendbuf.ensure((size_t)(plan.n_send() > 0 ? plan.n_send() : 1) * 32));
        if (kind == ARAP_TRANSPORT_NCCL) {
            if (!id || id_bytes < 128) return fail(ARAP_ERR_INVALID, "attach_partition: NCCL needs the 128-byte unique id");
            std::unique_ptr<NcclTransport> t(new NcclTransport());
            if (t->init(rank, world, id)) return fail(ARAP_ERR_CUDA, t->error);
            transport = std::move(t);
        } else if (kind == ARAP_TRANSPORT_IN_PROCESS) {
            if (!id || id_bytes < 4) return fail(ARAP_ERR_INVALID, "attach_partition: in-process transport needs an int32 group key");
            std::unique_ptr<LocalTransport> t(new LocalTransport());
            if (t->init(rank, world, *(const int32_t *)id)) return fail(ARAP_ERR_INVALID, t->error);
            transport = std::move(t);
        } else if (kind == ARAP_TRANSPORT_PEER || kind == ARAP_TRANSPORT_PEER_IN_PROCESS) {
            // direct stores into the peers' memory; the other transport of the pair only bootstraps (setup-time all-gathers)
            std::unique_ptr<Transport> boot;
            if (kind == ARAP_TRANSPORT_PEER) {
                if (!id || id_bytes < 128) return fail(ARAP_ERR_INVALID, "attach_partition: the peer transport bootstraps over NCCL and needs the 128-byte unique id");
                std::unique_ptr<NcclTransport> t(new NcclTransport());
                if (t->init(rank, world, id)) return fail(ARAP_ERR_CUDA, t->error);
                boot = std::move(t);
            } else {
                if (!id || id_bytes < 4) return fail(ARAP_ERR_INVALID, "attach_partition: in-process transport needs an int32 group key");
                std::unique_ptr<LocalTransport> t(new LocalTransport());
                if (t->init(rank, world, *(const int32_t *)id)) return fail(ARAP_ERR_INVALID, t->error);
                boot = std::move(t);
            }
            std::unique_ptr<PeerTransport> t(new PeerTransport());
            if (t->init(std::move(boot), kind == ARAP_TRANSPORT_PEER_IN_PROCESS, rank, world)) return fail(ARAP_ERR_CUDA, t->error);
            transport = std::move(t);
        } else {
            return fail(ARAP_ERR_INVALID, "attach_partition: unknown transport");
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        dirty = true;
        prepared = false;
        have_perm = false;
        tiles_built = false;              // the owned block is renumbered on the next prepare: the tile structure follows
        use_tiles = false;
        return ARAP_OK;
    }

    // refresh the halo slots of a per-vertex array from their owners (no-op on a single GPU)
    int exchange_halo(void *array, size_t elem_bytes, int site) {
        if (!transport) return ARAP_OK;
        pdl_next_plain = true;
        if (comm_counting) { comm_exchanges += 1; comm_bytes += (long long)plan.n_send() * (long long)elem_bytes; }
        begin_launch(ARAP_K_HALO_PACK);
        const int rc = transport->exchange(stream, site, plan, send_index_dev.ptr, (char *)halo_sendbuf.ptr, (char *)array, elem_bytes);
        end_launch();
        if (rc) return fail(ARAP_ERR_CUDA, transport->error);
        return ARAP_OK;
    }

    int set_global_mesh(const arap_global_mesh *g) override {
        if (!transport) return fail(ARAP_ERR_INVALID, "set_global_mesh: call arap_attach_partition first");
        if (!g || g->n_vertices <= 0 || g->n_faces < 0 || !g->faces || !g->rest_xyz || !g->owner || !g->local_to_global ||
            (g->rest_scalar_bytes != 4 && g->rest_scalar_bytes != 8) || g->n_constrained < 0 || (g->n_constrained > 0 && !g->constrained))
            return fail(ARAP_ERR_INVALID, "set_global_mesh: bad arguments");
        GlobalMesh &m = global_mesh;
        m.n_vertices = g->n_vertices;
        m.n_faces = g->n_faces;
        m.faces.assign(g->faces, g->faces + 3 * (size_t)g->n_faces);
        for (int v : m.faces) if (v < 0 || v >= m.n_vertices) return fail(ARAP_ERR_INVALID, "set_global_mesh: face references a vertex out of range");
        m.owner.assign(g->owner, g->owner + g->n_vertices);
        for (int o : m.owner) if (o < 0 || o >= world_size) return fail(ARAP_ERR_INVALID, "set_global_mesh: owner rank out of range");
        m.local_to_global.assign(g->local_to_global, g->local_to_global + n_vertices);
        for (int v : m.local_to_global) if (v < 0 || v >= m.n_vertices) return fail(ARAP_ERR_INVALID, "set_global_mesh: local_to_global out of range");
        m.rest.resize(3 * (size_t)m.n_vertices);
        for (size_t k = 0; k < m.rest.size(); ++k)
            m.rest[k] = g->rest_scalar_bytes == 4 ? (double)((const float *)g->rest_xyz)[k] : ((const double *)g->rest_xyz)[k];
        m.constrained.assign((size_t)m.n_vertices, 0);
        for (int k = 0; k < g->n_constrained; ++k) {
            if (g->constrained[k] < 0 || g->constrained[k] >= m.n_vertices) return fail(ARAP_ERR_INVALID, "set_global_mesh: constrained index out of range");
            m.constrained[(size_t)g->constrained[k]] = 1;
        }
        have_global = true;
        length_scale = 0;              // re-measure on the global box
        dirty = true;
        return ARAP_OK;
    }

    // refresh the halo slots of one multigrid level's vector (global hierarchy, partitioned mode); which: see SITE_LEVEL_BASE
    int exchange_level(int level, int which, MgVec *array) {
        MgLevelDev &lv = *mg[(size_t)level];
        pdl_next_plain = true;
        if (comm_counting) { comm_exchanges += 1; comm_bytes += (long long)lv.plan.n_send() * (long long)sizeof(MgVec); }
        begin_launch(ARAP_K_HALO_PACK);
        const int rc = transport->exchange(stream, SITE_LEVEL_BASE + 4 * level + which, lv.plan, lv.send_index.ptr, (char *)mg_sendbuf.ptr,
                                           (char *)array, sizeof(MgVec));
        end_launch();
        if (rc) return fail(ARAP_ERR_CUDA, transport->error);
        return ARAP_OK;
    }

    // tell the transport about every call site (peer transport: mailboxes and flags per site); collective
    int configure_transport() {
        std::vector<SiteSpec> sites((size_t)SITE_COUNT);
        auto ex = [&](int site, const HaloPlan *pl, int bytes) { sites[(size_t)site].kind = SiteSpec::EXCHANGE; sites[(size_t)site].plan = pl; sites[(size_t)site].elem_bytes = bytes; };
        ex(SITE_CUR4, &plan, (int)sizeof(Vec4T<S>));
        ex(SITE_QUAT, &plan, (int)sizeof(Vec4T<S>));
        ex(SITE_CG_D, &plan, use_mg ? (int)sizeof(MgVec) : (int)sizeof(Vec3d));       // what the CG's matrix-vector product gathers: z (fp32) or d
        for (int st = 0; st <= CG_STAGE_MERGED; ++st) { sites[(size_t)(SITE_RED_BASE + st)].kind = SiteSpec::REDUCE_F64; sites[(size_t)(SITE_RED_BASE + st)].n = 8; }
        if (use_mg && mg_global) {
            ex(SITE_MG0_X_PRE, &plan, (int)sizeof(MgVec));
            ex(SITE_MG0_R, &plan, (int)sizeof(MgVec));
            ex(SITE_MG0_X_POST, &plan, (int)sizeof(MgVec));
            for (int l = 1; l < mg_first_replicated; ++l)
                for (int w = 0; w < 4; ++w) {
                    if (SITE_LEVEL_BASE + 4 * l + w >= SITE_COUNT) return fail(ARAP_ERR_SOLVER, "too many multigrid levels for the transport's site table");
                    ex(SITE_LEVEL_BASE + 4 * l + w, &mg[(size_t)l]->plan, (int)sizeof(MgVec));
                }
            sites[SITE_COARSE_B].kind = SiteSpec::REDUCE_F32;
            sites[SITE_COARSE_B].n = 4 * mg[(size_t)mg_first_replicated]->n;
        }
        if (transport->configure(stream, sites)) return fail(ARAP_ERR_CUDA, transport->error);
        return ARAP_OK;
    }

    int warm_up_transport() {
        if (use_mg) { int rc = exchange_halo(mg[0]->x2.ptr, sizeof(MgVec), SITE_CG_D); if (rc) return rc; }
        else { int rc = exchange_halo(cg_d.ptr, sizeof(Vec3d), SITE_CG_D); if (rc) return rc; }
        { int rc = exchange_halo(cur4.ptr, sizeof(Vec4T<S>), SITE_CUR4); if (rc) return rc; }
        double *red = (double *)((char *)cg.ptr + offsetof(CgScalars, red));
        if (transport->allreduce_sum(stream, SITE_RED_BASE, red, 8)) return fail(ARAP_ERR_CUDA, transport->error);
        if (use_mg && mg_global) {
            { int rc = exchange_halo(mg[0]->x.ptr, sizeof(MgVec), SITE_MG0_X_PRE); if (rc) return rc; }
            for (int l = 1; l < mg_first_replicated; ++l) { int rc = exchange_level(l, 0, mg[(size_t)l]->x.ptr); if (rc) return rc; }
            MgLevelDev &cl = *mg[(size_t)mg_first_replicated];
            if (transport->allreduce_sum_f32(stream, SITE_COARSE_B, (float *)cl.b.ptr, 4 * cl.n)) return fail(ARAP_ERR_CUDA, transport->error);
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }

    // partitioned mode: sum the stage's partial sums over the ranks, then finish the stage (alpha / beta / convergence)
    int reduce_stage(int stage, int n_values) {
        if (!transport) return ARAP_OK;
        double *red = (double *)((char *)cg.ptr + offsetof(CgScalars, red));
        (void)n_values;             // always the whole red[8] block: one site layout for every stage
        if (comm_counting) comm_allreduces += 1;
        begin_launch(ARAP_K_ALLREDUCE_SCALARS);
        const int rc_red = transport->allreduce_sum(stream, SITE_RED_BASE + stage, red, 8);
        end_launch();
        if (rc_red) return fail(ARAP_ERR_CUDA, transport->error);
        pdl_next_plain = true;
        begin_launch(ARAP_K_CG_FINALIZE);
        cg_finalize_kernel<<<1, 1, 0, stream>>>(cg.ptr, stage);
        end_launch();
        return ARAP_OK;
    }

    // Coarsening stops at the first level with at most 2048 rows (ARAP_MG_COARSE_ROWS): up to 256 rows are inverted on the
    // host, larger ones on the device. A 1234-row coarsest level instead of 1234 -> 125 saves four launches per V-cycle.
    static MgSetupOptions engine_mg_options() {
        MgSetupOptions mo;
        const char *env = getenv("ARAP_MG_COARSE_ROWS");
        mo.coarse_size = env ? atoi(env) : 2048;
        if (mo.coarse_size > mo.max_dense) mo.coarse_size = mo.max_dense;
        if (mo.coarse_size < 16) mo.coarse_size = 16;
        mo.host_dense_max = 256;
        return mo;
    }

    // ---- multigrid setup: host analysis of L (the reference's _solver.compute(_L), arap.h:337) ------------
    // Batches of small meshes: every member has the same operator, so the preconditioner is member 0's dense inverse applied to
    // all members (mg_kernels.cuh, mg_batch_dense_kernel). Returns with mg_batch_dense == false if that is not possible.
    int setup_batch_dense(const std::vector<int> &h_rowptr, const std::vector<int> &h_colidx, const std::vector<S> &h_w,
                          const std::vector<unsigned char> &h_con) {
        mg_batch_dense = false;
        const int Vm = member_vertices;
        if (batch_members <= 1 || transport || Vm <= 0 || Vm > 2048 || (long long)Vm * batch_members != n_vertices) return ARAP_OK;
        if (getenv("ARAP_BATCH_DENSE") && atoi(getenv("ARAP_BATCH_DENSE")) == 0) return ARAP_OK;
        const int e = h_rowptr[(size_t)Vm];
        for (int k = 0; k < e; ++k) if (h_colidx[(size_t)k] >= Vm) return ARAP_OK;            // not block diagonal after all
        // one operator for all members only if every member has member 0's constrained SET (arap_batch_set_constraints guarantees
        // it; constraints set through the raw handle might not): otherwise the general hierarchy over the whole batch
        for (int m = 1; m < batch_members; ++m)
            if (std::memcmp(&h_con[(size_t)m * Vm], &h_con[0], (size_t)Vm) != 0) return ARAP_OK;
        MgHierarchyHost H;
        MgSetupOptions mo = engine_mg_options();
        mo.coarse_size = std::max(mo.coarse_size, Vm);
        mg_build_hierarchy<S>(Vm, h_rowptr.data(), h_colidx.data(), h_w.data(), h_con.data(), mo, H, nullptr);
        if (H.levels.size() != 1) return ARAP_OK;
        if (!H.coarse_inv.empty()) { int rc = upload_coarse_inverse(H.coarse_inv, H.n_coarse); if (rc) return rc; mg_dense = true; }
        else if (H.coarse_dense_on_device) { int rc = invert_coarsest_on_device(H.levels.back().A); if (rc) return rc; }
        else return ARAP_OK;
        if (!mg_dense) return ARAP_OK;
        mg.clear();
        std::unique_ptr<MgLevelDev> d(new MgLevelDev());
        d->n = n_vertices;
        d->omega = H.levels[0].omega;
        ARAP_CUDA(d->x.ensure((size_t)n_vertices));
        ARAP_CUDA(d->x2.ensure((size_t)n_vertices));
        ARAP_CUDA(cudaMemsetAsync(d->x.ptr, 0, sizeof(MgVec) * (size_t)n_vertices, stream));
        ARAP_CUDA(cudaMemsetAsync(d->x2.ptr, 0, sizeof(MgVec) * (size_t)n_vertices, stream));
        mg.push_back(std::move(d));
        // tensor-core path: pack the inverse once into stage images, hi + lo TF32 parts (ARAP_BATCH_TC=0 keeps the SIMT fp32 GEMM)
        mg_batch_tc = !(getenv("ARAP_BATCH_TC") && atoi(getenv("ARAP_BATCH_TC")) == 0);
        if (mg_batch_tc) {
            const int n_mt = (Vm + kTcM - 1) / kTcM, n_nt = (batch_members + kTcMembers - 1) / kTcMembers, n_ks = (Vm + kTcK - 1) / kTcK;
            ARAP_CUDA(mg_a_pack.ensure((size_t)n_mt * n_ks * 2 * kTcAFloats));
            ARAP_CUDA(mg_b_pack.ensure((size_t)n_nt * n_ks * 2 * kTcBFloats));
            const size_t items = (size_t)n_mt * n_ks * kTcKcores * kTcM;
            begin_launch(ARAP_K_MISC);
            batch_pack_a_kernel<<<grid_for(items), kBlock, 0, stream>>>(Vm, mg_coarse_ld, n_mt, n_ks, mg_coarse_inv.ptr, mg_a_pack.ptr);
            end_launch();
            ARAP_CUDA(cudaFuncSetAttribute(mg_batch_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        tail_first = 0;
        stats.mg_levels = 1;
        stats.mg_operator_complexity = 1.0;
        mg_batch_dense = true;
        return ARAP_OK;
    }


    // ---- multigrid setup ON THE DEVICE (mg_setup_device.cuh): the same smoothed-aggregation hierarchy as mg_setup.cpp ----
    struct DevCsr {
        DeviceBuffer<int> rowptr, colidx;
        DeviceBuffer<double> val;
        int n_rows = 0, n_cols = 0, nnz = 0;
    };
    // row lengths in `len` (n entries) -> out.rowptr (n + 1), allocates colidx / val; one host round trip for nnz
    int csr_allocate(DevCsr &out, int n_rows_, int n_cols_, const int *len) {
        out.n_rows = n_rows_;
        out.n_cols = n_cols_;
        ARAP_CUDA(out.rowptr.ensure((size_t)n_rows_ + 1));
        { int rc = exclusive_scan(len, n_rows_, out.rowptr.ptr); if (rc) return rc; }
        ARAP_CUDA(cudaMemcpyAsync(&out.nnz, out.rowptr.ptr + n_rows_, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(out.colidx.ensure((size_t)out.nnz));
        ARAP_CUDA(out.val.ensure((size_t)out.nnz));
        return ARAP_OK;
    }
    int upload_level_matrix(const DevCsr &m, DeviceBuffer<int> &rp, DeviceBuffer<int> &ci, DeviceBuffer<float> &v) {
        ARAP_CUDA(rp.ensure((size_t)m.n_rows + 1));
        ARAP_CUDA(ci.ensure((size_t)m.nnz));
        ARAP_CUDA(v.ensure((size_t)m.nnz));
        ARAP_CUDA(cudaMemcpyAsync(rp.ptr, m.rowptr.ptr, sizeof(int) * ((size_t)m.n_rows + 1), cudaMemcpyDeviceToDevice, stream));
        if (m.nnz > 0) {
            ARAP_CUDA(cudaMemcpyAsync(ci.ptr, m.colidx.ptr, sizeof(int) * (size_t)m.nnz, cudaMemcpyDeviceToDevice, stream));
            mgdev::to_float_kernel<<<grid_for((size_t)m.nnz), kBlock, 0, stream>>>((size_t)m.nnz, m.val.ptr, v.ptr);
        }
        return ARAP_OK;
    }

    // One level of a hierarchy as the device setup leaves it (fp64): A, 1/diag, omega, and -- unless it is the coarsest -- P, R.
    struct DevLevel {
        std::unique_ptr<DevCsr> A, P, R;
        DeviceBuffer<double> inv_diag;
        DeviceBuffer<int> block;               // owner rank per row (partitioned hierarchies), else empty
        DeviceBuffer<int> block_next;          // ... of the next level's rows, until that level has copied it
        double omega = 2.0 / 3.0;
        int n = 0, active = 0;
    };

    // The hierarchy of L = D - W restricted to the rows with is_free != 0 (one-ring CSR rowptr / colidx / weight over V vertices),
    // built on the device. block0 (optional, device, V entries): aggregates never mix two blocks (partitioned mode).
    // *built = false (and ARAP_OK) when a row outgrew its accumulator: the caller falls back to the host setup.
    template <typename W>
    int build_hierarchy_device(int V, const int *d_rowptr, const int *d_colidx, const W *d_weight, const unsigned char *d_is_free,
                               const int *block0, const MgSetupOptions &mo, std::vector<std::unique_ptr<DevLevel>> &levels,
                               double *operator_complexity, bool *built, const void *pos_src = nullptr, int pos_scalar_bytes = 8,
                               int pos_stride = 3) {
        using namespace mgdev;
        *built = false;
        levels.clear();
        const double theta2 = mo.theta * mo.theta;
        const bool timing = getenv("ARAP_MG_TIMING") != nullptr;
        bool sweep_keys = false;
        DeviceBuffer<int> lex_work, lex_work_next, lex_roots, lex_adj, lex_far, lex_stamp, lex_count;
        DeviceBuffer<double> pos, pos_next, length_total;      // sweep keys: where the level's rows sit, the size of a sweep cell
        double cell = 0.0;
        DeviceBuffer<int> len, agg, status, flag, root_id, joined, scalars, cursor;
        DeviceBuffer<unsigned long long> m1, keys, gersh;
        DeviceBuffer<double> vx, vy, sums;
        ARAP_CUDA(scalars.ensure(4));        // [0] active rows, [1] newly elected roots, [2] still undecided, [3] accumulator overflow
        ARAP_CUDA(sums.ensure(4));
        ARAP_CUDA(gersh.ensure(1));
        ARAP_CUDA(cudaMemsetAsync(scalars.ptr, 0, 4 * sizeof(int), stream));
        // the reductions below use the engine's per-block partial sums; V may be the GLOBAL mesh of a partitioned handle, larger
        // than the local mesh the buffer was sized for
        ARAP_CUDA(partials.ensure((size_t)(std::min(grid_for((size_t)V), sm_count * 4) + 1) * 8));
        // ---- level 0 as an explicit CSR
        std::unique_ptr<DevCsr> A(new DevCsr());
        ARAP_CUDA(len.ensure((size_t)V + 1));
        level0_rows_kernel<W><<<grid_for((size_t)V), kBlock, 0, stream>>>(V, V, d_rowptr, d_colidx, d_weight, d_is_free, 0, len.ptr, nullptr, nullptr, nullptr);
        { int rc = csr_allocate(*A, V, V, len.ptr); if (rc) return rc; }
        level0_rows_kernel<W><<<grid_for((size_t)V), kBlock, 0, stream>>>(V, V, d_rowptr, d_colidx, d_weight, d_is_free, 1, nullptr, A->rowptr.ptr, A->colidx.ptr,
                                                                           A->val.ptr);
        const double fine_nnz = std::max(1, A->nnz);
        double total_nnz = 0;
        const int *block = block0;
        int h_scalars[4] = {0, 0, 0, 0};
        for (;;) {
            const int n = A->n_rows;
            const int l = (int)levels.size();
            std::unique_ptr<DevLevel> d(new DevLevel());
            d->n = n;
            if (block) {
                ARAP_CUDA(d->block.ensure((size_t)n));
                ARAP_CUDA(cudaMemcpyAsync(d->block.ptr, block, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
                block = d->block.ptr;
            }
            // ---- 1 / diagonal, number of active rows
            ARAP_CUDA(d->inv_diag.ensure((size_t)n));
            ARAP_CUDA(cudaMemsetAsync(scalars.ptr, 0, sizeof(int), stream));
            inv_diag_kernel<<<grid_for((size_t)n), kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, d->inv_diag.ptr, scalars.ptr);
            // ---- omega = 4 / (3 rho(D^-1 A)): 12 power iterations + Gershgorin bound
            ARAP_CUDA(vx.ensure((size_t)n));
            ARAP_CUDA(vy.ensure((size_t)n));
            ARAP_CUDA(cudaMemsetAsync(gersh.ptr, 0, sizeof(unsigned long long), stream));
            rho_init_kernel<<<grid_for((size_t)n), kBlock, 0, stream>>>(n, d->inv_diag.ptr, vx.ptr);
            const int rgrid = std::min(grid_for((size_t)n), sm_count * 4);
            for (int it = 0; it < 12; ++it) {
                rho_step_kernel<<<rgrid, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, d->inv_diag.ptr, vx.ptr, vy.ptr, partials.ptr, counter.ptr,
                                                              sums.ptr, it == 0 ? gersh.ptr : nullptr);
                rho_scale_kernel<<<grid_for((size_t)n), kBlock, 0, stream>>>(n, vy.ptr, sums.ptr, vx.ptr);
            }
            double h_sums[4] = {0, 0, 0, 0};
            unsigned long long h_gersh = 0;
            ARAP_CUDA(cudaMemcpyAsync(h_sums, sums.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaMemcpyAsync(&h_gersh, gersh.ptr, sizeof(h_gersh), cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaMemcpyAsync(h_scalars, scalars.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaStreamSynchronize(stream));
            double rho = h_sums[1] > 0 ? h_sums[0] / h_sums[1] : 1.0;
            double gbound;
            std::memcpy(&gbound, &h_gersh, sizeof(double));
            rho *= 1.1;
            if (gbound > 0) rho = std::min(rho, gbound);
            rho = std::max(rho, 1.0);
            d->omega = 4.0 / (3.0 * rho);
            const int active = h_scalars[0];
            d->active = active;
            total_nnz += A->nnz;
            bool coarsened = false;
            const bool last = active <= mo.coarse_size || l + 1 >= mo.max_levels;
            if (!last) {
                // ---- aggregation: roots = maximal independent set of the squared strength graph
                ARAP_CUDA(agg.ensure((size_t)n));
                ARAP_CUDA(status.ensure((size_t)n));
                ARAP_CUDA(m1.ensure((size_t)n));
                ARAP_CUDA(keys.ensure((size_t)n));
                ARAP_CUDA(flag.ensure((size_t)n + 1));
                ARAP_CUDA(root_id.ensure((size_t)n + 1));
                ARAP_CUDA(joined.ensure((size_t)n));
                const int G = grid_for((size_t)n);
                const double *idg = d->inv_diag.ptr;
                agg_init_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, agg.ptr, status.ptr);
                // which independent set: the lexicographically first one (a wavefront, lex_*_kernel) for a quad-like strength graph, rim
                // growth from sparse seeds (agg_key_kernel) otherwise; decided once, on the finest level (ARAP_MG_AGG_KEY=sweep|rim
                // overrides; ARAP_MG_WAVEFRONT=0 replaces the wavefront by the round-2b sweeps over cells with hashed priorities)
                if (l == 0) {
                    const bool no_wavefront = getenv("ARAP_MG_WAVEFRONT") && atoi(getenv("ARAP_MG_WAVEFRONT")) == 0;
                    if (pos_src && no_wavefront && !(getenv("ARAP_MG_SWEEP_CELLS") && atoi(getenv("ARAP_MG_SWEEP_CELLS")) == 0)) {
                        ARAP_CUDA(pos.ensure(3 * (size_t)n));
                        pos_gather_kernel<<<G, kBlock, 0, stream>>>(n, pos_src, pos_scalar_bytes, pos_stride, pos.ptr);
                    }
                    ARAP_CUDA(length_total.ensure(1));
                    ARAP_CUDA(cudaMemsetAsync(length_total.ptr, 0, sizeof(double), stream));
                    ARAP_CUDA(cudaMemsetAsync(gersh.ptr, 0, sizeof(unsigned long long), stream));
                    agg_strong_count_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, pos.ptr, gersh.ptr,
                                                                      length_total.ptr);
                    unsigned long long strong_total = 0;
                    double h_length = 0.0;
                    ARAP_CUDA(cudaMemcpyAsync(&strong_total, gersh.ptr, sizeof(strong_total), cudaMemcpyDeviceToHost, stream));
                    ARAP_CUDA(cudaMemcpyAsync(&h_length, length_total.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream));
                    ARAP_CUDA(cudaStreamSynchronize(stream));
                    const double mean_strong = active > 0 ? (double)strong_total / active : 0.0;
                    const char *kenv = getenv("ARAP_MG_AGG_KEY");
                    // (aggregates confined to MANY partition blocks take rim growth. Measured at 16M vertices, CG iterations per ARAP iteration:
                    //  8 GPUs, strips / blocks: rim growth 9.64 / 8.56, wavefront 11.08 / 8.84 -- coarse levels of 58-220 entries per row and a
                    //  seventh level; 2 and 4 partitions: wavefront 7.8 / 7.25, rim growth 8.2 / 8.4-9.2)
                    sweep_keys = kenv ? std::strcmp(kenv, "sweep") == 0 : (mean_strong <= 4.5 && (block == nullptr || world_size <= 4));
                    cell = strong_total > 0 ? 8.0 * h_length / (double)strong_total : 0.0;        // 8 mean edges: ~64 rows of a surface per cell
                    if (!(cell > 0.0) || !sweep_keys) pos.release();
                    if (timing) std::fprintf(stderr, "[mg device setup] %.2f strong connections per row: %s\n", mean_strong,
                                             !sweep_keys ? "rim growth" : !no_wavefront ? "wavefront (lexicographically first set)" : pos.ptr ? "sweeps over spatial cells" : "sweeps over runs of 64 rows");
                }
                // rim growth: 1 vertex in 2^bits is a seed (ARAP_MG_AGG_SEED_BITS, 0 = every vertex; small levels still get a handful);
                // where a round elects nothing although vertices remain (no seed in that component), the seed set is made 8x denser
                int seed_bits = getenv("ARAP_MG_AGG_SEED_BITS") ? atoi(getenv("ARAP_MG_AGG_SEED_BITS")) : 12;
                int log2n = 0;
                while ((2 << log2n) <= active) ++log2n;
                seed_bits = std::max(0, std::min(std::min(23, seed_bits), log2n - 3));
                int rounds_used = 0;
                const bool wavefront = sweep_keys && !(getenv("ARAP_MG_WAVEFRONT") && atoi(getenv("ARAP_MG_WAVEFRONT")) == 0);
                bool finish_with_rim = false;
                if (wavefront) {
                    // the lexicographically first independent set (the host's greedy walk), as a wavefront over worklists
                    ARAP_CUDA(lex_work.ensure((size_t)n));
                    ARAP_CUDA(lex_work_next.ensure((size_t)n));
                    ARAP_CUDA(lex_roots.ensure((size_t)n));
                    ARAP_CUDA(lex_adj.ensure((size_t)n));
                    ARAP_CUDA(lex_far.ensure((size_t)n));
                    ARAP_CUDA(lex_stamp.ensure((size_t)n));
                    ARAP_CUDA(lex_count.ensure(8));
                    ARAP_CUDA(cudaMemsetAsync(lex_count.ptr, 0, 8 * sizeof(int), stream));
                    LexLists L;
                    L.work = lex_work.ptr; L.work_next = lex_work_next.ptr; L.roots = lex_roots.ptr; L.adj = lex_adj.ptr; L.far = lex_far.ptr;
                    L.count = lex_count.ptr; L.stamp = lex_stamp.ptr;
                    lex_fill_kernel<<<G, kBlock, 0, stream>>>(n, status.ptr, L.work, L.count, L.stamp);
                    const int Gw = std::min(G, sm_count * 8);
                    const int lex_exact = getenv("ARAP_MG_WAVEFRONT_EXACT") && atoi(getenv("ARAP_MG_WAVEFRONT_EXACT")) != 0 ? 1 : 0;
                    int h_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    // round budget: a few times the side length of a surface mesh (the dependency chain of a compact mesh); a mesh with
                    // a longer chain (a ribbon a few rows wide) finishes with rim growth from what the wavefront has decided so far
                    int max_rounds = std::max(4096, 8 * (int)std::sqrt((double)n));
                    if (getenv("ARAP_MG_WAVEFRONT_ROUNDS")) max_rounds = std::max(1, atoi(getenv("ARAP_MG_WAVEFRONT_ROUNDS")));
                    for (int round = 0; round < max_rounds; ++round) {
                        // the first rounds sweep long lists (round 0: every undecided row); later ones the front only
                        const int Gr = Gw;                 // one warp per list entry in the elect / next kernels
                        lex_elect_kernel<<<Gr, kBlock, 0, stream>>>(L, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr);
                        lex_cover1_kernel<<<Gr, kBlock, 0, stream>>>(L, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr, lex_exact);
                        lex_cover2_kernel<<<Gr, kBlock, 0, stream>>>(L, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr);
                        lex_next_kernel<<<Gr, kBlock, 0, stream>>>(L, round, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr);
                        lex_advance_kernel<<<1, 1, 0, stream>>>(L.count, L.count + 5);
                        std::swap(L.work, L.work_next);
                        if ((round & 63) == 63) {
                            ARAP_CUDA(cudaMemcpyAsync(h_count, L.count, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream));
                            ARAP_CUDA(cudaStreamSynchronize(stream));
                            if (h_count[0] == 0) break;
                        }
                    }
                    ARAP_CUDA(cudaMemcpyAsync(h_count, L.count, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream));
                    ARAP_CUDA(cudaStreamSynchronize(stream));
                    ARAP_CUDA(cudaGetLastError());
                    rounds_used = h_count[5] + 1;
                    finish_with_rim = h_count[0] != 0;
                    if (finish_with_rim && timing)
                        std::fprintf(stderr, "[mg device setup] level %d: wavefront stopped after %d rounds, rim growth takes over\n", l, max_rounds);
                }
                for (int round = 0; round < 8192 && (!wavefront || finish_with_rim); ++round) {
                    ++rounds_used;
                    ARAP_CUDA(cudaMemsetAsync(scalars.ptr + 1, 0, 2 * sizeof(int), stream));
                    agg_key_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr, keys.ptr,
                                                             (1u << seed_bits) - 1u, sweep_keys && !finish_with_rim ? 1 : 0, pos.ptr, cell > 0.0 ? 1.0 / cell : 0.0);
                    agg_max1_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, keys.ptr, m1.ptr);
                    agg_elect_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, keys.ptr, m1.ptr, status.ptr, scalars.ptr + 1);
                    agg_cover1_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr);
                    agg_cover2_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr, scalars.ptr + 2);
                    ARAP_CUDA(cudaMemcpyAsync(h_scalars + 1, scalars.ptr + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
                    ARAP_CUDA(cudaStreamSynchronize(stream));
                    if (h_scalars[2] == 0) break;
                    if (h_scalars[1] == 0) seed_bits = std::max(0, seed_bits - 3);
                }
                agg_root_flag_kernel<<<G, kBlock, 0, stream>>>(n, status.ptr, flag.ptr);
                { int rc = exclusive_scan(flag.ptr, n, root_id.ptr); if (rc) return rc; }
                int n_agg = 0;
                ARAP_CUDA(cudaMemcpyAsync(&n_agg, root_id.ptr + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
                agg_assign1_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, status.ptr, root_id.ptr, agg.ptr);
                agg_assign2_kernel<<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, block, theta2, agg.ptr, joined.ptr);
                ARAP_CUDA(cudaStreamSynchronize(stream));
                if (n_agg > 0 && n_agg < 0.8 * active) {
                    // ---- P = (I - omega D^-1 A) T
                    d->P.reset(new DevCsr());
                    d->R.reset(new DevCsr());
                    DevCsr &P = *d->P, &R = *d->R;
                    DevCsr AP;
                    std::unique_ptr<DevCsr> Ac(new DevCsr());
                    // (rows of P: 64 entries, 512 on the dense coarse levels thin partition blocks produce -- 111 entries per row of A there)
                    bool wide_p = A->nnz > 40 * (long long)n;
                    for (int attempt = wide_p ? 1 : 0; attempt < 2; ++attempt) {
                        wide_p = attempt == 1;
                        ARAP_CUDA(cudaMemsetAsync(scalars.ptr + 3, 0, sizeof(int), stream));
                        if (!wide_p) prolongator_kernel<64><<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, joined.ptr, d->omega, 0, len.ptr,
                                                                                     nullptr, nullptr, nullptr, scalars.ptr + 3);
                        else prolongator_kernel<512><<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, joined.ptr, d->omega, 0, len.ptr,
                                                                               nullptr, nullptr, nullptr, scalars.ptr + 3);
                        ARAP_CUDA(cudaMemcpyAsync(h_scalars + 3, scalars.ptr + 3, sizeof(int), cudaMemcpyDeviceToHost, stream));
                        ARAP_CUDA(cudaStreamSynchronize(stream));
                        if (h_scalars[3] == 0) break;
                    }
                    if (h_scalars[3] != 0) {
                        if (timing) std::fprintf(stderr, "[mg device setup] level %d (%d rows, %d nnz): a row of P outgrew 512 entries, host setup instead\n", l, n, A->nnz);
                        return ARAP_OK;
                    }
                    { int rc = csr_allocate(P, n, n_agg, len.ptr); if (rc) return rc; }
                    if (!wide_p) prolongator_kernel<64><<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, joined.ptr, d->omega, 1, nullptr,
                                                                                 P.rowptr.ptr, P.colidx.ptr, P.val.ptr, scalars.ptr + 3);
                    else prolongator_kernel<512><<<G, kBlock, 0, stream>>>(n, A->rowptr.ptr, A->colidx.ptr, A->val.ptr, idg, joined.ptr, d->omega, 1, nullptr,
                                                                           P.rowptr.ptr, P.colidx.ptr, P.val.ptr, scalars.ptr + 3);
                    // ---- R = P^T, rows sorted
                    ARAP_CUDA(cursor.ensure((size_t)n_agg + 1));
                    ARAP_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(int) * ((size_t)n_agg + 1), stream));
                    if (P.nnz > 0) col_count_kernel<<<grid_for((size_t)P.nnz), kBlock, 0, stream>>>(P.nnz, P.colidx.ptr, cursor.ptr);
                    { int rc = csr_allocate(R, n_agg, n, cursor.ptr); if (rc) return rc; }
                    ARAP_CUDA(cudaMemsetAsync(cursor.ptr, 0, sizeof(int) * ((size_t)n_agg + 1), stream));
                    transpose_fill_kernel<<<G, kBlock, 0, stream>>>(n, P.rowptr.ptr, P.colidx.ptr, P.val.ptr, R.rowptr.ptr, cursor.ptr, R.colidx.ptr, R.val.ptr);
                    sort_rows_kernel<<<grid_for((size_t)n_agg), kBlock, 0, stream>>>(n_agg, R.rowptr.ptr, R.colidx.ptr, R.val.ptr);
                    // ---- AP = A P ; A_c = R (A P). Rows are accumulated in per-thread sorted arrays of 128 entries; a product whose rows
                    // outgrow that (coarse operators of irregular meshes) is redone with 512, and only then is the host setup the fallback.
                    auto product = [&](const DevCsr &X, const DevCsr &Y, DevCsr &Z) -> int {
                        const int rows = X.n_rows, Gp = grid_for((size_t)rows);
                        for (int attempt = 0; attempt < 3; ++attempt) {
                            if (attempt == 2 && rows > 200000) break;              // 24 KB of accumulator per thread: small levels only
                            ARAP_CUDA(cudaMemsetAsync(scalars.ptr + 3, 0, sizeof(int), stream));
                            if (attempt == 0)
                                spgemm_rows_kernel<128><<<Gp, kBlock, 0, stream>>>(rows, X.rowptr.ptr, X.colidx.ptr, X.val.ptr, Y.rowptr.ptr, Y.colidx.ptr, Y.val.ptr, 0, len.ptr,
                                                                                   nullptr, nullptr, nullptr, scalars.ptr + 3);
                            else if (attempt == 1)
                                spgemm_rows_kernel<512><<<Gp, kBlock, 0, stream>>>(rows, X.rowptr.ptr, X.colidx.ptr, X.val.ptr, Y.rowptr.ptr, Y.colidx.ptr, Y.val.ptr, 0, len.ptr,
                                                                                   nullptr, nullptr, nullptr, scalars.ptr + 3);
                            else
                                spgemm_rows_kernel<2048><<<Gp, kBlock, 0, stream>>>(rows, X.rowptr.ptr, X.colidx.ptr, X.val.ptr, Y.rowptr.ptr, Y.colidx.ptr, Y.val.ptr, 0, len.ptr,
                                                                                    nullptr, nullptr, nullptr, scalars.ptr + 3);
                            ARAP_CUDA(cudaMemcpyAsync(h_scalars + 3, scalars.ptr + 3, sizeof(int), cudaMemcpyDeviceToHost, stream));
                            ARAP_CUDA(cudaStreamSynchronize(stream));
                            if (h_scalars[3] != 0) continue;
                            { int rc = csr_allocate(Z, rows, Y.n_cols, len.ptr); if (rc) return rc; }
                            if (attempt == 0)
                                spgemm_rows_kernel<128><<<Gp, kBlock, 0, stream>>>(rows, X.rowptr.ptr, X.colidx.ptr, X.val.ptr, Y.rowptr.ptr, Y.colidx.ptr, Y.val.ptr, 1, nullptr,
                                                                                   Z.rowptr.ptr, Z.colidx.ptr, Z.val.ptr, scalars.ptr + 3);
                            else if (attempt == 1)
                                spgemm_rows_kernel<512><<<Gp, kBlock, 0, stream>>>(rows, X.rowptr.ptr, X.colidx.ptr, X.val.ptr, Y.rowptr.ptr, Y.colidx.ptr, Y.val.ptr, 1, nullptr,
                                                                                   Z.rowptr.ptr, Z.colidx.ptr, Z.val.ptr, scalars.ptr + 3);
                            else
                                spgemm_rows_kernel<2048><<<Gp, kBlock, 0, stream>>>(rows, X.rowptr.ptr, X.colidx.ptr, X.val.ptr, Y.rowptr.ptr, Y.colidx.ptr, Y.val.ptr, 1, nullptr,
                                                                                    Z.rowptr.ptr, Z.colidx.ptr, Z.val.ptr, scalars.ptr + 3);
                            return ARAP_OK;
                        }
                        return ARAP_OK;                                  // h_scalars[3] != 0: the caller gives up on the device path
                    };
                    // the prolongator's own flag first (it shares scalars[3] with the products)
                    ARAP_CUDA(cudaMemcpyAsync(h_scalars + 3, scalars.ptr + 3, sizeof(int), cudaMemcpyDeviceToHost, stream));
                    ARAP_CUDA(cudaStreamSynchronize(stream));
                    auto declined = [&](const char *what) {
                        if (timing) std::fprintf(stderr, "[mg device setup] level %d (%d rows, %d nnz): a row of %s outgrew its accumulator, host setup instead\n", l, n, A->nnz, what);
                        return ARAP_OK;
                    };
                    if (h_scalars[3] != 0) return declined("P");
                    { int rc = product(*A, P, AP); if (rc) return rc; }
                    if (h_scalars[3] != 0) return declined("A P");
                    { int rc = product(R, AP, *Ac); if (rc) return rc; }
                    if (h_scalars[3] != 0) return declined("R A P");
                    DeviceBuffer<int> block_c;
                    if (block) {
                        ARAP_CUDA(block_c.ensure((size_t)n_agg));
                        agg_block_kernel<<<G, kBlock, 0, stream>>>(n, status.ptr, root_id.ptr, block, block_c.ptr);
                    }
                    if (pos.ptr) {                                   // sweep cells of the next level: ~64 of ITS rows each
                        ARAP_CUDA(pos_next.ensure(3 * (size_t)n_agg));
                        pos_coarse_kernel<<<G, kBlock, 0, stream>>>(n, status.ptr, root_id.ptr, pos.ptr, pos_next.ptr);
                        cell *= std::sqrt((double)std::max(1, active) / (double)n_agg);
                    }
                    ARAP_CUDA(cudaMemcpyAsync(h_scalars + 3, scalars.ptr + 3, sizeof(int), cudaMemcpyDeviceToHost, stream));
                    ARAP_CUDA(cudaStreamSynchronize(stream));       // AP dies at scope exit
                    ARAP_CUDA(cudaGetLastError());
                    if (h_scalars[3] != 0) return ARAP_OK;             // a row outgrew its accumulator: the host setup handles it
                    if (timing) std::fprintf(stderr, "[mg device setup] level %d: %d rows (%d active, %d nnz) -> %d aggregates in %d rounds, omega %.4f\n", l, n, active, A->nnz, n_agg, rounds_used, d->omega);
                    d->A = std::move(A);
                    levels.push_back(std::move(d));
                    A = std::move(Ac);
                    if (pos.ptr) { std::swap(pos.ptr, pos_next.ptr); std::swap(pos.count, pos_next.count); }
                    if (block) {
                        // the next level copies its block array out of this temporary at the top of the loop; keep it alive there
                        levels.back()->block_next.ptr = block_c.ptr; levels.back()->block_next.count = block_c.count;
                        block_c.ptr = nullptr; block_c.count = 0;
                        block = levels.back()->block_next.ptr;
                    }
                    coarsened = true;
                }
            }
            if (coarsened) continue;
            if (timing) std::fprintf(stderr, "[mg device setup] level %d: %d rows (%d active, %d nnz), coarsest\n", l, n, active, A->nnz);
            d->A = std::move(A);
            levels.push_back(std::move(d));
            break;
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        if (operator_complexity) *operator_complexity = total_nnz / fine_nnz;
        *built = true;
        return ARAP_OK;
    }

    int download_csr(const DevCsr &m, HostCsr &h) {
        h.n_rows = m.n_rows;
        h.n_cols = m.n_cols;
        h.rowptr.resize((size_t)m.n_rows + 1);
        h.colidx.resize((size_t)m.nnz);
        h.val.resize((size_t)m.nnz);
        ARAP_CUDA(cudaMemcpyAsync(h.rowptr.data(), m.rowptr.ptr, sizeof(int) * ((size_t)m.n_rows + 1), cudaMemcpyDeviceToHost, stream));
        if (m.nnz > 0) {
            ARAP_CUDA(cudaMemcpyAsync(h.colidx.data(), m.colidx.ptr, sizeof(int) * (size_t)m.nnz, cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaMemcpyAsync(h.val.data(), m.val.ptr, sizeof(double) * (size_t)m.nnz, cudaMemcpyDeviceToHost, stream));
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }

    // Builds `mg` for the single-GPU solver on the device. *built = false (and ARAP_OK) when the device path declines (an
    // accumulator row overflowed, switched off with ARAP_MG_DEVICE_SETUP=0): the caller then runs the host setup.
    int setup_multigrid_device(bool *built) {
        using namespace mgdev;
        *built = false;
        if (getenv("ARAP_MG_DEVICE_SETUP") && atoi(getenv("ARAP_MG_DEVICE_SETUP")) == 0) return ARAP_OK;
        if (transport || batch_members > 1) return ARAP_OK;
        const auto t0 = std::chrono::steady_clock::now();
        const MgSetupOptions mo = engine_mg_options();
        std::vector<std::unique_ptr<DevLevel>> dl;
        double complexity = 0;
        bool ok = false;
        DeviceBuffer<int> fake_block;
        if (getenv("ARAP_MG_FAKE_BLOCKS") && atoi(getenv("ARAP_MG_FAKE_BLOCKS")) > 1) {
            ARAP_CUDA(fake_block.ensure((size_t)n_vertices));
            mgdev::fake_block_kernel<<<grid_for((size_t)n_vertices), kBlock, 0, stream>>>(n_vertices, atoi(getenv("ARAP_MG_FAKE_BLOCKS")), fake_block.ptr);
        }
        { int rc = build_hierarchy_device<S>(n_vertices, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, free_mask.ptr, fake_block.ptr, mo, dl, &complexity, &ok,
                                                 rest4.ptr, (int)sizeof(S), 4); if (rc) return rc; }
        if (!ok) return ARAP_OK;
        std::vector<std::unique_ptr<MgLevelDev>> levels;
        const size_t L = dl.size();
        for (size_t l = 0; l < L; ++l) {
            DevLevel &src = *dl[l];
            const int n = src.n;
            std::unique_ptr<MgLevelDev> d(new MgLevelDev());
            d->n = n;
            d->omega = src.omega;
            if (l > 0) {
                { int rc = upload_level_matrix(*src.A, d->a_rowptr, d->a_colidx, d->a_val); if (rc) return rc; }
                ARAP_CUDA(d->b.ensure((size_t)n));
                d->a_lanes = pick_lanes((size_t)src.A->nnz, (size_t)n);
            }
            ARAP_CUDA(d->inv_diag.ensure((size_t)n));
            to_float_kernel<<<grid_for((size_t)n), kBlock, 0, stream>>>((size_t)n, src.inv_diag.ptr, d->inv_diag.ptr);
            if (l + 1 < L) {
                { int rc = upload_level_matrix(*src.P, d->p_rowptr, d->p_colidx, d->p_val); if (rc) return rc; }
                { int rc = upload_level_matrix(*src.R, d->r_rowptr, d->r_colidx, d->r_val); if (rc) return rc; }
                ARAP_CUDA(d->r.ensure((size_t)n));
                d->r_lanes = pick_lanes((size_t)src.R->nnz, (size_t)src.R->n_rows);
            }
            ARAP_CUDA(d->x.ensure((size_t)n));
            ARAP_CUDA(d->x2.ensure((size_t)n));
            ARAP_CUDA(cudaMemsetAsync(d->x.ptr, 0, sizeof(MgVec) * (size_t)(n > 0 ? n : 1), stream));
            ARAP_CUDA(cudaMemsetAsync(d->x2.ptr, 0, sizeof(MgVec) * (size_t)(n > 0 ? n : 1), stream));
            levels.push_back(std::move(d));
        }
        // ---- coarsest level: dense inverse on the device, or one more smoothing level when it is too large
        mg_dense = false;
        if (dl.back()->n <= mo.max_dense) {
            HostCsr hA;
            { int rc = download_csr(*dl.back()->A, hA); if (rc) return rc; }
            int rc = invert_coarsest_on_device(hA);
            if (rc) return rc;
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        mg.swap(levels);
        { int rc = plan_tail(); if (rc) return rc; }
        stats.mg_levels = (int)mg.size();
        stats.mg_operator_complexity = complexity;
        stats.setup_host_ms = 0.0;
        stats.setup_device_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *built = true;
        return ARAP_OK;
    }

    int setup_multigrid() {
        {   // the device path first; the host path below remains for batches, partitions and as the fallback
            bool built = false;
            int rc = setup_multigrid_device(&built);
            if (rc) return rc;
            if (built) return ARAP_OK;
        }
        const int V = n_vertices;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<int> h_rowptr((size_t)V + 1), h_colidx((size_t)nnz);
        std::vector<S> h_w((size_t)nnz);
        std::vector<unsigned char> h_con((size_t)V);
        ARAP_CUDA(cudaMemcpyAsync(h_rowptr.data(), hot_rowptr.ptr, sizeof(int) * ((size_t)V + 1), cudaMemcpyDeviceToHost, stream));
        if (nnz > 0) {
            ARAP_CUDA(cudaMemcpyAsync(h_colidx.data(), hot_colidx.ptr, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaMemcpyAsync(h_w.data(), hot_weight.ptr, sizeof(S) * (size_t)nnz, cudaMemcpyDeviceToHost, stream));
        }
        std::vector<unsigned char> h_con_user((size_t)V);
        std::vector<int> h_perm((size_t)V);
        if (V > 0) {
            ARAP_CUDA(cudaMemcpyAsync(h_con_user.data(), is_constrained.ptr, (size_t)V, cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaMemcpyAsync(h_perm.data(), perm.ptr, sizeof(int) * (size_t)V, cudaMemcpyDeviceToHost, stream));
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        for (int v = 0; v < V; ++v) h_con[(size_t)v] = h_con_user[(size_t)h_perm[(size_t)v]];
        for (int v = n_rows; v < V; ++v) h_con[(size_t)v] = 1;      // partitioned mode: block-Jacobi across ranks, halo = Dirichlet
        { int rc = setup_batch_dense(h_rowptr, h_colidx, h_w, h_con); if (rc) return rc; }
        if (mg_batch_dense) {
            stats.setup_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            return ARAP_OK;
        }
        MgHierarchyHost H;
        MgSetupOptions mo = engine_mg_options();
        mg_build_hierarchy<S>(V, h_rowptr.data(), h_colidx.data(), h_w.data(), h_con.data(), mo, H,
                              (int)mg_visit_order.size() == V ? mg_visit_order.data() : nullptr);
        mg.clear();
        std::vector<float> fscratch;
        for (size_t l = 0; l < H.levels.size(); ++l) {
            const MgLevelHost &hl = H.levels[l];
            if (getenv("ARAP_MG_TIMING"))
                std::fprintf(stderr, "[mg host setup] level %zu: %d rows (%zu nnz) -> %d aggregates, omega %.4f\n", l, hl.A.n_rows,
                             hl.A.colidx.size(), hl.P.n_cols, hl.omega);
            std::unique_ptr<MgLevelDev> d(new MgLevelDev());
            d->n = hl.A.n_rows;
            d->omega = hl.omega;
            if (l > 0) {
                ARAP_CUDA(upload_vector(d->a_rowptr, hl.A.rowptr, stream));
                ARAP_CUDA(upload_vector(d->a_colidx, hl.A.colidx, stream));
                ARAP_CUDA(upload_as_float(d->a_val, hl.A.val, stream, fscratch));
                ARAP_CUDA(d->b.ensure((size_t)d->n));
            }
            ARAP_CUDA(upload_as_float(d->inv_diag, hl.inv_diag, stream, fscratch));
            if (l + 1 < H.levels.size()) {
                ARAP_CUDA(upload_vector(d->p_rowptr, hl.P.rowptr, stream));
                ARAP_CUDA(upload_vector(d->p_colidx, hl.P.colidx, stream));
                ARAP_CUDA(upload_as_float(d->p_val, hl.P.val, stream, fscratch));
                ARAP_CUDA(upload_vector(d->r_rowptr, hl.R.rowptr, stream));
                ARAP_CUDA(upload_vector(d->r_colidx, hl.R.colidx, stream));
                ARAP_CUDA(upload_as_float(d->r_val, hl.R.val, stream, fscratch));
                ARAP_CUDA(d->r.ensure((size_t)d->n));
            }
            ARAP_CUDA(d->x.ensure((size_t)d->n));
            ARAP_CUDA(d->x2.ensure((size_t)d->n));
            ARAP_CUDA(cudaMemsetAsync(d->x.ptr, 0, sizeof(MgVec) * (size_t)(d->n > 0 ? d->n : 1), stream));
            ARAP_CUDA(cudaMemsetAsync(d->x2.ptr, 0, sizeof(MgVec) * (size_t)(d->n > 0 ? d->n : 1), stream));
            d->a_lanes = pick_lanes(hl.A.colidx.size(), (size_t)hl.A.n_rows);
            d->r_lanes = pick_lanes(hl.R.colidx.size(), (size_t)hl.R.n_rows);
            mg.push_back(std::move(d));
        }
        mg_dense = !H.coarse_inv.empty();
        if (mg_dense) { int rc = upload_coarse_inverse(H.coarse_inv, H.n_coarse); if (rc) return rc; }
        else if (H.coarse_dense_on_device) { int rc = invert_coarsest_on_device(H.levels.back().A); if (rc) return rc; }
        ARAP_CUDA(cudaStreamSynchronize(stream));     // host vectors die at scope exit
        { int rc = plan_tail(); if (rc) return rc; }
        stats.mg_levels = (int)mg.size();
        stats.mg_operator_complexity = H.operator_complexity;
        stats.setup_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return ARAP_OK;
    }

    // the host-computed dense inverse (n x n doubles) -> mg_coarse_inv (floats, rows padded to a multiple of 4)
    int upload_coarse_inverse(const std::vector<double> &inv, int n) {
        mg_coarse_ld = (n + 3) & ~3;
        std::vector<float> padded((size_t)n * mg_coarse_ld, 0.f);
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) padded[(size_t)r * mg_coarse_ld + c] = (float)inv[(size_t)r * n + c];
        ARAP_CUDA(mg_coarse_inv.ensure(padded.size()));
        ARAP_CUDA(cudaMemcpyAsync(mg_coarse_inv.ptr, padded.data(), sizeof(float) * padded.size(), cudaMemcpyHostToDevice, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }

    // Dense fp32 inverse of the coarsest operator A (host CSR) into mg_coarse_inv, inverted on the device (mg_kernels.cuh).
    // Sets mg_dense; on a bad pivot the level simply stays a smoothing level (mg_dense = false).
    int invert_coarsest_on_device(const HostCsr &A) {
        const int n = A.n_rows;
        mg_dense = false;
        if (n <= 0) return ARAP_OK;
        const auto t0 = std::chrono::steady_clock::now();
        DeviceBuffer<int> d_rowptr, d_colidx, d_bad;
        DeviceBuffer<double> d_val, d_M, d_D, d_C, d_R;
        ARAP_CUDA(upload_vector(d_rowptr, A.rowptr, stream));
        ARAP_CUDA(upload_vector(d_colidx, A.colidx, stream));
        ARAP_CUDA(upload_vector(d_val, A.val, stream));
        ARAP_CUDA(d_M.ensure((size_t)n * n));
        ARAP_CUDA(d_D.ensure((size_t)kGjB * kGjB));
        ARAP_CUDA(d_C.ensure((size_t)n * kGjB));
        ARAP_CUDA(d_R.ensure((size_t)n * kGjB));
        ARAP_CUDA(d_bad.ensure(1));
        ARAP_CUDA(cudaMemsetAsync(d_M.ptr, 0, sizeof(double) * (size_t)n * n, stream));
        ARAP_CUDA(cudaMemsetAsync(d_bad.ptr, 0, sizeof(int), stream));
        double trace = 0;
        for (int i = 0; i < n; ++i)
            for (int k = A.rowptr[(size_t)i]; k < A.rowptr[(size_t)i + 1]; ++k) if (A.colidx[(size_t)k] == i) trace += A.val[(size_t)k];
        const double shift = 1e-13 * trace / n;             // keeps a pure-Neumann component invertible (as mg_setup.cpp does)
        begin_launch(ARAP_K_MISC);
        dense_from_csr_kernel<<<grid_for((size_t)n), kBlock, 0, stream>>>(n, d_rowptr.ptr, d_colidx.ptr, d_val.ptr, shift, d_M.ptr);
        const dim3 ugrid((unsigned)((n + 15) / 16), (unsigned)((n + 15) / 16), 1);
        const int copy_ctas = std::max(1, std::min(64, (n * kGjB + 1023) / 1024));
        for (int k0 = 0; k0 < n; k0 += kGjB) {                  // blocked Gauss-Jordan: three launches per panel of 32 pivots
            const int nb = std::min(kGjB, n - k0);
            gj_panel_kernel<<<1 + copy_ctas, 1024, 0, stream>>>(n, k0, nb, d_M.ptr, d_D.ptr, d_C.ptr, d_bad.ptr);
            gj_row_kernel<<<grid_for((size_t)n), kBlock, 0, stream>>>(n, k0, nb, d_M.ptr, d_D.ptr, d_R.ptr);
            gj_update_kernel<<<ugrid, 256, 0, stream>>>(n, k0, nb, d_M.ptr, d_C.ptr, d_R.ptr, d_D.ptr);
        }
        mg_coarse_ld = (n + 3) & ~3;
        ARAP_CUDA(mg_coarse_inv.ensure((size_t)n * mg_coarse_ld));
        dense_to_float_kernel<<<grid_for((size_t)n * mg_coarse_ld), kBlock, 0, stream>>>(n, mg_coarse_ld, d_M.ptr, mg_coarse_inv.ptr);
        end_launch();
        int bad = 0;
        ARAP_CUDA(cudaMemcpyAsync(&bad, d_bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        mg_dense = (bad == 0);
        if (getenv("ARAP_MG_TIMING"))
            std::fprintf(stderr, "[mg setup] dense inverse of %d rows on the device %7.1f ms%s\n", n,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), bad ? " (bad pivot)" : "");
        return ARAP_OK;
    }

    // ---- partitioned mode: this rank's share of the GLOBAL hierarchy (mg_partition.h) -----------------------------
    // Returns ARAP_OK with mg_global == false when the global hierarchy is unusable for a reason every rank sees alike
    // (single level / no dense coarsest level); the caller then falls back to the per-rank hierarchy.
    // ---- partitioned mode: the GLOBAL hierarchy built on this rank's GPU (every rank builds the same one: the kernels are
    // deterministic), then downloaded for the host-side slicing (mg_partition.cpp). Replaces ~16 s of host work per rank at 16M
    // vertices by the cotan CSR + setup kernels on the whole mesh and a download. *built = false: use the host path.
    int build_global_hierarchy_device(MgHierarchyHost &H, bool *built) {
        *built = false;
        if (getenv("ARAP_MG_DEVICE_SETUP") && atoi(getenv("ARAP_MG_DEVICE_SETUP")) == 0) return ARAP_OK;
        const GlobalMesh &gm = global_mesh;
        const int Vg = gm.n_vertices, Fg = gm.n_faces;
        if (Vg <= 0 || Fg <= 0) return ARAP_OK;
        const auto t0 = std::chrono::steady_clock::now();
        DeviceBuffer<int> d_faces, d_count, d_rawptr, d_cursor, d_rawcol, d_unique, d_rowptr, d_colidx, d_owner;
        DeviceBuffer<double> d_rest, d_rawval, d_weight;
        DeviceBuffer<unsigned> d_tag;
        DeviceBuffer<unsigned char> d_free;
        ARAP_CUDA(upload_vector(d_faces, gm.faces, stream));
        ARAP_CUDA(upload_vector(d_rest, gm.rest, stream));
        ARAP_CUDA(upload_vector(d_owner, gm.owner, stream));
        {
            std::vector<unsigned char> is_free((size_t)Vg);
            for (int v = 0; v < Vg; ++v) is_free[(size_t)v] = gm.constrained[(size_t)v] ? 0 : 1;
            ARAP_CUDA(upload_vector(d_free, is_free, stream));
            ARAP_CUDA(cudaStreamSynchronize(stream));
        }
        // cotan weights + CSR of the global mesh: the same kernels as arap_prepare's computeCotanWeights (arap.h:182-239)
        ARAP_CUDA(d_count.ensure((size_t)Vg + 1));
        ARAP_CUDA(d_rawptr.ensure((size_t)Vg + 1));
        ARAP_CUDA(d_cursor.ensure((size_t)Vg + 1));
        ARAP_CUDA(d_unique.ensure((size_t)Vg + 1));
        ARAP_CUDA(d_rowptr.ensure((size_t)Vg + 1));
        ARAP_CUDA(d_rawcol.ensure(6 * (size_t)Fg));
        ARAP_CUDA(d_rawval.ensure(6 * (size_t)Fg));
        ARAP_CUDA(d_tag.ensure(6 * (size_t)Fg));
        ARAP_CUDA(cudaMemsetAsync(d_count.ptr, 0, sizeof(int) * ((size_t)Vg + 1), stream));
        ARAP_CUDA(cudaMemsetAsync(d_cursor.ptr, 0, sizeof(int) * ((size_t)Vg + 1), stream));
        weights_count_kernel<<<grid_for((size_t)Fg), kBlock, 0, stream>>>(d_faces.ptr, Fg, d_count.ptr);
        { int rc = exclusive_scan(d_count.ptr, Vg, d_rawptr.ptr); if (rc) return rc; }
        weights_fill_kernel<double><<<grid_for((size_t)Fg), kBlock, 0, stream>>>(d_faces.ptr, Fg, d_rest.ptr, d_rawptr.ptr, d_cursor.ptr, d_rawcol.ptr, d_rawval.ptr, d_tag.ptr);
        ARAP_CUDA(cudaMemsetAsync(d_cursor.ptr + Vg, 0, sizeof(int), stream));
        row_sort_merge_kernel<double><<<grid_for((size_t)Vg), kBlock, 0, stream>>>(Vg, d_rawptr.ptr, d_rawcol.ptr, d_rawval.ptr, d_tag.ptr, d_unique.ptr, d_cursor.ptr,
                                                                                   d_cursor.ptr + Vg);
        row_sort_long_kernel<double><<<sm_count, kBlock, 0, stream>>>(d_cursor.ptr, d_cursor.ptr + Vg, d_rawptr.ptr, d_rawcol.ptr, d_rawval.ptr, d_tag.ptr, d_unique.ptr);
        { int rc = exclusive_scan(d_unique.ptr, Vg, d_rowptr.ptr); if (rc) return rc; }
        int g_nnz = 0;
        ARAP_CUDA(cudaMemcpyAsync(&g_nnz, d_rowptr.ptr + Vg, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(d_colidx.ensure((size_t)g_nnz));
        ARAP_CUDA(d_weight.ensure((size_t)g_nnz));
        csr_compact_kernel<double><<<grid_for((size_t)Vg), kBlock, 0, stream>>>(Vg, d_rawptr.ptr, d_rawcol.ptr, d_rawval.ptr, d_rowptr.ptr, d_colidx.ptr, d_weight.ptr);
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        d_rawcol.release(); d_rawval.release(); d_tag.release(); d_faces.release();
        // the hierarchy, aggregates confined to one owner each
        const MgSetupOptions mo = engine_mg_options();
        std::vector<std::unique_ptr<DevLevel>> dl;
        double complexity = 0;
        bool ok = false;
        { int rc = build_hierarchy_device<double>(Vg, d_rowptr.ptr, d_colidx.ptr, d_weight.ptr, d_free.ptr, d_owner.ptr, mo, dl, &complexity, &ok,
                                                      d_rest.ptr, 8, 3); if (rc) return rc; }
        if (!ok) return ARAP_OK;
        // download: mg_slice_hierarchy works on host matrices
        H.levels.clear();
        H.levels.resize(dl.size());
        for (size_t l = 0; l < dl.size(); ++l) {
            DevLevel &src = *dl[l];
            MgLevelHost &hl = H.levels[l];
            { int rc = download_csr(*src.A, hl.A); if (rc) return rc; }
            if (src.P) { int rc = download_csr(*src.P, hl.P); if (rc) return rc; }
            if (src.R) { int rc = download_csr(*src.R, hl.R); if (rc) return rc; }
            hl.inv_diag.resize((size_t)src.n);
            hl.block.resize((size_t)src.n);
            ARAP_CUDA(cudaMemcpyAsync(hl.inv_diag.data(), src.inv_diag.ptr, sizeof(double) * (size_t)src.n, cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaMemcpyAsync(hl.block.data(), src.block.ptr, sizeof(int) * (size_t)src.n, cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaStreamSynchronize(stream));
            hl.omega = src.omega;
            dl[l].reset();                                      // free this level's device copy as soon as it is on the host
        }
        H.n_coarse = H.levels.back().A.n_rows;
        H.coarse_inv.clear();
        H.coarse_dense_on_device = H.n_coarse <= mo.max_dense;
        H.operator_complexity = complexity;
        stats.setup_device_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *built = true;
        return ARAP_OK;
    }

    int setup_multigrid_global() {
        auto t0 = std::chrono::steady_clock::now();
        mg_global = false;
        const GlobalMesh &gm = global_mesh;
        MgHierarchyHost H;
        bool on_device = false;
        { int rc = build_global_hierarchy_device(H, &on_device); if (rc) return rc; }
        if (!on_device) {
            std::vector<int> g_rowptr, g_colidx, visit;
            std::vector<double> g_w;
            build_global_csr(gm.n_vertices, gm.n_faces, gm.faces.data(), gm.rest.data(), g_rowptr, g_colidx, g_w);
            morton_sequence(gm.n_vertices, gm.rest.data(), visit);
            MgSetupOptions mo = engine_mg_options();
            mg_build_hierarchy<double>(gm.n_vertices, g_rowptr.data(), g_colidx.data(), g_w.data(), gm.constrained.data(), mo, H,
                                       visit.data(), gm.owner.data());
        }
        if (H.levels.size() < 2 || (H.coarse_inv.empty() && !H.coarse_dense_on_device)) return ARAP_OK;
        const int V = n_vertices;
        std::vector<int> h_perm((size_t)V), global_of_local((size_t)V);
        if (V > 0) ARAP_CUDA(cudaMemcpyAsync(h_perm.data(), perm.ptr, sizeof(int) * (size_t)V, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < V; ++i) global_of_local[(size_t)i] = gm.local_to_global[(size_t)h_perm[(size_t)i]];
        MgLocalHierarchy LH;
        std::string err;
        // levels of at most this many rows are replicated on every rank (mg_partition.h): 4 halo exchanges less per level
        // a level kept whole on every rank costs its full sweep time on each of them plus an all-reduce of its right-hand side; a
        // partitioned one costs 1 / world of the sweeps plus four halo exchanges (~65 us). Measured break-even at 16M vertices:
        // the 300k-row level is better replicated on 2 GPUs and better partitioned on 8 (8.4 -> 7.5 ms per ARAP iteration).
        const int replicate_rows = getenv("ARAP_MG_REPLICATE_ROWS") ? atoi(getenv("ARAP_MG_REPLICATE_ROWS")) : std::max(50000, 800000 / std::max(1, world_size));
        if (!mg_slice_hierarchy(H, my_rank, n_rows, V, global_of_local.data(), LH, err, replicate_rows)) return fail(ARAP_ERR_SOLVER, err);
        mg_first_replicated = LH.first_replicated;
        { MgHierarchyHost().levels.swap(H.levels); }          // the global matrices are no longer needed
        mg.clear();
        std::vector<float> fscratch;
        size_t max_send = 1;
        const size_t L = LH.levels.size();
        for (size_t l = 0; l < L; ++l) {
            const MgLocalLevel &hl = LH.levels[l];
            std::unique_ptr<MgLevelDev> d(new MgLevelDev());
            d->n = (l == 0) ? V : hl.n_own;                  // level 0 vectors are indexed like every other per-vertex array
            d->n_ext = (l == 0) ? V : hl.n_own + hl.n_halo;
            d->omega = hl.omega;
            const size_t len = (size_t)(d->n_ext > 0 ? d->n_ext : 1);
            if (l > 0 && l + 1 < L) {
                ARAP_CUDA(upload_vector(d->a_rowptr, hl.A.rowptr, stream));
                ARAP_CUDA(upload_vector(d->a_colidx, hl.A.colidx, stream));
                ARAP_CUDA(upload_as_float(d->a_val, hl.A.val, stream, fscratch));
                d->plan = hl.plan;
                ARAP_CUDA(upload_vector(d->send_index, hl.plan.send_index, stream));
                max_send = std::max(max_send, (size_t)hl.plan.n_send());
                d->a_lanes = pick_lanes(hl.A.colidx.size(), (size_t)hl.A.n_rows);
            }
            if (l > 0) {
                ARAP_CUDA(d->b.ensure(len));
                ARAP_CUDA(cudaMemsetAsync(d->b.ptr, 0, sizeof(MgVec) * len, stream));
                ARAP_CUDA(upload_as_float(d->inv_diag, hl.inv_diag, stream, fscratch));
            }
            if (l + 1 < L) {
                ARAP_CUDA(upload_vector(d->p_rowptr, hl.P.rowptr, stream));
                ARAP_CUDA(upload_vector(d->p_colidx, hl.P.colidx, stream));
                ARAP_CUDA(upload_as_float(d->p_val, hl.P.val, stream, fscratch));
                ARAP_CUDA(upload_vector(d->r_rowptr, hl.R.rowptr, stream));
                ARAP_CUDA(upload_vector(d->r_colidx, hl.R.colidx, stream));
                ARAP_CUDA(upload_as_float(d->r_val, hl.R.val, stream, fscratch));
                ARAP_CUDA(d->r.ensure(len));
                ARAP_CUDA(cudaMemsetAsync(d->r.ptr, 0, sizeof(MgVec) * len, stream));
                d->r_lanes = pick_lanes(hl.R.colidx.size(), (size_t)hl.R.n_rows);
            }
            ARAP_CUDA(d->x.ensure(len));
            ARAP_CUDA(d->x2.ensure(len));
            ARAP_CUDA(cudaMemsetAsync(d->x.ptr, 0, sizeof(MgVec) * len, stream));
            ARAP_CUDA(cudaMemsetAsync(d->x2.ptr, 0, sizeof(MgVec) * len, stream));
            mg.push_back(std::move(d));
        }
        ARAP_CUDA(mg_sendbuf.ensure(max_send * sizeof(MgVec)));
        mg_dense = true;
        if (LH.coarse_dense_on_device) {
            int rc = invert_coarsest_on_device(LH.coarse_A);
            if (rc) return rc;
            if (!mg_dense) return fail(ARAP_ERR_SOLVER, "global multigrid: the coarsest operator could not be inverted");
        } else {
            int rc = upload_coarse_inverse(LH.coarse_inv, LH.n_coarse);
            if (rc) return rc;
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        stats.mg_levels = (int)mg.size();
        stats.mg_operator_complexity = LH.operator_complexity;
        stats.setup_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        mg_global = true;
        return ARAP_OK;
    }

    // The V(1,1) cycle of vcycle() on a hierarchy whose rows are spread over the ranks: same kernels on the owned rows,
    // with the halo of every gathered vector refreshed right before the kernel that gathers it.
    int vcycle_partitioned() {
        const int R = n_rows;
        const int L = (int)mg.size();
        const int Lr = mg_first_replicated;                      // levels >= Lr are whole on every rank: no exchanges there
        MgLevelDev &m0 = *mg[0];
        MgVec *z = m0.x2.ptr;
        // down
        { int rc = exchange_halo(m0.x.ptr, sizeof(MgVec), SITE_MG0_X_PRE); if (rc) return rc; }      // x0 = omega D^-1 r was made on owned rows
        launch_fine_residual();
        { int rc = exchange_halo(m0.r.ptr, sizeof(MgVec), SITE_MG0_R); if (rc) return rc; }
        for (int l = 0; l + 1 < L; ++l) {
            MgLevelDev &f = *mg[l], &c = *mg[l + 1];
            ARAP_DISPATCH_LANES(f.r_lanes, LAUNCH_PDL(ARAP_K_MG_RESTRICT, mg_restrict_presmooth_kernel<LN>, grid_for((size_t)c.n * LN), c.n,
                                                      f.r_rowptr.ptr, f.r_colidx.ptr, f.r_val.ptr, f.r.ptr, c.inv_diag.ptr, (float)c.omega, c.b.ptr,
                                                      c.x.ptr, cg.ptr));
            if (l + 1 == Lr) {
                // into the replicated part: every rank restricted the rows it owns (zeros elsewhere); sum them, then redo the
                // pre-smoothing x = omega D^-1 b with the complete right-hand side
                if (comm_counting) comm_allreduces += 1;
                begin_launch(ARAP_K_ALLREDUCE_LEVEL);
                const int rc_red = transport->allreduce_sum_f32(stream, SITE_COARSE_B, (float *)c.b.ptr, 4 * c.n);
                end_launch();
                if (rc_red) return fail(ARAP_ERR_CUDA, transport->error);
                pdl_next_plain = true;
                if (l + 1 < L - 1)
                    LAUNCH_PDL(ARAP_K_MG_RESTRICT, mg_jacobi_kernel, grid_for((size_t)c.n), c.n, c.inv_diag.ptr, (float)c.omega, c.b.ptr, c.x.ptr);
            }
            if (l + 1 == L - 1) break;
            if (l + 1 < Lr) { int rc = exchange_level(l + 1, 0, c.x.ptr); if (rc) return rc; }
            ARAP_DISPATCH_LANES(c.a_lanes, LAUNCH_PDL(ARAP_K_MG_CSR_RESIDUAL, mg_csr_residual_kernel<LN>, grid_for((size_t)c.n * LN), c.n,
                                                      c.a_rowptr.ptr, c.a_colidx.ptr, c.a_val.ptr, c.b.ptr, c.x.ptr, c.r.ptr, cg.ptr));
            if (l + 1 < Lr) { int rc = exchange_level(l + 1, 1, c.r.ptr); if (rc) return rc; }
        }
        // coarsest: solved redundantly on every rank
        MgLevelDev &cl = *mg[L - 1];
        launch_dense_solve(cl.n, cl.b.ptr, cl.x2.ptr);
        // up
        for (int l = L - 2; l >= 0; --l) {
            MgLevelDev &f = *mg[l], &c = *mg[l + 1];
            const int rows = (l == 0) ? R : f.n;
            if (l + 1 < Lr) { int rc = exchange_level(l + 1, 2, c.x2.ptr); if (rc) return rc; }
            LAUNCH_PDL(ARAP_K_MG_PROLONG, mg_prolong_add_kernel, grid_for((size_t)rows), rows, f.p_rowptr.ptr, f.p_colidx.ptr, f.p_val.ptr,
                       c.x2.ptr, f.x.ptr, cg.ptr);
            if (l == 0) {
                { int rc = exchange_halo(f.x.ptr, sizeof(MgVec), SITE_MG0_X_POST); if (rc) return rc; }
                launch_fine_postsmooth();
            } else {
                if (l < Lr) { int rc = exchange_level(l, 3, f.x.ptr); if (rc) return rc; }
                ARAP_DISPATCH_LANES(f.a_lanes, LAUNCH_PDL(ARAP_K_MG_CSR_POSTSMOOTH, mg_csr_postsmooth_kernel<LN>, grid_for((size_t)f.n * LN), f.n,
                                                          f.a_rowptr.ptr, f.a_colidx.ptr, f.a_val.ptr, f.inv_diag.ptr, (float)f.omega, f.b.ptr,
                                                          f.x.ptr, f.x2.ptr, cg.ptr));
            }
        }
        return ARAP_OK;      // gamma = r.z and the position-error norm sit in cg->red until CG_STAGE_MERGED
    }

    // The coarse part of the V-cycle as one kernel. Two forms (ARAP_TAIL): "grid" (default) -- every level below the fine one,
    // with the restriction out of and the prolongation into the fine level, in one COOPERATIVE kernel that fills the GPU and
    // separates its phases with a grid barrier (mg_tail_grid_kernel); "cluster" -- the round-1 experiment (levels of at most
    // ARAP_TAIL_ROWS rows in one thread-block cluster; measured slower than separate launches); "0" -- one launch per phase.
    DeviceBuffer<GridBarrier> tail_barrier;
    bool tail_grid = false;
    int tail_grid_ctas = 0;
    size_t tail_grid_smem = 0;
    int plan_tail() {
        tail_first = 0;
        tail_grid = false;
        const int L = (int)mg.size();
        const char *env_mode = getenv("ARAP_TAIL");
        const std::string mode = env_mode ? env_mode : "0";          // measured at 1M vertices: 105 us for the cooperative kernel vs ~55 us of graph-replayed launches
        if (L < 2 || mg_global || transport || mode == "0") return ARAP_OK;
        int t = 0;
        if (mode == "grid") {
            if (L > kTailMaxLevels) return ARAP_OK;
            t = 1;
        } else {
            const char *env_rows = getenv("ARAP_TAIL_ROWS"), *env_parent = getenv("ARAP_TAIL_PARENT_ROWS"), *env_cluster = getenv("ARAP_TAIL_CLUSTER");
            const int max_rows = env_rows ? atoi(env_rows) : 0, max_parent = env_parent ? atoi(env_parent) : 65536;
            tail_cluster = env_cluster ? atoi(env_cluster) : 8;
            if (max_rows <= 0) return ARAP_OK;
            for (int l = 1; l < L; ++l) if (mg[(size_t)l]->n <= max_rows) { t = l; break; }
            if (t == 0 || mg[(size_t)t - 1]->n > max_parent || L - t + 1 > kTailMaxLevels) return ARAP_OK;
        }
        std::memset(&tail_args, 0, sizeof(tail_args));
        tail_args.n_levels = L - t + 1;
        tail_args.dense = mg_dense ? 1 : 0;
        tail_args.coarse_inv = mg_coarse_inv.ptr;
        tail_args.coarse_ld = mg_coarse_ld;
        for (int l = t - 1; l < L; ++l) {
            MgLevelDev &d = *mg[(size_t)l];
            MgTailLevel &a = tail_args.lv[l - (t - 1)];
            a.n = (l == 0) ? n_rows : d.n;
            a.a_lanes = d.a_lanes; a.r_lanes = d.r_lanes; a.omega = (float)d.omega;
            a.a_rowptr = d.a_rowptr.ptr; a.a_colidx = d.a_colidx.ptr; a.a_val = d.a_val.ptr; a.inv_diag = d.inv_diag.ptr;
            a.p_rowptr = d.p_rowptr.ptr; a.p_colidx = d.p_colidx.ptr; a.p_val = d.p_val.ptr;
            a.r_rowptr = d.r_rowptr.ptr; a.r_colidx = d.r_colidx.ptr; a.r_val = d.r_val.ptr;
            a.b = d.b.ptr; a.x = d.x.ptr; a.x2 = d.x2.ptr; a.r = d.r.ptr;
        }
        if (mode == "grid") {
            int coop = 0;
            ARAP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
            if (!coop) return ARAP_OK;
            tail_grid_smem = mg_dense ? sizeof(MgVec) * (size_t)mg_coarse_ld : 0;
            if (tail_grid_smem > 48 * 1024) ARAP_CUDA(cudaFuncSetAttribute(mg_tail_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mg_tail_grid_kernel, kBlock, tail_grid_smem) != cudaSuccess || occ <= 0) { cudaGetLastError(); return ARAP_OK; }
            const char *env_occ = getenv("ARAP_TAIL_CTAS_PER_SM");
            if (env_occ && atoi(env_occ) > 0 && atoi(env_occ) < occ) occ = atoi(env_occ);
            tail_grid_ctas = sm_count * occ;
            ARAP_CUDA(tail_barrier.ensure(1));
            ARAP_CUDA(cudaMemsetAsync(tail_barrier.ptr, 0, sizeof(GridBarrier), stream));
            tail_grid = true;
        } else if (tail_cluster > 8) {
            ARAP_CUDA(cudaFuncSetAttribute(mg_tail_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        tail_first = t;
        return ARAP_OK;
    }

    int launch_tail() {
        if (tail_grid) {
            const CgScalars *cgp = cg.ptr;
            GridBarrier *bar = tail_barrier.ptr;
            void *params[3] = {(void *)&tail_args, (void *)&cgp, (void *)&bar};
            pdl_next_plain = true;
            begin_launch(ARAP_K_MG_TAIL);
            const cudaError_t e = cudaLaunchCooperativeKernel((const void *)mg_tail_grid_kernel, dim3((unsigned)tail_grid_ctas, 1, 1), dim3(kBlock, 1, 1), params,
                                                              tail_grid_smem, stream);
            end_launch();
            pdl_next_plain = true;
            if (e != cudaSuccess) return fail(ARAP_ERR_CUDA, std::string("mg_tail_grid_kernel launch: ") + cudaGetErrorString(e));
            return ARAP_OK;
        }
        cudaLaunchConfig_t cfg;
        cfg = cudaLaunchConfig_t();
        cfg.gridDim = dim3((unsigned)tail_cluster, 1, 1);
        cfg.blockDim = dim3(kTailThreads, 1, 1);
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)tail_cluster;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        begin_launch(ARAP_K_MG_TAIL);
        cudaError_t e = cudaLaunchKernelEx(&cfg, mg_tail_kernel, tail_args, (const CgScalars *)cg.ptr);
        end_launch();
        if (e != cudaSuccess) return fail(ARAP_ERR_CUDA, std::string("mg_tail_kernel launch: ") + cudaGetErrorString(e));
        return ARAP_OK;
    }

    void launch_dense_solve(int n, const MgVec *b, MgVec *x) {
        begin_launch(ARAP_K_MG_DENSE_SOLVE);
        launch_pdl(mg_dense_solve_kernel, (unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock), (unsigned)kBlock, sizeof(MgVec) * (size_t)mg_coarse_ld,
                   n, mg_coarse_ld, (const float *)mg_coarse_inv.ptr, b, x, (const CgScalars *)cg.ptr);
        end_launch();
    }

    // z = M^-1 r by one V(1,1) cycle; the last kernel also produces rho = r.z and beta.
    // Buffer roles are fixed (no pointer swapping) so that the launch sequence can be captured in a CUDA graph:
    // on every level x = iterate before post-smoothing, x2 = the level's result; level 0's result is z = mg[0]->x2.
    int vcycle() {
        if (mg_global) return vcycle_partitioned();
        const int R = n_rows;                 // owned rows (== n_vertices on a single GPU)
        const int L = (int)mg.size();
        MgLevelDev &m0 = *mg[0];
        MgVec *z = m0.x2.ptr;
        if (L == 1) {
            // tiny meshes: the whole system is the "coarsest level"; b = the fp64 CG residual converted to fp32
            if (mg_batch_dense) {       // Z = Inv . R over all members at once, straight from the fp64 residual
                begin_launch(ARAP_K_MG_DENSE_SOLVE);
                if (mg_batch_tc) {
                    const int n_mt = (member_vertices + kTcM - 1) / kTcM, n_nt = (batch_members + kTcMembers - 1) / kTcMembers;
                    const int n_ks = (member_vertices + kTcK - 1) / kTcK;
                    const size_t items = (size_t)n_nt * n_ks * kTcKcores * kTcN;
                    batch_pack_b_kernel<<<grid_for(items), kBlock, 0, stream>>>(member_vertices, batch_members, n_nt, n_ks, cg_r.ptr, mg_b_pack.ptr, cg.ptr);
                    mg_batch_dense_tc_kernel<<<dim3((unsigned)n_mt, (unsigned)n_nt, 1), 128, kTcSmemBytes, stream>>>(member_vertices, batch_members, n_ks,
                                                                                                                  mg_a_pack.ptr, mg_b_pack.ptr, z, cg.ptr);
                } else {
                    const dim3 grid((unsigned)((member_vertices + kBgM - 1) / kBgM), (unsigned)((batch_members + kBgMembers - 1) / kBgMembers), 1);
                    mg_batch_dense_kernel<<<grid, 256, 0, stream>>>(member_vertices, mg_coarse_ld, batch_members, mg_coarse_inv.ptr, cg_r.ptr, z, cg.ptr);
                }
                end_launch();
                LAUNCH_PDL(ARAP_K_CG_DOT, cg_dot_rho_f_kernel, reduce_grid(cg_dot_rho_f_kernel, (size_t)R), R, cg_r.ptr, z, partials.ptr, counter.ptr, cg.ptr);
                return ARAP_OK;      // gamma = r.z and the position-error norm sit in cg (one GPU) / cg->red (partitioned: CG_STAGE_MERGED)
            }
            LAUNCH(ARAP_K_MISC, mg_to_float_kernel, grid_for((size_t)m0.n), m0.n, cg_r.ptr, m0.x.ptr);
            if (mg_dense) {
                launch_dense_solve(m0.n, m0.x.ptr, z);
            } else {      // no dense inverse (singular coarse operator): plain Jacobi, z = omega D^-1 r
                LAUNCH(ARAP_K_MISC, mg_jacobi_kernel, grid_for((size_t)m0.n), m0.n, m0.inv_diag.ptr, (float)m0.omega, m0.x.ptr, z);
            }
            LAUNCH_PDL(ARAP_K_CG_DOT, cg_dot_rho_f_kernel, reduce_grid(cg_dot_rho_f_kernel, (size_t)R), R, cg_r.ptr, z, partials.ptr, counter.ptr, cg.ptr);
            return ARAP_OK;      // gamma = r.z and the position-error norm sit in cg (one GPU) / cg->red (partitioned: CG_STAGE_MERGED)
        }
        // down
        // levels [tail_first, L) run inside mg_tail_kernel, which also restricts into and prolongates out of them
        const int top = tail_first > 0 ? tail_first - 1 : L - 1;      // the coarsest level handled launch by launch
        for (int l = 0; l <= top && l + 1 < L; ++l) {
            MgLevelDev &f = *mg[l], &c = *mg[l + 1];
            if (l == 0) {
                launch_fine_residual();
            } else {
                ARAP_DISPATCH_LANES(f.a_lanes, LAUNCH_PDL(ARAP_K_MG_CSR_RESIDUAL, mg_csr_residual_kernel<LN>, grid_for((size_t)f.n * LN), f.n,
                                                      f.a_rowptr.ptr, f.a_colidx.ptr, f.a_val.ptr, f.b.ptr, f.x.ptr, f.r.ptr, cg.ptr));
            }
            if (l == top) break;                                       // the tail kernel restricts out of this level itself
            ARAP_DISPATCH_LANES(f.r_lanes, LAUNCH_PDL(ARAP_K_MG_RESTRICT, mg_restrict_presmooth_kernel<LN>, grid_for((size_t)c.n * LN), c.n,
                                                  f.r_rowptr.ptr, f.r_colidx.ptr, f.r_val.ptr, f.r.ptr, c.inv_diag.ptr, (float)c.omega, c.b.ptr,
                                                  c.x.ptr, cg.ptr));
        }
        // coarsest: exact dense solve, or one more damped-Jacobi step when the level is too large for a dense inverse
        MgLevelDev &cl = *mg[L - 1];
        if (tail_first > 0) {
            int rc = launch_tail();
            if (rc) return rc;
        } else if (mg_dense) {
            launch_dense_solve(cl.n, cl.b.ptr, cl.x2.ptr);
        } else {
            ARAP_DISPATCH_LANES(cl.a_lanes, LAUNCH_PDL(ARAP_K_MG_CSR_POSTSMOOTH, mg_csr_postsmooth_kernel<LN>, grid_for((size_t)cl.n * LN), cl.n,
                                                   cl.a_rowptr.ptr, cl.a_colidx.ptr, cl.a_val.ptr, cl.inv_diag.ptr, (float)cl.omega, cl.b.ptr,
                                                   cl.x.ptr, cl.x2.ptr, cg.ptr));
        }
        // up
        for (int l = (tail_first > 0 ? top : L - 2); l >= 0; --l) {
            MgLevelDev &f = *mg[l], &c = *mg[l + 1];
            const int rows = (l == 0) ? R : f.n;
            if (!(tail_first > 0 && l == top))                         // the tail kernel already prolongated into its parent
                LAUNCH_PDL(ARAP_K_MG_PROLONG, mg_prolong_add_kernel, grid_for((size_t)rows), rows, f.p_rowptr.ptr, f.p_colidx.ptr, f.p_val.ptr,
                       c.x2.ptr, f.x.ptr, cg.ptr);
            if (l == 0) {
                launch_fine_postsmooth();
            } else {
                ARAP_DISPATCH_LANES(f.a_lanes, LAUNCH_PDL(ARAP_K_MG_CSR_POSTSMOOTH, mg_csr_postsmooth_kernel<LN>, grid_for((size_t)f.n * LN), f.n,
                                                      f.a_rowptr.ptr, f.a_colidx.ptr, f.a_val.ptr, f.inv_diag.ptr, (float)f.omega, f.b.ptr,
                                                      f.x.ptr, f.x2.ptr, cg.ptr));
            }
        }
        return ARAP_OK;      // gamma = r.z and the position-error norm sit in cg (one GPU) / cg->red (partitioned: CG_STAGE_MERGED)
    }

    bool use_tma = getenv("ARAP_TMA") != nullptr && atoi(getenv("ARAP_TMA")) != 0;
    void launch_spmv() {
        const int R = n_rows;
        if (use_tma) {
            const size_t smem = tma_spmv_smem_bytes<S>();
            begin_launch(ARAP_K_CG_SPMV);
            cg_spmv_tma_kernel<S><<<reduce_grid(cg_spmv_tma_kernel<S>, (size_t)R, smem), kBlock, smem, stream>>>(
                R, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, free_mask.ptr, cg_d.ptr, cg_ad.ptr, partials.ptr, counter.ptr, cg.ptr);
            end_launch();
        } else {
            LAUNCH_PDL(ARAP_K_CG_SPMV, cg_spmv_kernel<S>, reduce_grid(cg_spmv_kernel<S>, (size_t)R), R, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, free_mask.ptr,
                   cg_d.ptr, cg_ad.ptr, partials.ptr, counter.ptr, cg.ptr);
        }
    }

    int cg_iteration_jacobi() {
        const int R = n_rows;
        { int rc = exchange_halo(cg_d.ptr, sizeof(Vec3d), SITE_CG_D); if (rc) return rc; }
        launch_spmv();
        { int rc = reduce_stage(CG_STAGE_ALPHA, 3); if (rc) return rc; }
        LAUNCH_PDL(ARAP_K_CG_UPDATE, cg_update_kernel, reduce_grid(cg_update_kernel, (size_t)R), R, inv_diag.ptr, cg_d.ptr, cg_ad.ptr, cg_x.ptr, cg_r.ptr, partials.ptr,
               counter.ptr, cg.ptr);
        { int rc = reduce_stage(CG_STAGE_UPDATE_JACOBI, 4); if (rc) return rc; }
        LAUNCH_PDL(ARAP_K_CG_DIRECTION, cg_direction_kernel, grid_for((size_t)R), R, inv_diag.ptr, cg_r.ptr, cg_d.ptr, cg.ptr);
        return ARAP_OK;
    }

    // One iteration of the multigrid-preconditioned CG in its single-reduction form (kernels.cuh, CgStage):
    //   z = V-cycle(r) [gamma = r.z, position-error norm]; halo of z; w = A z [delta = z.w]; ONE all-reduce (partitioned);
    //   d = z + beta d, s = w + beta s, x += alpha d, r -= alpha s, x0 = omega D^-1 r for the next V-cycle [|r|^2].
    // `loop`: the conditional handle of the step graph's WHILE node when this is captured as its body, else 0.
    int cg_iteration_mg(unsigned long long loop = 0) {
        const int R = n_rows;
        const int n3 = 3 * R;
        MgLevelDev &m0 = *mg[0];
        { int rc = vcycle(); if (rc) return rc; }
        { int rc = exchange_halo(m0.x2.ptr, sizeof(MgVec), SITE_CG_D); if (rc) return rc; }
        launch_spmv_z();
        { int rc = reduce_stage(CG_STAGE_MERGED, 8); if (rc) return rc; }
        LAUNCH_PDL(ARAP_K_CG_UPDATE_MG, cg_fused_update_kernel, reduce_grid(cg_fused_update_kernel, ((size_t)n3 + 1) / 2), n3, inv_diag.ptr, m0.omega,
                   (const float *)m0.x2.ptr, (const double *)cg_w.ptr, (double *)cg_d.ptr, (double *)cg_ad.ptr, (double *)cg_x.ptr, (double *)cg_r.ptr,
                   (float *)m0.x.ptr, partials.ptr, counter.ptr, cg.ptr, loop);
        return ARAP_OK;
    }

    // Capture one CG iteration into a CUDA graph: ~20 small launches collapse into one graph launch.
    int build_cg_graph() {
        destroy_cg_graph();
        std::memset(graph_counts, 0, sizeof(graph_counts));
        ARAP_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        capturing = true;
        pdl_next_plain = true;
        comm_count_begin();
        const int rc_body = use_mg ? cg_iteration_mg() : cg_iteration_jacobi();
        comm_count_end();
        capturing = false;
        pdl_next_plain = true;
        cudaError_t e = cudaStreamEndCapture(stream, &cg_graph);
        if (e != cudaSuccess || rc_body) { cg_graph = nullptr; return fail(ARAP_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e)); }
        std::memcpy(graph_counts_iter, graph_counts, sizeof(graph_counts_iter));
        ARAP_CUDA(cudaGraphInstantiate(&cg_graph_exec, cg_graph, 0));
        return ARAP_OK;
    }
    int64_t graph_counts_iter[ARAP_K_COUNT_MAX] = {0};

    // ---- the whole ARAP iteration as ONE graph with a device-side loop ---------------------------------------------------
    //   local step, redo list, right-hand side + residual   ->   WHILE (not converged) { one CG iteration }   ->   p' += x
    // The WHILE node's condition is set on the device by the kernel that ends the CG iteration (cg_fused_update_kernel), so a
    // global step needs no host round trip at all: arap_iterate(n) is n graph launches and one synchronisation at the end
    // (reference loop: arap.h:122-129). Single GPU + multigrid solver; everything else keeps the host-driven loop below.
    cudaGraph_t step_graph = nullptr;
    cudaGraphExec_t step_graph_exec = nullptr;
    int64_t step_counts[ARAP_K_COUNT_MAX] = {0}, body_counts[ARAP_K_COUNT_MAX] = {0};
    void destroy_step_graph() {
        if (step_graph_exec) cudaGraphExecDestroy(step_graph_exec);
        if (step_graph) cudaGraphDestroy(step_graph);
        step_graph_exec = nullptr;
        step_graph = nullptr;
    }

    void launch_step_head() {
        const int R = n_rows, G = grid_for((size_t)R);
        MgLevelDev &m0 = *mg[0];
        // quat[] starts as identity (initializeRotations), which is already a usable Newton seed: the hot kernel
        // certifies convergence to the SVD's rotation per vertex and lists the vertices that need the Jacobi SVD.
        launch_local_step();
        
        launch_rhs_residual<true>(mg[0]->omega, mg[0]->x.ptr);
    }

    int build_step_graph() {
        destroy_step_graph();
        if (transport || !use_mg || mg.empty()) return ARAP_OK;
        if (getenv("ARAP_STEP_GRAPH") && atoi(getenv("ARAP_STEP_GRAPH")) == 0) return ARAP_OK;
        const int R = n_rows, G = grid_for((size_t)R);
        cudaGraphConditionalHandle handle = 0;
        cudaGraph_t body = nullptr, captured = nullptr;
        bool ok = cudaGraphCreate(&step_graph, 0) == cudaSuccess &&
                  cudaGraphConditionalHandleCreate(&handle, step_graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
        // head: local step + right-hand side
        if (ok) ok = cudaStreamBeginCaptureToGraph(stream, step_graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            capturing = true;
            std::memset(graph_counts, 0, sizeof(graph_counts));
            pdl_next_plain = true;
            launch_step_head();
            // the WHILE node hangs off whatever the capture currently ends in; the capture then continues behind it
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            const cudaGraphNode_t *deps = nullptr;
            const cudaGraphEdgeData *edges = nullptr;
            size_t n_deps = 0;
            ok = cudaStreamGetCaptureInfo_v3(stream, &st, nullptr, nullptr, &deps, &edges, &n_deps) == cudaSuccess && st == cudaStreamCaptureStatusActive;
            cudaGraphNode_t cond = nullptr;
            if (ok) {
                cudaGraphNodeParams np = {};
                np.type = cudaGraphNodeTypeConditional;
                np.conditional.handle = handle;
                np.conditional.type = cudaGraphCondTypeWhile;
                np.conditional.size = 1;
                std::vector<cudaGraphNode_t> dep_nodes(deps, deps + n_deps);
                ok = cudaGraphAddNode(&cond, step_graph, dep_nodes.data(), n_deps, &np) == cudaSuccess && np.conditional.phGraph_out != nullptr;
                if (ok) body = np.conditional.phGraph_out[0];
            }
            if (ok) ok = cudaStreamUpdateCaptureDependencies(stream, &cond, 1, cudaStreamSetCaptureDependencies) == cudaSuccess;
            if (ok) {
                pdl_next_plain = true;                                  // no programmatic edge out of a conditional node
                LAUNCH_PDL(ARAP_K_APPLY, apply_update_kernel<S>, G, R, free_mask.ptr, cg_x.ptr, cur4.ptr, cg.ptr, 1);
            }
            std::memcpy(step_counts, graph_counts, sizeof(step_counts));
            capturing = false;
            pdl_next_plain = true;
            const cudaError_t e = cudaStreamEndCapture(stream, &captured);
            ok = ok && e == cudaSuccess;
        }
        // body: one CG iteration, closed by the kernel that sets the loop condition
        if (ok) ok = cudaStreamBeginCaptureToGraph(stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            capturing = true;
            std::memset(graph_counts, 0, sizeof(graph_counts));
            pdl_next_plain = true;
            const int rc_body = cg_iteration_mg((unsigned long long)handle);
            std::memcpy(body_counts, graph_counts, sizeof(body_counts));
            capturing = false;
            pdl_next_plain = true;
            cudaGraph_t body_out = nullptr;
            const cudaError_t e = cudaStreamEndCapture(stream, &body_out);
            ok = e == cudaSuccess && rc_body == ARAP_OK;
        }
        if (ok) ok = cudaGraphInstantiate(&step_graph_exec, step_graph, 0) == cudaSuccess;
        if (!ok) {                                                       // not fatal: the host-driven loop does the same work
            cudaGetLastError();
            destroy_step_graph();
        }
        return ARAP_OK;
    }

    inline int issue_cg_iteration() {
        if (cg_graph_exec && !profile_events) {
            for (int k = 0; k < ARAP_K_COUNT_MAX; ++k) profile.launches[k] += graph_counts_iter[k];
            ARAP_CUDA(cudaGraphLaunch(cg_graph_exec, stream));
        } else {
            const bool count = transport && !comm_counted;
            if (count) comm_count_begin();
            const int rc = use_mg ? cg_iteration_mg() : cg_iteration_jacobi();
            if (count) comm_count_end();
            return rc;
        }
        return ARAP_OK;
    }

    int comm_benchmark(int rounds, double *us_exchange, double *us_allreduce) override {
        if (!transport || !prepared) return fail(ARAP_ERR_INVALID, "comm_benchmark: needs a prepared partitioned handle");
        if (rounds <= 0) rounds = 1;
        double *red = (double *)((char *)cg.ptr + offsetof(CgScalars, red));
        float ms = 0.f;
        for (int pass = 0; pass < 2; ++pass) {                   // pass 0 warms up
            ARAP_CUDA(cudaEventRecord(timer_start, stream));
            for (int k = 0; k < rounds; ++k) {
                int rc;
                if (use_mg) rc = exchange_halo(mg[0]->x2.ptr, sizeof(MgVec), SITE_CG_D);
                else rc = exchange_halo(cg_d.ptr, sizeof(Vec3d), SITE_CG_D);
                if (rc) return rc;
            }
            ARAP_CUDA(cudaEventRecord(timer_stop, stream));
            ARAP_CUDA(cudaEventSynchronize(timer_stop));
            ARAP_CUDA(cudaEventElapsedTime(&ms, timer_start, timer_stop));
        }
        if (us_exchange) *us_exchange = 1e3 * (double)ms / rounds;
        ARAP_CUDA(cudaMemsetAsync(red, 0, 8 * sizeof(double), stream));
        for (int pass = 0; pass < 2; ++pass) {
            ARAP_CUDA(cudaEventRecord(timer_start, stream));
            for (int k = 0; k < rounds; ++k)
                if (transport->allreduce_sum(stream, SITE_RED_BASE + CG_STAGE_MERGED, red, 8)) return fail(ARAP_ERR_CUDA, transport->error);
            ARAP_CUDA(cudaEventRecord(timer_stop, stream));
            ARAP_CUDA(cudaEventSynchronize(timer_stop));
            ARAP_CUDA(cudaEventElapsedTime(&ms, timer_start, timer_stop));
        }
        if (us_allreduce) *us_allreduce = 1e3 * (double)ms / rounds;
        pdl_next_plain = true;
        if (transport->poll_error()) return fail(ARAP_ERR_CUDA, transport->error);
        return ARAP_OK;
    }

    // Host-driven global step (partitioned mode, Jacobi-PCG, profiling passes, or when the step graph is unavailable).
    int global_step() {
        const int R = n_rows, G = grid_for((size_t)R);
        if (use_mg) {
            MgLevelDev &m0 = *mg[0];
            launch_rhs_residual<true>(mg[0]->omega, mg[0]->x.ptr);
            { int rc = reduce_stage(CG_STAGE_START_MG, 5); if (rc) return rc; }
        } else {
            launch_rhs_residual<false>(1.0, (float4 *)nullptr);
            { int rc = reduce_stage(CG_STAGE_START_JACOBI, 5); if (rc) return rc; }
        }
        const int max_it = opt.max_cg_iterations > 0 ? opt.max_cg_iterations : 20000;
        int check = opt.cg_check_interval > 0 ? opt.cg_check_interval : 32;
        if (use_mg) check = 1;
        // The iteration count of a warm-started solve is very close to the previous global step's, so
        // that many iterations (minus a margin) are enqueued blind; after that, batches of `check`
        // iterations are enqueued one batch ahead of the convergence poll so the device never idles
        // waiting for the host. Once converged every kernel returns immediately.
        int issued = 0, slot = 0;
        bool have_poll[2] = {false, false};
        bool done = false;
        int blind = stats.last_cg_iterations - (use_mg ? 1 : check);
        if (blind > max_it) blind = max_it;
        for (; issued < blind; ++issued) { int rc = issue_cg_iteration(); if (rc) return rc; }
        while (!done) {
            const int batch = (max_it - issued < check) ? (max_it - issued) : check;
            for (int it = 0; it < batch; ++it) { int rc = issue_cg_iteration(); if (rc) return rc; }
            issued += batch;
            ARAP_CUDA(cudaMemcpyAsync(&cg_host[slot], cg.ptr, sizeof(CgScalars), cudaMemcpyDeviceToHost, stream));
            ARAP_CUDA(cudaEventRecord(poll_event[slot], stream));
            pdl_next_plain = true;
            have_poll[slot] = true;
            const int prev = slot ^ 1;
            if (have_poll[prev]) {
                ARAP_CUDA(cudaEventSynchronize(poll_event[prev]));
                if (cg_host[prev].converged) done = true;
            }
            if (issued >= max_it || batch == 0) done = true;
            slot ^= 1;
        }
        LAUNCH_PDL(ARAP_K_APPLY, apply_update_kernel<S>, G, R, free_mask.ptr, cg_x.ptr, cur4.ptr, cg.ptr, 1);
        { int rc = exchange_halo(cur4.ptr, sizeof(Vec4T<S>), SITE_CUR4); if (rc) return rc; }
        ARAP_CUDA(cudaMemcpyAsync(&cg_host[0], cg.ptr, sizeof(CgScalars), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        pdl_next_plain = true;
        if (transport && transport->poll_error()) return fail(ARAP_ERR_CUDA, transport->error);
        ARAP_CUDA(cudaGetLastError());
        stats.last_cg_iterations = cg_host[0].iterations;
        return ARAP_OK;
    }

    // Fold the device-side bookkeeping of the global steps since the last call (CgScalars::steps ...) into `stats`.
    // cg_host[0] must hold a copy of *cg taken after the last step; the device counters are reset for the next batch.
    int book_steps() {
        const CgScalars &c = cg_host[0];
        if (c.steps > 0) {
            if (use_mg) {
                if (mg_fresh && mg_fresh_iterations == 0) mg_fresh_iterations = c.first_step_iterations > 0 ? c.first_step_iterations : 1;
                else if (!mg_fresh && mg_fresh_iterations > 0 && c.iterations > (3 * mg_fresh_iterations) / 2 + 3) mg_stale = true;
            }
            stats.global_steps += c.steps;
            stats.cg_iterations_total += c.iterations_total;
            stats.last_cg_iterations = c.iterations;
            stats.last_converged = c.converged;
            unconverged_steps += c.unconverged_steps;
            stats.last_relative_residual = c.ref2 > 0 ? sqrt(c.rr / c.ref2) : 0.0;
            stats.last_position_error = c.z8_tol > 0 ? std::pow(c.z8, 0.125) : 0.0;
            if (step_graph_exec && !profile_events)                  // launches the step graphs made (body runs: one per CG iteration +
                for (int k = 0; k < ARAP_K_COUNT_MAX; ++k)           // the pass that notices convergence)
                    profile.launches[k] += (int64_t)c.steps * step_counts[k] + c.body_runs_total * body_counts[k];
        }
        ARAP_CUDA(cudaMemsetAsync((char *)cg.ptr + offsetof(CgScalars, iterations_total), 0, sizeof(CgScalars) - offsetof(CgScalars, iterations_total), stream));
        if (!(c.rr == c.rr)) return fail(ARAP_ERR_SOLVER, "global step: CG residual is NaN");
        return ARAP_OK;
    }
    int unconverged_steps = 0;

    // `defer_sync`: leave the final synchronisation to the caller (arap_deform: the write-back's copy synchronises anyway);
    // finish_iterate() must follow once the stream is known to be idle.
    bool iterate_pending = false;
    int iterate(int n, bool defer_sync = false) override {
        if (!prepared) return fail(ARAP_ERR_INVALID, "iterate: arap_prepare has not succeeded");
        if (dirty) return fail(ARAP_ERR_INVALID, "iterate: the handle is dirty (constraints changed since the last arap_prepare): call arap_prepare or arap_deform");
        const int R = n_rows, G = grid_for((size_t)R);
        unconverged_steps = 0;
        if (n <= 0) return ARAP_OK;
        if (step_graph_exec && !profile_events) {
            for (int it = 0; it < n; ++it) ARAP_CUDA(cudaGraphLaunch(step_graph_exec, stream));
            ARAP_CUDA(cudaMemcpyAsync(&cg_host[0], cg.ptr, sizeof(CgScalars), cudaMemcpyDeviceToHost, stream));
            pdl_next_plain = true;
            iterate_pending = true;
            if (defer_sync) return ARAP_OK;
            return finish_iterate();
        }
        for (int it = 0; it < n; ++it) {
            launch_local_step();
            
            { int rc = exchange_halo(quat.ptr, sizeof(Vec4T<S>), SITE_QUAT); if (rc) return rc; }
            int rc = global_step();
            if (rc) return rc;
        }
        ARAP_CUDA(cudaGetLastError());
        { int rc = book_steps(); if (rc) return rc; }
        return unconverged_steps > 0 ? ARAP_NOT_CONVERGED : ARAP_OK;
    }

    int finish_iterate() override {
        if (!iterate_pending) return ARAP_OK;
        iterate_pending = false;
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        { int rc = book_steps(); if (rc) return rc; }
        return unconverged_steps > 0 ? ARAP_NOT_CONVERGED : ARAP_OK;
    }

    int get_positions(void *out, int scalar_bytes) override {
        if (!out || (scalar_bytes != 4 && scalar_bytes != 8)) return fail(ARAP_ERR_INVALID, "get_positions: bad arguments");
        if (!cur4.ptr) return fail(ARAP_ERR_INVALID, "get_positions: no state (call arap_prepare first)");
        const int V = n_vertices;
        const size_t bytes = (size_t)scalar_bytes * 3 * (size_t)V;
        ARAP_CUDA(staging.ensure(bytes));
        begin_launch(ARAP_K_MISC);
        if (scalar_bytes == 4) export_positions_kernel<S, float><<<grid_for((size_t)V), kBlock, 0, stream>>>(V, perm.ptr, cur4.ptr, (float *)staging.ptr);
        else export_positions_kernel<S, double><<<grid_for((size_t)V), kBlock, 0, stream>>>(V, perm.ptr, cur4.ptr, (double *)staging.ptr);
        end_launch();
        ARAP_CUDA(cudaMemcpyAsync(out, staging.ptr, bytes, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }

    // ---- pipelined deform: the write-back of frame k runs on a second stream while frame k + 1 iterates --------------------
    // deform_async enqueues n iterations, a snapshot of p' in the caller's order and scalar type, and the snapshot's copy to the
    // caller's (page-locked) buffer on `copy_stream`; nothing blocks the host. deform_wait blocks until the OLDEST frame in flight
    // is in its buffer. At most two frames are in flight (two host buffers on the caller's side).
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t snap_ready[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
    DeviceBuffer<unsigned char> snapshot;
    unsigned async_head = 0, async_tail = 0;       // frames enqueued / frames waited for
    int async_status = ARAP_OK;

    int deform_async(void *out, int scalar_bytes, int n) override {
        if (!out || (scalar_bytes != 4 && scalar_bytes != 8)) return fail(ARAP_ERR_INVALID, "deform_async: bad arguments");
        if (async_head - async_tail >= 2) return fail(ARAP_ERR_INVALID, "deform_async: two frames are in flight already (call arap_deform_wait)");
        if (!copy_stream) {
            ARAP_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
            for (int k = 0; k < 2; ++k) {
                ARAP_CUDA(cudaEventCreateWithFlags(&snap_ready[k], cudaEventDisableTiming));
                ARAP_CUDA(cudaEventCreateWithFlags(&copy_done[k], cudaEventDisableTiming));
            }
        }
        const int rc = iterate(n, /*defer_sync=*/true);
        if (rc < 0) return rc;
        if (rc > 0) async_status = rc;
        const int V = n_vertices;
        const size_t bytes = (size_t)scalar_bytes * 3 * (size_t)V;
        const unsigned slot = async_head & 1u;
        if (async_head > 0) ARAP_CUDA(cudaStreamWaitEvent(stream, copy_done[(async_head - 1) & 1u], 0));   // the snapshot is still being read
        ARAP_CUDA(snapshot.ensure(bytes));
        begin_launch(ARAP_K_MISC);
        if (scalar_bytes == 4) export_positions_kernel<S, float><<<grid_for((size_t)V), kBlock, 0, stream>>>(V, perm.ptr, cur4.ptr, (float *)snapshot.ptr);
        else export_positions_kernel<S, double><<<grid_for((size_t)V), kBlock, 0, stream>>>(V, perm.ptr, cur4.ptr, (double *)snapshot.ptr);
        end_launch();
        pdl_next_plain = true;
        ARAP_CUDA(cudaEventRecord(snap_ready[slot], stream));
        ARAP_CUDA(cudaStreamWaitEvent(copy_stream, snap_ready[slot], 0));
        ARAP_CUDA(cudaMemcpyAsync(out, snapshot.ptr, bytes, cudaMemcpyDeviceToHost, copy_stream));
        ARAP_CUDA(cudaEventRecord(copy_done[slot], copy_stream));
        ++async_head;
        return ARAP_OK;
    }
    bool deform_wait_pending() override { return async_head != async_tail; }
    int deform_wait() override {
        if (async_head == async_tail) return ARAP_OK;
        ARAP_CUDA(cudaEventSynchronize(copy_done[async_tail & 1u]));
        ++async_tail;
        if (async_head != async_tail) return ARAP_OK;          // the convergence status is booked when the pipeline has drained
        const int rc = finish_iterate();
        const int st = async_status;
        async_status = ARAP_OK;
        if (rc < 0) return rc;
        return rc > 0 ? rc : st;
    }

    int get_csr_nnz(int *out) override {
        if (!rowptr.ptr) return fail(ARAP_ERR_INVALID, "get_csr: weights not built (call arap_prepare first)");
        *out = nnz;
        return ARAP_OK;
    }
    int get_csr(int *rp, int *ci, void *w) override {
        if (!rowptr.ptr) return fail(ARAP_ERR_INVALID, "get_csr: weights not built (call arap_prepare first)");
        ARAP_CUDA(cudaMemcpyAsync(rp, rowptr.ptr, sizeof(int) * ((size_t)n_vertices + 1), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaMemcpyAsync(ci, colidx.ptr, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaMemcpyAsync(w, weight.ptr, sizeof(S) * (size_t)nnz, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }
    int get_free_map(int *fi, int *nf) override {
        if (!free_idx.ptr) return fail(ARAP_ERR_INVALID, "get_free_map: call arap_prepare first");
        if (fi) ARAP_CUDA(cudaMemcpyAsync(fi, free_idx.ptr, sizeof(int) * (size_t)n_vertices, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        if (nf) *nf = n_free;
        return ARAP_OK;
    }
    int get_rotations(void *rot9) override {
        if (!quat.ptr) return fail(ARAP_ERR_INVALID, "get_rotations: call arap_prepare first");
        const int V = n_vertices;
        ARAP_CUDA(staging.ensure(sizeof(S) * 9 * (size_t)V));
        LAUNCH(ARAP_K_MISC, export_rotations_kernel<S>, grid_for((size_t)V), V, perm.ptr, quat.ptr, (S *)staging.ptr);
        ARAP_CUDA(cudaMemcpyAsync(rot9, staging.ptr, sizeof(S) * 9 * (size_t)V, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }
    // _b (arap.h:393-414) for the current rotations, nF x 3 doubles in free-index order: runs the hot right-hand-side kernel
    // exactly as a global step would and rebuilds b from its output (export_rhs_kernel). Inspection / tests only.
    int get_rhs(double *out) override {
        if (!prepared || dirty) return fail(ARAP_ERR_INVALID, "get_rhs: call arap_prepare first");
        if (transport) return fail(ARAP_ERR_INVALID, "get_rhs: not available on a partitioned handle");
        const int R = n_rows;
        pdl_next_plain = true;
        if (use_mg) {
            MgLevelDev &m0 = *mg[0];
            launch_rhs_residual<true>(mg[0]->omega, mg[0]->x.ptr);
        } else {
            launch_rhs_residual<false>(1.0, (float4 *)nullptr);
        }
        pdl_next_plain = true;
        ARAP_CUDA(staging.ensure(sizeof(double) * 3 * (size_t)(n_free > 0 ? n_free : 1)));
        LAUNCH(ARAP_K_MISC, export_rhs_kernel<S>, grid_for((size_t)R), R, perm.ptr, free_idx.ptr, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr,
               free_mask.ptr, cur4.ptr, cg_r.ptr, (double *)staging.ptr);
        ARAP_CUDA(cudaMemcpyAsync(out, staging.ptr, sizeof(double) * 3 * (size_t)n_free, cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        return ARAP_OK;
    }
    // Viewer interop (reference examples/osg_viewer.cpp:45-72): float positions + vertex normals of the current pose, into host
    // memory or straight into a device buffer of the caller (location = ARAP_BUFFER_DEVICE: e.g. a mapped OpenGL VBO).
    DeviceBuffer<int> vf_ptr, vf_face, vf_cursor;
    DeviceBuffer<float4> face_normal;
    DeviceBuffer<float> render_staging;
    bool vf_built = false;
    int get_render_buffers(float *positions, float *normals, int location) override {
        if (!positions || (location != ARAP_BUFFER_HOST && location != ARAP_BUFFER_DEVICE)) return fail(ARAP_ERR_INVALID, "get_render_buffers: bad arguments");
        if (!cur4.ptr || !have_perm) return fail(ARAP_ERR_INVALID, "get_render_buffers: no state (call arap_prepare first)");
        if (transport) return fail(ARAP_ERR_INVALID, "get_render_buffers: not available on a partitioned handle");
        const int V = n_vertices, F = n_faces;
        if (normals && !vf_built) {
            ARAP_CUDA(vf_ptr.ensure((size_t)V + 1));
            ARAP_CUDA(vf_cursor.ensure((size_t)V + 1));
            ARAP_CUDA(vf_face.ensure(3 * (size_t)F));
            ARAP_CUDA(face_normal.ensure((size_t)F));
            ARAP_CUDA(cudaMemsetAsync(vf_cursor.ptr, 0, sizeof(int) * ((size_t)V + 1), stream));
            if (F > 0) LAUNCH(ARAP_K_MISC, vf_count_kernel, grid_for((size_t)F), F, faces.ptr, vf_cursor.ptr);
            { int rc = exclusive_scan(vf_cursor.ptr, V, vf_ptr.ptr); if (rc) return rc; }
            ARAP_CUDA(cudaMemsetAsync(vf_cursor.ptr, 0, sizeof(int) * ((size_t)V + 1), stream));
            if (F > 0) LAUNCH(ARAP_K_MISC, vf_fill_kernel, grid_for((size_t)F), F, faces.ptr, vf_ptr.ptr, vf_cursor.ptr, vf_face.ptr);
            if (V > 0) LAUNCH(ARAP_K_MISC, vf_sort_kernel, grid_for((size_t)V), V, vf_ptr.ptr, vf_face.ptr);
            vf_built = true;
        }
        float *d_pos = positions, *d_nrm = normals;
        if (location == ARAP_BUFFER_HOST) {
            ARAP_CUDA(render_staging.ensure(6 * (size_t)(V > 0 ? V : 1)));
            d_pos = render_staging.ptr;
            d_nrm = normals ? render_staging.ptr + 3 * (size_t)V : nullptr;
        }
        if (normals && F > 0) LAUNCH(ARAP_K_MISC, face_normals_kernel<S>, grid_for((size_t)F), F, faces.ptr, iperm.ptr, cur4.ptr, face_normal.ptr);
        if (V > 0) LAUNCH(ARAP_K_MISC, render_buffers_kernel<S>, grid_for((size_t)V), V, iperm.ptr, cur4.ptr, vf_ptr.ptr, vf_face.ptr, face_normal.ptr, d_pos, d_nrm);
        if (location == ARAP_BUFFER_HOST) {
            ARAP_CUDA(cudaMemcpyAsync(positions, d_pos, sizeof(float) * 3 * (size_t)V, cudaMemcpyDeviceToHost, stream));
            if (normals) ARAP_CUDA(cudaMemcpyAsync(normals, d_nrm, sizeof(float) * 3 * (size_t)V, cudaMemcpyDeviceToHost, stream));
        }
        ARAP_CUDA(cudaStreamSynchronize(stream));
        ARAP_CUDA(cudaGetLastError());
        return ARAP_OK;
    }
    int energy(double *e) override {
        if (!quat.ptr || !hot_rowptr.ptr) return fail(ARAP_ERR_INVALID, "energy: call arap_prepare first");
        const int V = prepared ? n_rows : n_vertices;      // partitioned mode: this rank's share (owned rows)
        LAUNCH(ARAP_K_ENERGY, energy_kernel<S>, reduce_grid(energy_kernel<S>, (size_t)V), V, hot_rowptr.ptr, hot_colidx.ptr, hot_weight.ptr, rest4.ptr, cur4.ptr, quat.ptr,
               partials.ptr, counter.ptr, energy_dev.ptr);
        ARAP_CUDA(cudaMemcpyAsync(e, energy_dev.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream));
        ARAP_CUDA(cudaStreamSynchronize(stream));
        return ARAP_OK;
    }
};

}  // namespace arap

// =================================================================================================
// C ABI
// =================================================================================================
struct arap_handle {
    arap::EngineBase *engine;
    int precision_bytes;
};

extern "C" {

int arap_abi_version(void) { return ARAP_B200_ABI_VERSION; }

void arap_default_options(arap_options *opt) {
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->struct_size = (int32_t)sizeof(arap_options);
    opt->device = -1;
    opt->solver = ARAP_SOLVER_AUTO;
    opt->max_cg_iterations = 20000;
    opt->cg_tolerance = 0.0;   /* 0 = per-solver default: 1e-6 (multigrid), 1e-9 (Jacobi) */
    opt->cg_check_interval = 32;
    opt->profile = 0;
    opt->position_tolerance = 0.0;   /* 0 = default (multigrid, when cg_tolerance is not given) */
}

const char *arap_create_error(void) { return arap::g_create_error.c_str(); }

int arap_create(const int32_t *faces, int32_t n_faces, int32_t n_vertices, int32_t precision_bytes, const arap_options *opt,
                arap_handle **out) {
    if (!out) return ARAP_ERR_INVALID;
    *out = nullptr;
    if (n_faces < 0 || n_vertices < 0 || (n_faces > 0 && !faces) || (precision_bytes != 4 && precision_bytes != 8)) {
        arap::g_create_error = "arap_create: bad arguments";
        return ARAP_ERR_INVALID;
    }
    for (size_t k = 0; k < 3 * (size_t)n_faces; ++k)
        if (faces[k] < 0 || faces[k] >= n_vertices) {
            arap::g_create_error = "arap_create: face references a vertex index out of range";
            return ARAP_ERR_INVALID;
        }
    arap_options o;
    arap_default_options(&o);
    if (opt) {
        size_t sz = opt->struct_size > 0 ? (size_t)opt->struct_size : sizeof(arap_options);
        if (sz > sizeof(arap_options)) sz = sizeof(arap_options);
        std::memcpy(&o, opt, sz);
        o.struct_size = (int32_t)sizeof(arap_options);
    }
    arap_handle *h = new (std::nothrow) arap_handle;
    if (!h) return ARAP_ERR_ALLOC;
    h->precision_bytes = precision_bytes;
    int rc;
    if (precision_bytes == 4) {
        auto *e = new (std::nothrow) arap::Engine<float>();
        if (!e) { delete h; return ARAP_ERR_ALLOC; }
        rc = e->init(faces, n_faces, n_vertices, o);
        h->engine = e;
    } else {
        auto *e = new (std::nothrow) arap::Engine<double>();
        if (!e) { delete h; return ARAP_ERR_ALLOC; }
        rc = e->init(faces, n_faces, n_vertices, o);
        h->engine = e;
    }
    if (rc != ARAP_OK) {
        arap::g_create_error = h->engine->last_error;
        delete h->engine;
        delete h;
        return rc;
    }
    *out = h;
    return ARAP_OK;
}

void arap_destroy(arap_handle *h) {
    if (!h) return;
    delete h->engine;
    delete h;
}

#define ARAP_ENGINE_OR_FAIL(h) do { if (!(h) || !(h)->engine) return ARAP_ERR_INVALID; (h)->engine->activate(); } while (0)

int arap_set_constraints(arap_handle *h, int32_t n, const int32_t *idx, const void *xyz, int32_t xyz_scalar_bytes) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->set_constraints(n, idx, xyz, xyz_scalar_bytes);
}

int arap_is_dirty(const arap_handle *h) { return (h && h->engine) ? (h->engine->dirty ? 1 : 0) : 1; }

int arap_prepare(arap_handle *h, const void *rest_xyz, int32_t rest_scalar_bytes) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->prepare(rest_xyz, rest_scalar_bytes);
}

int arap_iterate(arap_handle *h, int32_t n) {
    ARAP_ENGINE_OR_FAIL(h);
    if (n < 0) return h->engine->fail(ARAP_ERR_INVALID, "iterate: negative iteration count");
    return h->engine->iterate(n);
}

int arap_get_positions(arap_handle *h, void *out_xyz, int32_t out_scalar_bytes) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->get_positions(out_xyz, out_scalar_bytes);
}

int arap_deform(arap_handle *h, void *mesh_xyz, int32_t mesh_scalar_bytes, int32_t n_iterations) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!mesh_xyz) return h->engine->fail(ARAP_ERR_INVALID, "deform: null mesh");
    if (h->engine->dirty) {                                               /* arap.h:102 */
        int rc = h->engine->prepare(mesh_xyz, mesh_scalar_bytes);
        if (rc != ARAP_OK) return rc;                                     /* UNCONSTRAINED -> true, no write-back (:113-114) */
    }
    int rc = h->engine->iterate(n_iterations, /*defer_sync=*/true);      /* the write-back below synchronises once for both */
    if (rc < 0) return rc;
    const int rc_pos = h->engine->get_positions(mesh_xyz, mesh_scalar_bytes);   /* arap.h:133-135 */
    const int rc_fin = h->engine->finish_iterate();
    if (rc_fin < 0) return rc_fin;
    if (rc_pos != ARAP_OK) return rc_pos;
    return rc_fin > 0 ? rc_fin : rc;                                       /* ARAP_NOT_CONVERGED */
}

int arap_deform_async(arap_handle *h, void *mesh_xyz, int32_t mesh_scalar_bytes, int32_t n_iterations) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!mesh_xyz) return h->engine->fail(ARAP_ERR_INVALID, "deform_async: null mesh");
    if (h->engine->dirty) {                                               /* arap.h:102: the dirty block is synchronous */
        int rc = h->engine->deform_wait();
        while (rc >= 0 && h->engine->deform_wait_pending()) rc = h->engine->deform_wait();
        if (rc < 0) return rc;
        rc = h->engine->prepare(mesh_xyz, mesh_scalar_bytes);
        if (rc != ARAP_OK) return rc;
    }
    return h->engine->deform_async(mesh_xyz, mesh_scalar_bytes, n_iterations);
}
int arap_deform_wait(arap_handle *h) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->deform_wait();
}

int arap_get_csr_nnz(arap_handle *h, int32_t *nnz) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!nnz) return ARAP_ERR_INVALID;
    return h->engine->get_csr_nnz(nnz);
}
int arap_get_csr(arap_handle *h, int32_t *rowptr, int32_t *colidx, void *weights) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!rowptr || !colidx || !weights) return ARAP_ERR_INVALID;
    return h->engine->get_csr(rowptr, colidx, weights);
}
int arap_get_free_map(arap_handle *h, int32_t *free_idx, int32_t *n_free) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->get_free_map(free_idx, n_free);
}
int arap_get_rotations(arap_handle *h, void *rot9) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!rot9) return ARAP_ERR_INVALID;
    return h->engine->get_rotations(rot9);
}
int arap_get_rhs(arap_handle *h, double *rhs) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!rhs) return ARAP_ERR_INVALID;
    return h->engine->get_rhs(rhs);
}
int arap_get_render_buffers(arap_handle *h, float *positions, float *normals, int32_t location) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->get_render_buffers(positions, normals, location);
}
int arap_energy(arap_handle *h, double *energy) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!energy) return ARAP_ERR_INVALID;
    return h->engine->energy(energy);
}
int arap_get_solver_stats(arap_handle *h, arap_solver_stats *out) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!out) return ARAP_ERR_INVALID;
    *out = h->engine->stats;
    return ARAP_OK;
}

int arap_profile_enable(arap_handle *h, int32_t enable) {
    ARAP_ENGINE_OR_FAIL(h);
    h->engine->collect_profile();
    h->engine->profile_events = enable != 0;
    return ARAP_OK;
}
int arap_profile_reset(arap_handle *h) {
    ARAP_ENGINE_OR_FAIL(h);
    h->engine->collect_profile();
    std::memset(&h->engine->profile, 0, sizeof(arap_profile));
    return ARAP_OK;
}
int arap_profile_get(arap_handle *h, arap_profile *out) {
    ARAP_ENGINE_OR_FAIL(h);
    if (!out) return ARAP_ERR_INVALID;
    h->engine->collect_profile();
    *out = h->engine->profile;
    return ARAP_OK;
}
const char *arap_kernel_name(int32_t id) {
    if (id < 0 || id >= ARAP_K_COUNT_MAX || !arap::kKernelNames[id]) return "";
    return arap::kKernelNames[id];
}

int arap_timer_start(arap_handle *h) {
    ARAP_ENGINE_OR_FAIL(h);
    return cudaEventRecord(h->engine->timer_start, h->engine->stream) == cudaSuccess ? ARAP_OK : ARAP_ERR_CUDA;
}
int arap_timer_stop(arap_handle *h, double *ms) {
    ARAP_ENGINE_OR_FAIL(h);
    if (cudaEventRecord(h->engine->timer_stop, h->engine->stream) != cudaSuccess) return ARAP_ERR_CUDA;
    if (cudaEventSynchronize(h->engine->timer_stop) != cudaSuccess) return ARAP_ERR_CUDA;
    float f = 0.f;
    if (cudaEventElapsedTime(&f, h->engine->timer_start, h->engine->timer_stop) != cudaSuccess) return ARAP_ERR_CUDA;
    if (ms) *ms = (double)f;
    return ARAP_OK;
}
int arap_synchronize(arap_handle *h) {
    ARAP_ENGINE_OR_FAIL(h);
    return cudaStreamSynchronize(h->engine->stream) == cudaSuccess ? ARAP_OK : ARAP_ERR_CUDA;
}

// ---- batches: K copies of the mesh as one block-diagonal problem ----------------------------------------------
struct arap_batch {
    arap_handle *handle;
    int n_vertices, n_faces, batch;
};

int arap_batch_create(const int32_t *faces, int32_t n_faces, int32_t n_vertices, int32_t batch_size, int32_t precision_bytes,
                      const arap_options *opt, arap_batch **out) {
    if (!out) return ARAP_ERR_INVALID;
    *out = nullptr;
    if ((n_faces > 0 && !faces)) {
        arap::g_create_error = "arap_batch_create: null face array";
        return ARAP_ERR_INVALID;
    }
    if (batch_size <= 0 || n_vertices < 0 || n_faces < 0 || (long long)batch_size * n_vertices > 2000000000LL ||
        (long long)batch_size * n_faces * 6 > 2000000000LL) {
        arap::g_create_error = "arap_batch_create: bad sizes (batch x vertices must fit int32)";
        return ARAP_ERR_INVALID;
    }
    for (size_t k = 0; k < 3 * (size_t)n_faces; ++k)
        if (faces[k] < 0 || faces[k] >= n_vertices) {
            arap::g_create_error = "arap_batch_create: face references a vertex index out of range";
            return ARAP_ERR_INVALID;
        }
    std::vector<int32_t> all((size_t)3 * n_faces * batch_size);
    for (int b = 0; b < batch_size; ++b)
        for (size_t k = 0; k < 3 * (size_t)n_faces; ++k) all[(size_t)b * 3 * n_faces + k] = faces[k] + b * n_vertices;
    arap_handle *h = nullptr;
    int rc = arap_create(all.data(), n_faces * batch_size, n_vertices * batch_size, precision_bytes, opt, &h);
    if (rc != ARAP_OK) return rc;
    arap_batch *bt = new (std::nothrow) arap_batch;
    if (!bt) { arap_destroy(h); return ARAP_ERR_ALLOC; }
    h->engine->set_batch_layout(batch_size, n_vertices);
    bt->handle = h;
    bt->n_vertices = n_vertices;
    bt->n_faces = n_faces;
    bt->batch = batch_size;
    *out = bt;
    return ARAP_OK;
}

void arap_batch_destroy(arap_batch *b) {
    if (!b) return;
    arap_destroy(b->handle);
    delete b;
}

arap_handle *arap_batch_handle(arap_batch *b) { return b ? b->handle : nullptr; }

int arap_batch_set_constraints(arap_batch *b, int32_t n, const int32_t *vertex_idx, const void *xyz, int32_t xyz_scalar_bytes) {
    if (!b || n < 0 || (n > 0 && (!vertex_idx || !xyz))) return ARAP_ERR_INVALID;
    for (int k = 0; k < n; ++k)
        if (vertex_idx[k] < 0 || vertex_idx[k] >= b->n_vertices) return b->handle->engine->fail(ARAP_ERR_INVALID, "batch set_constraints: vertex index out of range");
    std::vector<int32_t> idx((size_t)n * b->batch);
    for (int m = 0; m < b->batch; ++m)
        for (int k = 0; k < n; ++k) idx[(size_t)m * n + k] = vertex_idx[k] + m * b->n_vertices;
    return arap_set_constraints(b->handle, n * b->batch, idx.data(), xyz, xyz_scalar_bytes);
}

int arap_batch_prepare(arap_batch *b, const void *rest_xyz, int32_t rest_scalar_bytes) {
    if (!b || !rest_xyz || (rest_scalar_bytes != 4 && rest_scalar_bytes != 8)) return ARAP_ERR_INVALID;
    const size_t one = (size_t)3 * b->n_vertices * rest_scalar_bytes;
    std::vector<unsigned char> all(one * b->batch);
    for (int m = 0; m < b->batch; ++m) std::memcpy(all.data() + one * m, rest_xyz, one);
    return arap_prepare(b->handle, all.data(), rest_scalar_bytes);
}

int arap_batch_iterate(arap_batch *b, int32_t n_iterations) { return b ? arap_iterate(b->handle, n_iterations) : ARAP_ERR_INVALID; }

int arap_batch_get_positions(arap_batch *b, void *out_xyz, int32_t out_scalar_bytes) {
    return b ? arap_get_positions(b->handle, out_xyz, out_scalar_bytes) : ARAP_ERR_INVALID;
}

// ---- handle poses: trajectory + rigid constraint front end (host arithmetic shared with inc/deform/trajectory.h) --------
struct arap_trajectory {
    deform::detail::SplineSE3<double> spline;
};

int arap_trajectory_create(arap_trajectory **out) {
    if (!out) return ARAP_ERR_INVALID;
    *out = new (std::nothrow) arap_trajectory;
    return *out ? ARAP_OK : ARAP_ERR_ALLOC;
}
void arap_trajectory_destroy(arap_trajectory *t) { delete t; }

int arap_trajectory_add_key_pose(arap_trajectory *t, const double *pose16) {
    if (!t || !pose16) return ARAP_ERR_INVALID;
    t->spline.addKeyPose(pose16);
    return ARAP_OK;
}

int arap_trajectory_evaluate(arap_trajectory *t, int32_t n, const double *times, double *poses16_out) {
    if (!t || n < 0 || (n > 0 && (!times || !poses16_out))) return ARAP_ERR_INVALID;
    if (t->spline.numberOfKeyPoses() < 4) {
        arap::g_create_error = "arap_trajectory_evaluate: a cubic trajectory needs at least 4 key poses";
        return ARAP_ERR_INVALID;
    }
    for (int k = 0; k < n; ++k) t->spline.pose(times[k], poses16_out + 16 * (size_t)k);
    return ARAP_OK;
}

int arap_rigid_conjugate(const double *origin16, const double *pose16, double *out16) {
    if (!origin16 || !pose16 || !out16) return ARAP_ERR_INVALID;
    deform::detail::conjugate_rigid<double>(origin16, pose16, out16);
    return ARAP_OK;
}

int arap_set_rigid_constraints(arap_handle *h, int32_t n, const int32_t *vertex_idx, const void *rest_xyz, int32_t rest_scalar_bytes,
                               const double *transform16) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->set_rigid_constraints(n, 1, h->engine->n_vertices, vertex_idx, rest_xyz, rest_scalar_bytes, transform16);
}

int arap_batch_set_rigid_constraints(arap_batch *b, int32_t n, const int32_t *vertex_idx, const void *rest_xyz,
                                     int32_t rest_scalar_bytes, const double *transforms16) {
    if (!b) return ARAP_ERR_INVALID;
    arap_handle *h = b->handle;
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->set_rigid_constraints(n, b->batch, b->n_vertices, vertex_idx, rest_xyz, rest_scalar_bytes, transforms16);
}

int arap_comm_unique_id(void *out_bytes, int32_t capacity) {
    if (!out_bytes || capacity < 128) return ARAP_ERR_INVALID;
    std::string err;
    arap::NcclApi *api = arap::NcclApi::get(&err);
    if (!api) { arap::g_create_error = err; return ARAP_ERR_CUDA; }
    arap::NcclApi::UniqueId id;
    if (api->GetUniqueId(&id) != 0) return ARAP_ERR_CUDA;
    std::memcpy(out_bytes, &id, 128);
    return ARAP_OK;
}

int arap_attach_partition(arap_handle *h, const arap_partition_plan *plan, int32_t rank, int32_t world_size, int32_t transport,
                          const void *id, int32_t id_bytes) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->attach_partition(plan, rank, world_size, transport, id, id_bytes);
}

int arap_partition_comm_benchmark(arap_handle *h, int32_t rounds, double *us_per_exchange, double *us_per_allreduce) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->comm_benchmark(rounds, us_per_exchange, us_per_allreduce);
}

int arap_partition_set_global_mesh(arap_handle *h, const arap_global_mesh *g) {
    ARAP_ENGINE_OR_FAIL(h);
    return h->engine->set_global_mesh(g);
}

int arap_host_alloc(size_t bytes, void **out) {
    if (!out) return ARAP_ERR_INVALID;
    *out = nullptr;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? ARAP_OK : ARAP_ERR_CUDA;
}
int arap_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? ARAP_OK : ARAP_ERR_CUDA; }

const char *arap_last_error(const arap_handle *h) { return (h && h->engine) ? h->engine->last_error.c_str() : "invalid handle"; }

}  // extern "C"
