/*
 * arap_b200.h -- C ABI of libarap_b200.so, the B200 (sm_100a) engine behind the header-only C++
 * API of cheind/mesh-deform (reference inc/deform/arap.h).
 *
 * The reference has no FFI/plugin layer: its boundary is the C++ template class
 * deform::AsRigidAsPossibleDeformation<MeshType, PrecisionType> (reference inc/deform/arap.h:49-466),
 * instantiated in the caller's translation unit. This C ABI is what that class's three public
 * members now bind to (see inc/deform/arap.h in this repo and INTEGRATION.md):
 *
 *   reference member (file:line)                         C-ABI entry point(s)
 *   ---------------------------------------------------  -------------------------------------------
 *   ctor + initializeMeshTopology   arap.h:66-70,149-155  arap_create
 *   (implicit destructor)                                 arap_destroy
 *   setConstraint                   arap.h:81-85          arap_set_constraints
 *   deform: the `_dirty` block      arap.h:102-120        arap_prepare   (geometry, cotan weights + CSR,
 *                                                          rotations, free map, constraints, system setup)
 *   deform: the iteration loop      arap.h:122-129        arap_iterate   (estimateRotations + estimatePositions)
 *   deform: write-back              arap.h:133-135        arap_get_positions
 *   deform, whole                   arap.h:101-138        arap_deform    (prepare-if-dirty + iterate + write-back)
 *   PrivateAccessor::cotanWeights   tests/accessor.h:16-22 arap_get_csr_nnz / arap_get_csr
 *
 * Conventions: plain pointers and sizes only; all pointers are HOST pointers owned by the caller
 * and are copied during the call; xyz arrays are V x 3 (x,y,z per vertex) in the handle's
 * precision unless a scalar-size argument says otherwise; indices are int32. Every function
 * returns an int status (ARAP_OK == 0; > 0 informational; < 0 error) and never throws.
 * A handle is not re-entrant; distinct handles may be used from distinct threads.
 * There is no CPU fallback: without a CUDA device every call fails with ARAP_ERR_CUDA.
 */
#ifndef ARAP_B200_H
#define ARAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARAP_B200_ABI_VERSION 1

enum {
    ARAP_OK = 0,
    ARAP_UNCONSTRAINED = 1,   /* prepare(): no vertex is constrained -> nothing to solve (arap.h:113-114) */
    ARAP_NOT_CONVERGED = 2,   /* iterate()/deform(): a global solve hit max_cg_iterations before its stopping rule was met (e.g. a free
                               * component without any constrained vertex makes L singular -- the case in which the reference's
                               * factorisation fails, arap.h:116-117). The iterations still ran and the positions are the best
                               * iterate; the C++ facade maps this to `false`. */
    ARAP_ERR_INVALID = -1,    /* bad argument / call order */
    ARAP_ERR_CUDA = -2,       /* CUDA runtime failure (message in arap_last_error) */
    ARAP_ERR_SOLVER = -3,     /* linear system unusable: the reference's `return false` (arap.h:116-117) */
    ARAP_ERR_ALLOC = -4
};

enum {
    ARAP_SOLVER_AUTO = 0,
    ARAP_SOLVER_PCG_JACOBI = 1,   /* warm-started Jacobi-preconditioned CG, matrix-free on the one-ring CSR */
    ARAP_SOLVER_PCG_MG = 2        /* CG preconditioned by a smoothed-aggregation multigrid V-cycle (AUTO picks this) */
};

typedef struct arap_handle arap_handle;

typedef struct arap_options {
    int32_t struct_size;        /* = sizeof(arap_options); lets the struct grow compatibly */
    int32_t device;             /* CUDA device ordinal, -1 = current device */
    int32_t solver;             /* ARAP_SOLVER_* */
    int32_t max_cg_iterations;  /* per global step; <= 0 -> default */
    double cg_tolerance;        /* stop when |r|_2 <= tol * |rhs|_2 (all three coordinates together); <= 0 -> per-solver default:
                                 * plain Jacobi: 1e-9; multigrid: no residual test (1e-13 floor), position_tolerance decides */
    int32_t cg_check_interval;  /* CG iterations between convergence polls; <= 0 -> default */
    int32_t profile;            /* != 0: time every kernel launch with CUDA events (see arap_profile_*) */
    double position_tolerance;  /* multigrid solver only: stop a global solve when the estimated position error of the
                                 * iterate (the 8-norm over the vertices of z = M^-1 r, M^-1 one V-cycle) is below
                                 * position_tolerance x bounding-box diagonal of the rest pose. This is the default stopping
                                 * rule (<= 0 -> 1e-8 when cg_tolerance is not given; ignored when cg_tolerance > 0 unless set
                                 * explicitly): unlike a residual tolerance it means the same thing on every mesh. */
} arap_options;

/* Fill `opt` with the defaults (and struct_size). */
void arap_default_options(arap_options *opt);

/* ctor: copies the F x 3 face array to the device. precision_bytes: 4 (float) or 8 (double) = PrecisionType. */
int arap_create(const int32_t *faces, int32_t n_faces, int32_t n_vertices, int32_t precision_bytes,
                const arap_options *opt, arap_handle **out);
void arap_destroy(arap_handle *h);

/* setConstraint for n vertices at once: overwrite-or-insert, constraints accumulate, marks the handle dirty.
 * xyz: n x 3 in scalars of xyz_scalar_bytes (4|8), cast to the handle precision like arap.h:83. */
int arap_set_constraints(arap_handle *h, int32_t n, const int32_t *vertex_idx, const void *xyz, int32_t xyz_scalar_bytes);

/* 1 if the next arap_deform would run the dirty block (arap.h:102). */
int arap_is_dirty(const arap_handle *h);

/* The dirty block (arap.h:107-119) from the rest pose `rest_xyz` (V x 3, scalars of rest_scalar_bytes).
 * Returns ARAP_OK, ARAP_UNCONSTRAINED (weights/CSR are still built; handle stays dirty) or an error. */
int arap_prepare(arap_handle *h, const void *rest_xyz, int32_t rest_scalar_bytes);

/* n x (local step, global step) on the device (arap.h:122-129). Requires a successful arap_prepare that is still current:
 * after arap_set_constraints / arap_set_rigid_constraints the handle is dirty and this call fails with ARAP_ERR_INVALID
 * until arap_prepare (or arap_deform) has run again. Returns ARAP_OK, ARAP_NOT_CONVERGED or an error. */
int arap_iterate(arap_handle *h, int32_t n_iterations);

/* Pipelined arap_deform (extension; the reference's deform is synchronous): enqueues the dirty block if needed (that part blocks),
 * n iterations and the write-back of p' into mesh_xyz on a second stream, and returns without waiting. arap_deform_wait blocks until
 * the OLDEST frame in flight has arrived in its buffer; at most two frames may be in flight, each with its own page-locked buffer
 * (arap_host_alloc), so that the write-back of frame k overlaps the iterations of frame k + 1:
 *     arap_deform_async(h, A, 4, 1);  loop { arap_deform_async(h, B, 4, 1); arap_deform_wait(h); use A; swap(A, B); }
 * The status of the iterations (ARAP_NOT_CONVERGED) is reported by the arap_deform_wait that drains the pipeline. */
int arap_deform_async(arap_handle *h, void *mesh_xyz, int32_t mesh_scalar_bytes, int32_t n_iterations);
int arap_deform_wait(arap_handle *h);

/* Current p' (V x 3) cast to scalars of out_scalar_bytes (arap.h:133-135). */
int arap_get_positions(arap_handle *h, void *out_xyz, int32_t out_scalar_bytes);

/* Whole deform() with `mesh_xyz` playing the role of the reference's Mesh& (arap.h:101-138):
 * if dirty, the rest pose is re-read from mesh_xyz; after the iterations p' is written back into mesh_xyz.
 * Returns ARAP_OK (true), ARAP_UNCONSTRAINED (true, nothing written) or an error (false). */
int arap_deform(arap_handle *h, void *mesh_xyz, int32_t mesh_scalar_bytes, int32_t n_iterations);

/* _edgeWeights (arap.h:453): CSR V x V, columns ascending, no diagonal, values in handle precision. */
int arap_get_csr_nnz(arap_handle *h, int32_t *nnz);
int arap_get_csr(arap_handle *h, int32_t *rowptr /*V+1*/, int32_t *colidx /*nnz*/, void *weights /*nnz*/);

/* _freeIdxMap / _numberOfFreeVariables (arap.h:456-457). Either pointer may be NULL. */
int arap_get_free_map(arap_handle *h, int32_t *free_idx /*V*/, int32_t *n_free);

/* _rotations (arap.h:454) as V x 9 row-major 3x3 matrices in handle precision. */
int arap_get_rotations(arap_handle *h, void *rot9);

/* _b (arap.h:393-414: bFixed + sum_j w_ij/2 (R_i + R_j)(p_i - p_j)) for the current rotations: n_free x 3 doubles, row f =
 * the vertex with free index f (arap_get_free_map). Runs the engine's right-hand-side kernel once; for tests / inspection. */
int arap_get_rhs(arap_handle *h, double *rhs /* n_free x 3 */);

/* Viewer interop (the per-frame work of the reference's viewer, examples/osg_viewer.cpp:45-72: update_normals(), then copy
 * positions and normals into float vertex arrays and re-upload them): float positions (V x 3) and -- unless `normals` is NULL --
 * per-vertex normals (V x 3; unit face normals of the current pose summed over the incident faces and normalised, what
 * OpenMesh's update_normals() computes), both made on the device. location = ARAP_BUFFER_HOST: the pointers are host memory;
 * ARAP_BUFFER_DEVICE: they are DEVICE memory of the caller on the handle's device, e.g. an OpenGL vertex buffer mapped with
 * cudaGraphicsResourceGetMappedPointer -- the frame's geometry then never leaves the GPU. */
enum { ARAP_BUFFER_HOST = 0, ARAP_BUFFER_DEVICE = 1 };
int arap_get_render_buffers(arap_handle *h, float *positions, float *normals, int32_t location);

/* ARAP energy sum_i sum_j w_ij |(p'_i-p'_j) - R_i (p_i-p_j)|^2 (Sorkine & Alexa eq. 3; the reference has none). */
int arap_energy(arap_handle *h, double *energy);

/* Statistics of the global solves since the last arap_prepare. */
typedef struct arap_solver_stats {
    int64_t cg_iterations_total;   /* CG iterations summed over all global steps */
    int32_t global_steps;
    int32_t last_cg_iterations;
    double last_relative_residual; /* |r| / |rhs| at the end of the last global step */
    int32_t last_converged;
    int32_t mg_levels;             /* 0 when the Jacobi preconditioner is in use */
    double mg_operator_complexity;
    double setup_host_ms;          /* host time spent building the multigrid hierarchy in the last arap_prepare (0: built on the device) */
    int32_t cg_graph;              /* 1: a CG iteration (with its exchanges, if partitioned) is replayed from a CUDA graph;
                                    * 2: the whole ARAP iteration is one graph whose CG loop runs on the device (WHILE node) */
    int32_t mg_global;             /* 1: partitioned mode with the global hierarchy (arap_partition_set_global_mesh) */
    double last_position_error;    /* multigrid: the estimate the stopping rule used, as a fraction of the bbox diagonal */
    /* partitioned mode: cross-GPU operations of ONE CG iteration (counted while it is captured / first issued) */
    int32_t comm_exchanges_per_cg_iteration;    /* halo exchanges (one per gathered vector and partitioned multigrid level) */
    int32_t comm_allreduces_per_cg_iteration;   /* all-reduces / all-gathers (CG scalars, replicated coarse levels) */
    int64_t comm_halo_bytes_per_cg_iteration;   /* bytes this rank sends in those exchanges */
    int32_t tile_max_halo;         /* > 0: the one-ring kernels stage their neighbourhood through shared memory in tiles of 256 rows;
                                    * this is the largest tile halo (distinct neighbours outside the tile). 0: untiled kernels */
    int32_t renumbered;            /* 1: the engine renumbered the vertices internally (Morton patches) for this handle */
    double setup_device_ms;        /* wall time of the multigrid setup when it ran on the device (setup_host_ms is then 0) */
    int32_t launches_per_cg_iteration;   /* kernels in the captured CG iteration (0 when no graph was built) */
    int32_t reserved1;
} arap_solver_stats;
int arap_get_solver_stats(arap_handle *h, arap_solver_stats *out);

/* Per-kernel profile (filled when options.profile != 0 or after arap_profile_enable(h, 1)). */
enum {
    ARAP_K_WEIGHTS_COUNT = 0, ARAP_K_WEIGHTS_FILL, ARAP_K_ROW_SORT_MERGE, ARAP_K_CSR_COMPACT, ARAP_K_SCAN,
    ARAP_K_INIT_STATE, ARAP_K_DIAGONAL, ARAP_K_LOCAL_STEP, ARAP_K_RHS_RESIDUAL, ARAP_K_CG_SPMV,
    ARAP_K_CG_UPDATE, ARAP_K_CG_DIRECTION, ARAP_K_APPLY, ARAP_K_ENERGY, ARAP_K_MISC,
    ARAP_K_MG_FINE_RESIDUAL, ARAP_K_MG_FINE_POSTSMOOTH, ARAP_K_MG_CSR_RESIDUAL, ARAP_K_MG_RESTRICT, ARAP_K_MG_PROLONG,
    ARAP_K_MG_CSR_POSTSMOOTH, ARAP_K_MG_DENSE_SOLVE, ARAP_K_CG_UPDATE_MG, ARAP_K_CG_DIRECTION_MG, ARAP_K_CG_DOT,
    ARAP_K_HALO_PACK, ARAP_K_CG_FINALIZE, ARAP_K_LOCAL_STEP_REDO, ARAP_K_MG_TAIL, ARAP_K_ALLREDUCE_SCALARS, ARAP_K_ALLREDUCE_LEVEL,
    ARAP_K_COUNT_MAX = 32
};
typedef struct arap_profile {
    int64_t launches[ARAP_K_COUNT_MAX];
    double milliseconds[ARAP_K_COUNT_MAX];   /* only when event timing is enabled */
} arap_profile;
int arap_profile_enable(arap_handle *h, int32_t enable);
int arap_profile_reset(arap_handle *h);
int arap_profile_get(arap_handle *h, arap_profile *out);
const char *arap_kernel_name(int32_t kernel_id);

/* Device-side timing of a span of calls with CUDA events on the handle's stream. */
int arap_timer_start(arap_handle *h);
int arap_timer_stop(arap_handle *h, double *milliseconds);
int arap_synchronize(arap_handle *h);

/* ---- batches of independent deformations of ONE mesh (BASELINE.json configs[3]) ------------------------------
 * K members share topology, rest pose and the SET of constrained vertices; only the targets differ (one pose
 * per key frame). Replaces the "fresh mesh copy + fresh solver per pose" loop of the reference's
 * examples/deform_example.cpp:37-53. All members advance together in one set of kernel launches: the batch is
 * solved as one block-diagonal system (the members' one-rings never touch), so every kernel of the single-mesh
 * path is reused unchanged. Results are per member, laid out [member][vertex][xyz]. */
typedef struct arap_batch arap_batch;
int arap_batch_create(const int32_t *faces, int32_t n_faces, int32_t n_vertices, int32_t batch_size, int32_t precision_bytes,
                      const arap_options *opt, arap_batch **out);
void arap_batch_destroy(arap_batch *b);
/* n constrained vertices (the same in every member); xyz: batch_size x n x 3 targets. */
int arap_batch_set_constraints(arap_batch *b, int32_t n, const int32_t *vertex_idx, const void *xyz, int32_t xyz_scalar_bytes);
int arap_batch_prepare(arap_batch *b, const void *rest_xyz /* V x 3, shared */, int32_t rest_scalar_bytes);
int arap_batch_iterate(arap_batch *b, int32_t n_iterations);
int arap_batch_get_positions(arap_batch *b, void *out_xyz /* batch_size x V x 3 */, int32_t out_scalar_bytes);
/* The underlying handle (statistics, profiling, timers, total energy over all members). */
arap_handle *arap_batch_handle(arap_batch *b);

/* ---- handle poses: DeformationUtil / TrajectorySE3 front end ---------------------------------------------------
 * The reference drives its handles with rigid transforms: DeformationUtil::updateConstraints
 * (reference inc/deform/deformation_util.h:48-57) calls setConstraint(h_i, origin * t * origin^-1 * p0_i) once per
 * handle, with t taken from TrajectorySE3::operator() (reference inc/deform/trajectory.h:61-73). These entry points
 * do the same per CALL instead of per vertex: the handle indices, their rest positions and one 4x4 transform per
 * batch member go to the device, and a kernel writes every member's targets into the constraint table.
 * All transforms are 4x4 ROW-MAJOR arrays of 16 doubles. */
typedef struct arap_trajectory arap_trajectory;
int arap_trajectory_create(arap_trajectory **out);
void arap_trajectory_destroy(arap_trajectory *t);
/* TrajectorySE3::addKeyPose (trajectory.h:55-59). */
int arap_trajectory_add_key_pose(arap_trajectory *t, const double *pose16);
/* TrajectorySE3::operator() at n times in [0,1] (trajectory.h:61-73); needs >= 4 key poses (cubic spline), else
 * ARAP_ERR_INVALID. poses16_out: n x 16. */
int arap_trajectory_evaluate(arap_trajectory *t, int32_t n, const double *times, double *poses16_out);
/* out = origin * pose * origin^-1, origin^-1 as an isometry inverse (deformation_util.h:38,51). */
int arap_rigid_conjugate(const double *origin16, const double *pose16, double *out16);
/* setConstraint(vertex_idx[k], transform * rest_xyz[k]) for k < n (deformation_util.h:51-55); marks the handle dirty. */
int arap_set_rigid_constraints(arap_handle *h, int32_t n, const int32_t *vertex_idx, const void *rest_xyz, int32_t rest_scalar_bytes,
                               const double *transform16);
/* The same for every member of a batch: member m gets transforms16[m] (batch_size x 16) applied to the shared rest_xyz. */
int arap_batch_set_rigid_constraints(arap_batch *b, int32_t n, const int32_t *vertex_idx, const void *rest_xyz,
                                     int32_t rest_scalar_bytes, const double *transforms16);

/* ---- one mesh partitioned over several GPUs (BASELINE.json configs[4]) ------------------------------------------
 * Every rank creates an ordinary handle for ITS local mesh: the vertices it owns first (local indices
 * [0, n_owned)), then its halo (one-ring neighbours owned by other ranks, grouped by owner), and every face that
 * touches an owned vertex. arap_attach_partition (before arap_prepare) tells the engine who owns what; from then on
 * arap_prepare / arap_iterate / arap_get_positions work as usual on the local mesh, exchanging the halo of p', R and
 * the CG direction and all-reducing the CG's dot products through the chosen transport. The reference has no
 * counterpart (single-threaded CPU code); the partition itself is made by the caller (mesh_deform_b200/partition.py). */
enum { ARAP_TRANSPORT_NCCL = 0,        /* one process per GPU; id = the 128 bytes from arap_comm_unique_id, broadcast by the caller */
       ARAP_TRANSPORT_IN_PROCESS = 1,  /* several partitions on one GPU, one host thread per partition; id = an int32 group key */
       /* Halo exchange and reductions by direct stores into the peers' memory over NVLink / NVSwitch (one put kernel and
        * one wait kernel per exchange, no library call in the iteration; see csrc/peer_transport.cuh). */
       ARAP_TRANSPORT_PEER = 2,             /* one process per GPU of ONE node; bootstraps over NCCL (id as for NCCL) + CUDA IPC */
       ARAP_TRANSPORT_PEER_IN_PROCESS = 3   /* several partitions on one GPU in one process (id = an int32 group key) */ };
typedef struct arap_partition_plan {
    int32_t n_owned;                /* owned vertices come first in the local numbering */
    int32_t n_neighbors;
    const int32_t *neighbor_rank;   /* [n_neighbors] */
    const int32_t *send_offset;     /* [n_neighbors + 1] into send_index */
    const int32_t *send_index;      /* owned local indices to send, grouped by neighbour, in the receiver's halo order */
    const int32_t *recv_offset;     /* [n_neighbors + 1]: halo from neighbour k = local indices n_owned + recv_offset[k] ... */
} arap_partition_plan;
int arap_comm_unique_id(void *out_bytes, int32_t capacity /* >= 128 */);
int arap_attach_partition(arap_handle *h, const arap_partition_plan *plan, int32_t rank, int32_t world_size, int32_t transport,
                          const void *id, int32_t id_bytes);

/* Measurement aid (partitioned handles, after arap_prepare; collective: every rank must call it with the same `rounds`):
 * device time per halo exchange of the CG's gathered vector and per all-reduce of the CG scalars, in microseconds, from
 * `rounds` back-to-back operations bracketed by CUDA events on the handle's stream. */
int arap_partition_comm_benchmark(arap_handle *h, int32_t rounds, double *us_per_exchange, double *us_per_allreduce);

/* Optional, after arap_attach_partition and before arap_prepare: the GLOBAL mesh, so that every rank can build the same
 * multigrid hierarchy for the whole mesh (aggregates never straddle two ranks) and keep its share of every level. The
 * V-cycle then exchanges halos level by level and converges like the single-GPU solver; without this call each rank
 * preconditions only its own block (block-Jacobi across ranks), which costs many more CG iterations. Every rank must
 * pass the same mesh, owner array and constrained set; call it again when the rest pose or the constrained set changes.
 * Host cost per rank: a single-GPU multigrid setup of the WHOLE mesh. All arrays are copied. */
typedef struct arap_global_mesh {
    int32_t n_vertices, n_faces;
    const int32_t *faces;             /* n_faces x 3 global vertex ids */
    const void *rest_xyz;             /* n_vertices x 3 */
    int32_t rest_scalar_bytes;        /* 4 | 8 */
    const int32_t *owner;             /* n_vertices: owning rank of every vertex */
    const int32_t *local_to_global;   /* this rank's local vertex -> global id (one per local vertex, halo included) */
    int32_t n_constrained;
    const int32_t *constrained;       /* global ids of ALL constrained vertices, whoever owns them */
} arap_global_mesh;
int arap_partition_set_global_mesh(arap_handle *h, const arap_global_mesh *g);

/* Page-locked host memory for mesh buffers handed to arap_deform / arap_prepare / arap_get_positions
 * (optional: pageable memory works too, pinned memory makes the copies run at full PCIe rate). */
int arap_host_alloc(size_t bytes, void **out);
int arap_host_free(void *ptr);

/* Message of the last error on this handle ("" if none). Valid until the next call on the handle. */
const char *arap_last_error(const arap_handle *h);
/* Message of the last failed arap_create on this thread. */
const char *arap_create_error(void);

int arap_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ARAP_B200_H */
