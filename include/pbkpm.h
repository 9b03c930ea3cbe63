/* pbkpm.h -- C ABI of libpbkpm.so, the B200-native kernel-polynomial-method engine.
 *
 * This is the drop-in boundary for pybinding's KPM hot path.  Two nested reference interfaces are
 * replaced (paths relative to the pybinding source tree):
 *
 *  (1) the compute-strategy plug-in point
 *        kpm::Compute::Interface::moments(MomentsRef, Starter const&, AlgorithmConfig const&,
 *                                         OptimizedHamiltonian const&)      cppcore/include/kpm/Core.hpp:21-38
 *      whose only implementation is DefaultCompute (cppcore/src/kpm/default/Compute.cpp:14-131).
 *      -> the pbk_moments_* entry points: raw (undamped) Chebyshev moments computed on the GPU.
 *
 *  (2) the per-quantity orchestration of kpm::Core              cppcore/src/kpm/Core.cpp:35-156
 *      as bound by the pybind11 module                          cppmodule/src/kpm.cpp:8-125
 *      -> pbk_create / pbk_set_hamiltonian / pbk_scaling_factors / pbk_calc_* / pbk_moments /
 *         pbk_get_stats / pbk_report.
 *
 * Conventions: plain pointers and sizes only; every function returns a pbk_status; the message of
 * the last failure is kept per context (pbk_last_error).  The caller owns all host buffers, which
 * are only read/written during the call; the library owns all device memory.  One context = one
 * CUDA device and one in-flight calculation (calls on the same context are serialised internally);
 * distinct contexts are fully concurrent, matching the reference's re-entrancy contract for
 * `parallel_for` workers (cppmodule/src/parallel.cpp:25-44).  There is NO CPU fallback: without a
 * usable CUDA device pbk_create fails with PBK_CUDA_ERROR.
 *
 * Complex numbers are interleaved (re, im) pairs; `c128` below means double[2] per element.
 */
#ifndef PBKPM_H
#define PBKPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBK_VERSION 100

typedef struct pbk_ctx pbk_ctx;

/* Scalar type of the Hamiltonian = scalar type of the KPM vectors
 * (reference: Hamiltonian variant, cppcore/include/hamiltonian/Hamiltonian.hpp:50-70). */
typedef enum pbk_dtype { PBK_F32 = 0, PBK_C64 = 1, PBK_F64 = 2, PBK_C128 = 3 } pbk_dtype;

/* Status codes. The Python binding maps them to the exception types pybind11 produces for the
 * reference: INVALID_ARGUMENT -> ValueError (std::invalid_argument), RUNTIME/LOGIC -> RuntimeError. */
typedef enum pbk_status {
    PBK_OK = 0,
    PBK_INVALID_ARGUMENT = 1,
    PBK_RUNTIME_ERROR = 2,
    PBK_LOGIC_ERROR = 3,
    PBK_CUDA_ERROR = 4,
    PBK_NCCL_ERROR = 5
} pbk_status;

/* Damping kernels (cppcore/src/kpm/Kernel.cpp:6-49). */
typedef enum pbk_kernel { PBK_JACKSON = 0, PBK_LORENTZ = 1, PBK_DIRICHLET = 2 } pbk_kernel;

/* kpm::Config (cppcore/include/kpm/Config.hpp:23-32) + cppmodule/src/kpm.cpp:8-37 keyword arguments. */
typedef struct pbk_config {
    float min_energy;        /* min == max (e.g. 0, 0) => bounds are found with the Lanczos procedure */
    float max_energy;
    int32_t kernel;          /* pbk_kernel */
    double lambda_value;     /* Lorentz kernel only */
    int32_t optimal_size;    /* AlgorithmConfig::optimal_size: light-cone row slicing for LDOS / Green's */
    int32_t interleaved;     /* accepted for API compatibility; a CPU cache optimisation with no GPU meaning */
    int32_t matrix_format;   /* 0 = CSR, 1 = ELL; accepted for API compatibility, the device layout is always ELL */
    float lanczos_precision; /* percent, default 0.002 */
    int32_t max_batch;       /* max. KPM vectors advanced together in one pass over H (0 = automatic) */
    int32_t locality_tile;   /* full-system runs (DOS, conductivity, moments): sites are relabelled into breadth-first
                                clusters of this many rows (grouped into macro-blocks of 256 clusters) so that gathers
                                stay on-chip; 0 = automatic, < 0 = keep the caller's site order.  Results do not
                                depend on it beyond summation rounding. */
} pbk_config;

/* kpm::Stats (cppcore/include/kpm/Stats.hpp:19-45, cppmodule/src/kpm.cpp:50-66) + GPU counters. */
typedef struct pbk_stats {
    int64_t num_moments;
    int32_t uses_full_system;
    uint64_t nnz;             /* processed non-zeros over all iterations, no optimisation */
    uint64_t opt_nnz;         /* same with light-cone slicing applied */
    uint64_t vec;
    uint64_t opt_vec;
    double multiplier;        /* repeated calculations (num_random / number of LDOS sites) */
    uint64_t matrix_memory;   /* bytes of the device ELL matrix */
    uint64_t vector_memory;   /* bytes of one KPM vector */
    double hamiltonian_time;  /* seconds: scale/reorder/ELL build + upload */
    double moments_time;      /* seconds: starter + recursion + reductions + allreduce (CUDA events) */
    double eps;               /* multiplier * opt_nnz / moments_time (Stats.cpp:49-51) */
    /* GPU-side evidence for bench.py */
    int64_t kernel_launches;  /* CUDA kernels of this library launched by the last calculation */
    int64_t step_launches;    /* of which fused Chebyshev step kernels */
    double step_ms;           /* summed device time of the step kernels (CUDA events on the library stream) */
    double step_bytes;        /* algorithmic bytes of those launches: sum rows*[k(s+4) + 3*R*s] */
    double starter_ms;        /* device time of starter generation */
    double gemm_ms;           /* Kubo-Bastin contraction device time */
    double gemm_flops;
    int64_t h2d_bytes;        /* host->device bytes moved by the last set_hamiltonian + calculation */
    int64_t d2h_bytes;
    int32_t batch;            /* vectors per pass used by the last calculation */
    int32_t num_batches;
    double moments_device_ms; /* device time of the whole moments phase (CUDA events on the library stream, from the
                                 first starter kernel to the moment copy-out, allreduce included) */
    int64_t bulk_launches;    /* step launches that ran the bulk-copy (TMA) staged kernel variant */
    int64_t res_launches;     /* step launches that ran the resident-tile kernel variant (x rows of a tile and its halo in shared memory) */
    int64_t persist_launches; /* recursions run by the persistent kernel (one launch for all steps of one vector; small systems);
                                 step_launches then counts the steps it executed */
    int64_t graph_launches;   /* recursions replayed as one CUDA graph (small, launch-bound systems); their kernels are
                                 still counted in kernel_launches / step_launches */
} pbk_stats;

/* Progress protocol of DefaultCompute (cppcore/src/kpm/default/Compute.cpp:133-145):
 * (-1, total) at start, (delta, total) per finished batch, (total, total) at the end. */
typedef void (*pbk_progress_fn)(int64_t delta, int64_t total, void* user);

/* ---- lifetime ------------------------------------------------------------------------------- */
int pbk_version(void);
/* replaces: cpb::KPM::KPM + kpm::Core::Core (cppcore/src/KPM.cpp:7-8, src/kpm/Core.cpp:17-23) */
int pbk_create(pbk_ctx** out, int device, const pbk_config* config);
void pbk_destroy(pbk_ctx* ctx);
/* Message of the last failed call on `ctx`; with ctx == NULL the last failed pbk_create of this thread. */
const char* pbk_last_error(const pbk_ctx* ctx);
int pbk_set_progress_callback(pbk_ctx* ctx, pbk_progress_fn fn, void* user);
int pbk_device_count(int* count);

/* ---- Hamiltonian / bounds ------------------------------------------------------------------- */
/* replaces: kpm::Core::set_hamiltonian (src/kpm/Core.cpp:25-29). Host CSR (int32 indices, rows sorted
 * by column), copied by the call. */
int pbk_set_hamiltonian(pbk_ctx* ctx, int dtype, int64_t n, const int32_t* indptr,
                        const int32_t* indices, const void* data);
/* replaces: kpm::Bounds (include/kpm/Bounds.hpp:41-70): Lanczos on the device when no range was given. */
int pbk_bounds(pbk_ctx* ctx, double* min_energy, double* max_energy, int32_t* lanczos_loops);
/* replaces: kpm::Core::scaling_factors (include/kpm/Core.hpp:54) */
int pbk_scaling_factors(pbk_ctx* ctx, double* a, double* b);
/* replaces: Kernel::required_num_moments(broadening / a) as used by every Core::* quantity */
int pbk_required_num_moments(pbk_ctx* ctx, double broadening, int32_t* num_moments);
/* context-free kernel helpers: KPMKernel.damping_coefficients / .required_num_moments (cppmodule/src/kpm.cpp:68-74) */
int pbk_kernel_damping(int kernel, double lambda_value, int32_t n, double* out);
int pbk_kernel_required_num_moments(int kernel, double lambda_value, double scaled_broadening, int32_t* out);

/* ---- compute-strategy level: raw (undamped) moments ----------------------------------------- */
/* BatchDiagonalMoments + RandomStarter + BatchAccumulator (Compute.cpp:52-88, Starter.cpp:48-83,
 * Moments.cpp:7-49): mean over `num_random` stochastic vectors of mu_n = <r|T_n(H)|r>.  The vectors are
 * the first `num_random` vectors of the reference's default-seeded MT19937 stream.  With an
 * initialised communicator (pbk_comm_init) the vectors are sharded over the ranks and the result is
 * all-reduced.  out: c128[num_moments]. num_moments must be of the form 4k+2. */
int pbk_moments_dos(pbk_ctx* ctx, int32_t num_moments, int32_t num_random, void* out);
/* BatchDiagonalMoments + UnitStarter + BatchConcatenator (Core.cpp:58-72): one unit vector per index.
 * out: c128[num_moments * nidx], moment-major (out[n * nidx + i]). */
int pbk_moments_ldos(pbk_ctx* ctx, int32_t num_moments, const int32_t* idx, int32_t nidx, void* out);
/* DiagonalMoments / MultiUnitMoments + UnitStarter (Core.cpp:96-117): mu_n = <col_i|T_n(H)|row>.
 * out: c128[ncols * num_moments], index-major (out[i * num_moments + n]). */
int pbk_moments_greens(pbk_ctx* ctx, int32_t num_moments, int32_t row, const int32_t* cols,
                       int32_t ncols, void* out);
/* DenseMatrixMoments x2 + MomentMultiplication (Core.cpp:119-146, Moments.cpp:92-130): the Kubo-Bastin
 * moment matrix mu_mn = 1/R sum_r <r| v_l T_m(H) ... >, i.e. L * R^H averaged over num_random vectors.
 * left/right: float32 site coordinates (length n) defining the two velocity operators.
 * out: c128[num_moments * num_moments], row-major. */
int pbk_moments_kubo(pbk_ctx* ctx, int32_t num_moments, const float* left, const float* right,
                     int32_t num_random, void* out);
/* DiagonalMoments for caller-supplied starter vectors (ConstantStarter, Starter.cpp:9-24), advanced
 * together as one block.  r0: c128[count * n], vector-major, in the original site order.
 * out: c128[num_moments * count], moment-major. Used by the parity tests to inject identical starters. */
int pbk_moments_diagonal(pbk_ctx* ctx, int32_t num_moments, const void* r0, int32_t count, void* out);
/* The first `count` starter vectors of the device random stream (test hook for the MT19937 kernels).
 * out: c128[count * n], vector-major. */
int pbk_random_vectors(pbk_ctx* ctx, int32_t count, void* out);

/* ---- kpm::Core level: damped moments and reconstructed functions ---------------------------- */
/* replaces: Core::moments (Core.cpp:35-56) / KPM.moments. alpha: c128[n]; beta: c128[n] or NULL;
 * op: CSR (c128 values) or op_rows == 0.  out: c128[num_moments] (damped, mu_0 carries the 1/2). */
int pbk_moments(pbk_ctx* ctx, int32_t num_moments, const void* alpha, const void* beta,
                int64_t op_rows, const int32_t* op_indptr, const int32_t* op_indices,
                const void* op_data, void* out);
/* replaces: Core::dos (Core.cpp:74-90). out: double[num_energy]. */
int pbk_calc_dos(pbk_ctx* ctx, const double* energy, int32_t num_energy, double broadening,
                 int32_t num_random, double* out);
/* replaces: Core::ldos (Core.cpp:58-72). out: double[num_energy * nidx], column-major (energy fastest). */
int pbk_calc_ldos(pbk_ctx* ctx, const double* energy, int32_t num_energy, double broadening,
                  const int32_t* idx, int32_t nidx, double* out);
/* replaces: Core::greens_vector (Core.cpp:96-117). out: c128[ncols * num_energy], index-major. */
int pbk_calc_greens(pbk_ctx* ctx, int32_t row, const int32_t* cols, int32_t ncols,
                    const double* energy, int32_t num_energy, double broadening, void* out);
/* replaces: Core::conductivity (Core.cpp:119-150). out: c128[num_mu] (the facade takes the real part). */
int pbk_calc_conductivity(pbk_ctx* ctx, const float* left, const float* right,
                          const double* chemical_potential, int32_t num_mu, double broadening,
                          double temperature, int32_t num_random, int32_t num_points, void* out);

/* The relabelling used for full-system runs (see pbk_config.locality_tile): order[new_row] = original row.
 * Host-only helper (no device needed); exposed so that callers / tests can inspect the layout. */
int pbk_locality_order(int64_t n, const int32_t* indptr, const int32_t* indices, int32_t tile, int32_t* order);
/* Two-level variant: macro-blocks of `macro_tiles` tiles first, clusters inside each block second (macro_tiles <= 1: the
 * one-level order above).  Tiles of a block are consecutive rows, so CTAs resident together share their halo rows in L2. */
int pbk_locality_order2(int64_t n, const int32_t* indptr, const int32_t* indices, int32_t tile, int32_t macro_tiles,
                        int32_t* order);

/* Host-only test hook: the scaled slot-major ELL (element (row, s) at s * pitch + row; padding = value 0, the row's own
 * index) exactly as it is uploaded -- OptimizedHamiltonian::create_scaled for `order` == NULL, create_reordered with the
 * relabelling order[new_row] = old_row otherwise (cppcore/src/kpm/OptimizedHamiltonian.cpp:55-152), csr_to_ell
 * (numeric/ellmatrix.hpp:65-82), scale factors from (min_energy, max_energy) like kpm::Scale (Bounds.hpp:19-25).
 * Call with val == NULL to learn k and pitch first; val holds k * pitch scalars, col k * pitch int32. */
int pbk_host_ell(int dtype, int64_t n, const int32_t* indptr, const int32_t* indices, const void* data, double min_energy,
                 double max_energy, const int32_t* order, int32_t* k, int64_t* pitch, void* val, int32_t* col);

/* Host-only helper: the light cone of `src` that LDOS runs on (engine.cu: light_cone / moments_ldos_cones) -- the sites within
 * `depth` bonds of `src` in the visiting order of the reference's relabelling (OptimizedHamiltonian.cpp:88-143: queue order,
 * row entries in ascending column order), queue[i] = site at position i, borders[j] = sites within distance j.
 * `queue` must hold n entries, `borders` depth + 1; *exhausted = 1 when the walk covered the whole connected component. */
int pbk_light_cone(int64_t n, const int32_t* indptr, const int32_t* indices, int32_t src, int32_t depth,
                   int32_t* queue, int64_t* queue_size, int32_t* borders, int32_t* num_borders, int32_t* exhausted);

/* Host-only check of the MT19937 jump-ahead used for segment-parallel starter generation: writes the 624-word
 * generator window positioned so that window[1..623] are the raw (untempered) words of draws
 * [position, position + 623) of a default-seeded std::mt19937. */
int pbk_mt_jump_window(uint64_t position, uint32_t* window);

/* ---- reporting ------------------------------------------------------------------------------ */
/* replaces: Core::get_stats / KPMStats (cppmodule/src/kpm.cpp:50-66) */
int pbk_get_stats(pbk_ctx* ctx, pbk_stats* out);
/* replaces: Core::report (Core.cpp:31-33); writes a NUL-terminated string of at most `size` bytes */
int pbk_report(pbk_ctx* ctx, int shortform, char* buffer, int64_t size);

/* ---- multi-GPU: one process per GPU, stochastic vectors / sites sharded over ranks ----------- */
/* No reference counterpart (the reference is single-process, thread-pool parallel over vector batches:
 * Compute.cpp:52-88).  The 128-byte id is an ncclUniqueId created on rank 0 and distributed by the
 * caller (bench.py uses torch.distributed for that plumbing). */
int pbk_comm_unique_id(char id[128]);
int pbk_comm_init(pbk_ctx* ctx, int32_t world_size, int32_t rank, const char id[128]);
int pbk_comm_destroy(pbk_ctx* ctx);
/* The partition every sharded entry point uses: rank `rank` of `world_size` owns units [first, first + count) of `total`
 * (contiguous blocks, remainder to the lowest ranks).  For stochastic vectors unit j consumes draws [j*N*w, (j+1)*N*w) of
 * the reference's single MT19937 stream, so the union over ranks is the single-GPU / CPU calculation.  Host-only. */
int pbk_shard(int32_t total, int32_t world_size, int32_t rank, int32_t* first, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* PBKPM_H */
