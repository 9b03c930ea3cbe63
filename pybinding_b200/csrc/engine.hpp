// engine.hpp -- host side of libpbkpm.so: the B200 counterpart of kpm::Core + DefaultCompute.
#pragma once
#include "../../include/pbkpm.h"
#include "kernels.cuh"

#include <complex>
#include <dlfcn.h>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <atomic>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace pbk {

using cd = std::complex<double>;

struct Error : std::runtime_error {
    int code;
    Error(int code, std::string const& msg) : std::runtime_error(msg), code(code) {}
};

void cuda_check(cudaError_t err, char const* what, char const* file, int line);
#define PBK_CUDA(call) ::pbk::cuda_check((call), #call, __FILE__, __LINE__)

/// RAII device allocation.  Blocks of a megabyte or more go back to a small per-process cache instead of cudaFree (which
/// unmaps the range under driver-wide locks: 0.1 - 0.5 s for the 3 GB of layouts that `set_hamiltonian` replaces when
/// several ranks share a host) and are handed out again to requests of exactly the same size -- the pattern of
/// re-setting a Hamiltonian of the same shape.  Release synchronises the device like cudaFree does; the cache is bounded,
/// flushed when an allocation fails and when the last engine goes away.
class DevBuf {
public:
    DevBuf() = default;
    explicit DevBuf(size_t bytes) { alloc(bytes); }
    DevBuf(DevBuf const&) = delete;
    DevBuf& operator=(DevBuf const&) = delete;
    DevBuf(DevBuf&& o) noexcept : ptr(o.ptr), size(o.size), dev(o.dev) { o.ptr = nullptr; o.size = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); ptr = o.ptr; size = o.size; dev = o.dev; o.ptr = nullptr; o.size = 0; } return *this; }
    ~DevBuf() { release(); }
    void alloc(size_t bytes);
    void ensure(size_t bytes) { if (bytes > size) { release(); alloc(bytes); } }
    void release();
    template<class T = void> T* as() const { return static_cast<T*>(ptr); }
    size_t bytes() const { return size; }
    static void flush_cache();   // cudaFree every cached block
    static size_t cached_bytes();
private:
    void* ptr = nullptr;
    size_t size = 0;
    int dev = -1;
};

/// Host array without value-initialisation: large buffers are first touched by the (parallel) code that fills them
template<class T> class RawVec {
public:
    RawVec() = default;
    RawVec(RawVec const&) = delete;
    RawVec& operator=(RawVec const&) = delete;
    RawVec(RawVec&& o) noexcept : p(o.p), n(o.n), cap(o.cap), pinned_(o.pinned_) { o.p = nullptr; o.n = o.cap = 0; o.pinned_ = false; }
    RawVec& operator=(RawVec&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; pinned_ = o.pinned_; o.p = nullptr; o.n = o.cap = 0; o.pinned_ = false; }
        return *this;
    }
    ~RawVec() { release(); }
    /// `pinned`: page-locked storage (uploads at the link rate; pageable memory when the driver refuses).  The storage is
    /// kept when it is large enough, so a context that sees several Hamiltonians pays for allocation and pinning once.
    void resize_uninit(size_t count, bool pinned = false);
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    bool is_pinned() const { return pinned_; }
    T& operator[](size_t i) { return p[i]; }
    T const& operator[](size_t i) const { return p[i]; }
private:
    void release();
    T* p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned_ = false;
};

/// Page-locked host staging buffer, kept by the context and grown on demand: uploads run at the link rate and the
/// pinning cost is paid once per context, not once per Hamiltonian
class PinnedBuf {
public:
    PinnedBuf() = default;
    PinnedBuf(PinnedBuf const&) = delete;
    PinnedBuf& operator=(PinnedBuf const&) = delete;
    ~PinnedBuf() { release(); }
    void* ensure(size_t bytes);
    void release();
private:
    void* ptr = nullptr;
    size_t size = 0;
};

/// kpm::Scale (cppcore/include/kpm/Bounds.hpp:11-33) including its single-precision constants
struct Scale {
    double a = 0, b = 0;
    Scale() = default;
    Scale(double min_energy, double max_energy);
};

/// kpm::SliceMap (cppcore/include/kpm/OptimizedHamiltonian.hpp:54-96)
struct SliceMap {
    std::vector<int32_t> data;
    int src_offset = 0, dest_offset = 0;
    int last_index() const { return static_cast<int>(data.size()) - 1; }
    int index(int n, int num_moments) const;
    int64_t optimal_size(int n, int num_moments) const { return data[index(n, num_moments)]; }
    bool uses_full_system(int num_moments) const { return static_cast<int>(data.size()) < num_moments / 2; }
};

struct Indices {
    std::vector<int32_t> src, dest;
    bool is_diagonal() const { return src == dest; }
    bool operator==(Indices const& o) const { return src == o.src && dest == o.dest; }
};

/// Device-resident scaled (and optionally BFS-reordered) Hamiltonian in slot-major ELL
enum OrderMode : int { ORDER_NATURAL = 0, ORDER_BFS = 1, ORDER_CLUSTER = 2 };

struct DeviceHamiltonian {
    bool valid = false;
    bool reordered = false;        // rows are relabelled: reorder_map / perm are set
    bool sliced = false;           // BFS order from the source: `map` holds the light-cone slices
    int64_t tile = 0;              // ORDER_CLUSTER: rows per locality cluster (the step kernel's CTA tile), else 0
    bool transient = false;        // rebuilt for every calculation (light-cone sub-systems): its launch sequence is not worth caching
    int64_t vec_rows = 0;          // length of a KPM vector on this layout (0: the system size); light-cone sub-systems are shorter
    std::vector<int32_t> order_queue;  // device row -> original (ORDER_CLUSTER only)
    Indices original_idx, idx;     // idx: positions in the device ordering
    SliceMap map;
    std::vector<int32_t> reorder_map;  // original -> device row (empty: identity)
    DevBuf val, col, perm;
    DevBuf queue_dev;              // row of the layout -> original site, on the device (layouts built by build.cu)
    bool host_order = true;        // reorder_map / order_queue hold the host copies of perm / queue_dev (else: downloaded on first use)
    // resident-tile variant of the step kernel (kernels_res.cu): chosen for layouts whose clusters are mostly surface
    bool res_enabled = false;      // the layout was ordered for it (tile = rows of a nominal resident tile)
    DevBuf res_tiles, res_halo, res_codes, res_vals;
    int res_ntiles = 0;
    uint32_t res_failed_row_bytes = 0;   // a row width for which the tiles did not fit (not retried)
    ResGeometry res_geo;
    double res_halo_frac = 0;      // halo rows / own rows over all tiles
    DevBuf packed;                 // row-major packed copy of the ELL arrays for the bulk-copy staged step kernel (ORDER_CLUSTER only)
    EllDev ell;
    double seconds = 0;
    uint64_t memory() const { return static_cast<uint64_t>(ell.rows) * ell.k * 0 + val.bytes() + col.bytes(); }
};

/// Result of the breadth-first relabelling from a source (host side of OptimizedHamiltonian::create_reordered)
struct BfsOrder {
    bool valid = false;
    Indices target, idx;               // requested (original) and relabelled indices
    std::vector<int32_t> queue;        // new -> original
    std::vector<int32_t> reorder_map;  // original -> new
    SliceMap map;
    bool valid_for(Indices const& t) const { return valid && target == t; }
};

/// Light cone of one source site: the breadth-first ball that the recursion can reach in `depth` steps.
/// queue[i] = original site at position i (shell by shell, the reference's BFS order: OptimizedHamiltonian.cpp:88-143),
/// borders[j] = number of sites within graph distance j.
struct Cone {
    std::vector<int32_t> queue;
    std::vector<int32_t> borders;
    bool exhausted = false;        // the ball is the whole connected component: every row has its full neighbourhood
    /// rows whose neighbours are all inside the ball
    int64_t complete_rows() const { return exhausted || borders.size() < 2 ? static_cast<int64_t>(queue.size()) : borders[borders.size() - 2]; }
    SliceMap map() const;
};

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: no link-time dependency, single-GPU use never touches it
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;   // optional
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    NcclApi() {
        for (char const* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) throw Error(PBK_NCCL_ERROR, std::string("cannot load libnccl.so.2: ") + dlerror());
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(handle, "ncclAllReduce"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(handle, "ncclBroadcast"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
        init_rank = dlsym(handle, "ncclCommInitRank");
        if (!GetUniqueId || !AllReduce || !CommDestroy || !init_rank) throw Error(PBK_NCCL_ERROR, "libnccl is missing required symbols");
    }
    void* init_rank = nullptr;
    void check(int r, char const* what) const {
        if (r != 0) throw Error(PBK_NCCL_ERROR, std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(r) : "?"));
    }
};
struct NcclId { char bytes[128]; };  // ncclUniqueId


/// Declared first in Engine, hence destroyed last: when the last engine of the process is gone (its buffers have just been
/// returned to the block cache) the cache is emptied.
struct DevCacheGuard {
    DevCacheGuard() { ++live(); }
    ~DevCacheGuard() { if (--live() == 0) DevBuf::flush_cache(); }
    DevCacheGuard(DevCacheGuard const&) = delete;
    DevCacheGuard& operator=(DevCacheGuard const&) = delete;
    static std::atomic<int>& live() { static std::atomic<int> n{0}; return n; }
};

class Engine {
    DevCacheGuard cache_guard;
public:
    Engine(int device, pbk_config const& config);
    ~Engine();

    std::mutex mutex;            // one in-flight calculation per context
    std::string last_error;

    void set_progress(pbk_progress_fn fn, void* user) { progress_fn = fn; progress_user = user; }
    void set_hamiltonian(int dtype, int64_t n, const int32_t* indptr, const int32_t* indices, const void* data);

    void bounds(double* mn, double* mx, int32_t* loops);
    Scale scaling_factors();
    int required_num_moments(double broadening);

    // compute-strategy level
    void moments_dos(int M, int num_random, cd* out);
    void moments_ldos(int M, const int32_t* idx, int nidx, cd* out);
    void moments_greens(int M, int row, const int32_t* cols, int ncols, cd* out);
    void moments_kubo(int M, const float* left, const float* right, int num_random, cd* out);
    void moments_diagonal(int M, const cd* r0, int count, cd* out);
    void random_vectors(int count, cd* out);

    // kpm::Core level
    void core_moments(int num_moments, const cd* alpha, const cd* beta, int64_t op_rows, const int32_t* op_indptr,
                      const int32_t* op_indices, const cd* op_data, cd* out);
    void calc_dos(const double* energy, int ne, double broadening, int num_random, double* out);
    void calc_ldos(const double* energy, int ne, double broadening, const int32_t* idx, int nidx, double* out);
    void calc_greens(int row, const int32_t* cols, int ncols, const double* energy, int ne, double broadening, cd* out);
    void calc_conductivity(const float* left, const float* right, const double* mu, int nmu, double broadening,
                           double temperature, int num_random, int num_points, cd* out);

    pbk_stats get_stats() const { return stats; }
    std::string report(bool shortform) const;

    void comm_init(int world, int rank, const char* id);
    void comm_destroy();

private:
    // ---- configuration / device ----
    int device = 0;
    int num_sms = 148;
    int step_tpb = 256, step_blocks_per_sm = 0, step_prefetch = 0, step_prefetch_mask = 0;
    int res_mode = 1;            // PBK_RES: resident-tile step kernel -- 0 off, 1 where the locality clusters are mostly surface, 2 always
    int64_t res_tile = 384;      // PBK_RES_TILE: rows of a nominal resident tile (the locality clusters of such layouts)
    int res_buffers = 1;         // PBK_RES_BUFS: resident-tile buffers per CTA (2: next tile loads during this one; slower: fewer warps)
    int res_row_bytes = 64;      // PBK_RES_ROW: bytes per row of a pass of the resident kernel (16 float lanes)
    int res_ctas = 3, res_stages = 2;   // PBK_RES_CTAS, PBK_RES_STAGES
    int64_t layout_tile = 0;     // cluster size of the current Hamiltonian's full-system layout
    bool layout_res = false;     // ... and whether it was ordered for the resident kernel
    int bulk_stages = 4;         // pipeline depth of the bulk-copy staged step kernel (0: general kernel only)
    bool bulk_xstage = true;     // staged kernel: the CTA's own x rows go through shared memory too
    int bulk_release = 0;        // PBK_RELEASE (experiment): when a warp hands a stage back to the producer
    int64_t locality_tile = 0;   // rows per locality cluster of the full-system layout (0: keep the caller's order)
    pbk_config config{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev_begin = nullptr, ev_end = nullptr;
    pbk_progress_fn progress_fn = nullptr;
    void* progress_user = nullptr;

    // ---- host copy of the Hamiltonian (original order, unscaled) ----
    int dtype = -1;
    int64_t n = 0;
    RawVec<int32_t> h_indptr, h_indices;   // page-locked mirror of the caller's arrays (host-side graph walks: BFS, light cones)
    RawVec<char> h_data;
    bool has_h = false;
    DevBuf d_indptr, d_indices, d_data;    // the same CSR on the device: every layout is built from it by a kernel (build.cu)
    bool dev_csr = false;
    int dev_build = 1;                     // PBK_DEVBUILD=0: build the layouts on the host (the round-1 path; A/B and fallback)
    int bcast_order = 1;                   // PBK_BCAST_ORDER=0: every rank computes the locality ordering itself
    // locality ordering of the full-system layout, computed while set_hamiltonian copies the arrays (rank 0 only when a
    // communicator is attached: the other ranks receive it over NVLink)
    std::vector<int32_t> cluster_queue, cluster_rmap;
    int64_t cluster_tile = 0;
    DevBuf cluster_queue_dev;              // the ordering on the device (uploaded or received by broadcast)
    bool cluster_on_device = false, cluster_on_host = false;
    PinnedBuf stage_val, stage_col;   // ELL staging of the full-system layout (host build path)
    DevBuf width_dev;                 // one int: widest row of the layout being built

    // ---- bounds ----
    bool have_bounds = false;
    double bounds_min = 0, bounds_max = 0;
    int lanczos_loops = 0;
    double bounds_seconds = 0;

    // ---- device Hamiltonians ----
    DeviceHamiltonian natural;    // scaled, original order (DOS / conductivity / moments)
    DeviceHamiltonian optimized;  // scaled + BFS-reordered for the last (src, dest) request (LDOS / Green's)
    DeviceHamiltonian unscaled;   // original values, original order (Lanczos)

    // ---- work buffers ----
    DevBuf cone_val, cone_col, cone_queue, cone_gmap, cone_table;   // light-cone sub-system of the site being processed
    int64_t cone_gmap_rows = 0;  // cone_gmap holds -1 for this many rows
    int64_t coarse_sites = 16;   // PBK_COARSE: consecutive sites per super-node of the macro-block pass (1: no coarsening)
    int64_t macro_tiles = 256;   // PBK_MACRO: tiles per macro-block of the two-level locality ordering (0: one level)
    bool identity_order = false; // PBK_IDENTITY_ORDER=1: tiles of consecutive rows in the caller's order instead of locality clusters
    int cone_mode = 1;
    int cone_group_cap = 0;              // PBK_CONE_GROUP: at most this many light-cone sub-systems per launch (0: bounded by memory only)           // PBK_CONE=0: the previous host-side full BFS relabelling for LDOS
    DevBuf vec_a, vec_b, vec_t, raw, mom, m01, acc, partials, counter, scratch, mt_state, mt_states, idx_buf;
    static constexpr int MT_MAX_SEGMENTS = 2048;
    uint64_t stream_pos = 0;     // next draw of the reference's random stream
    bool mt_sequential = false;  // PBK_MT_SEQUENTIAL=1: single-CTA generator (cross-check of the jump-ahead path)

    // ---- multi-GPU ----
    std::unique_ptr<NcclApi> nccl;
    void* comm = nullptr;
    int world = 1, rank = 0;

    // ---- CUDA graphs of launch-bound recursions (small systems): the whole sequence of step launches of one
    //      diagonal run is captured once and replayed, keyed by every baked-in pointer and row count ----
    struct RecursionGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0, step_launches = 0, bulk_launches = 0; double step_bytes = 0; };
    std::map<std::vector<int64_t>, RecursionGraph> graph_cache;
    int persist_mode = 1;            // PBK_PERSIST=0: no persistent single-launch recursion for small systems
    DevBuf persist_table, persist_barrier;
    DevBuf kubo_l, kubo_r, kubo_ws;   // Kubo-Bastin stacks and split-K workspace, kept between conductivity calls
    bool keep_kubo_buffers = false;
    int graph_mode = 1;
    double graph_max_bytes = 64e6;   // vector block size up to which a recursion counts as launch-bound
    void clear_graphs();

    pbk_stats stats{};
    double last_total_seconds = 0;

    // ---- helpers ----
    void require_hamiltonian() const;
    void compute_bounds();
    BfsOrder bfs_order(Indices const& target) const;
    BfsOrder bfs_ready;           // relabelling computed ahead of build_device_hamiltonian (moments_ldos)
    void build_device_hamiltonian(DeviceHamiltonian& dh, bool scaled, int order, Indices const& target);
    /// the layout's rows written by build.cu from the resident CSR; false: not applicable (rows too long), use the host path
    bool build_layout_on_device(DeviceHamiltonian& dh, int mode, Scale s, const float* positions_dev);
    bool ensure_res_meta(DeviceHamiltonian& dh, int R);   // tiles, halo lists and local codes of the resident-tile kernel
    void ensure_host_order(DeviceHamiltonian& dh);   // host copies of a layout's order maps (downloaded on first use)
    void broadcast(void* dev, int64_t count_int32, int root);
    DeviceHamiltonian& natural_hamiltonian();
    DeviceHamiltonian& optimized_for(Indices const& target);
    DeviceHamiltonian& unscaled_hamiltonian();
    Cone bfs_cone(int32_t src, int depth, std::vector<int32_t>& mark) const;
    /// LDOS moments on per-site light-cone sub-systems cut out of the resident Hamiltonian; false: the full-system batch is cheaper
    bool moments_ldos_cones(int M, Indices const& target, cd* out);
    void upload_operator(DeviceHamiltonian& dh, const float* positions, DeviceHamiltonian& like);  // velocity operator
    void upload_csr_operator(DeviceHamiltonian& dh, int64_t rows, const int32_t* indptr, const int32_t* indices, const cd* data,
                             DeviceHamiltonian& like);

    int pick_batch(int vectors, int extra_blocks) const;
    int lane_pad(int R) const;
    void ensure_moment_buffers(int R, int M);
    void step(DeviceHamiltonian const& h, const void* x, void* y, void* y2, int64_t nrows, int R, bool subtract, bool sums,
              double scale, int M, int nstep, int fin, int64_t y_block_stride = 0, int64_t y2_block_stride = 0);
    /// diagonal recursion for the R vectors in vec_a (r0); moments land in `mom` ([R][M] c128)
    void run_diagonal(DeviceHamiltonian const& h, int R, int M, bool opt_size);
    /// off-diagonal recursion for the single vector in vec_a; `collect(n, r, half)` is called for every moment
    void run_offdiagonal(DeviceHamiltonian const& h, int M, bool opt_size, std::function<void(int, void*, double)> const& collect);
    void reset_stats(int M, DeviceHamiltonian const& h, bool opt_size, double multiplier);
    void begin_moments();
    void end_moments();
    void progress(int64_t delta, int64_t total) { if (progress_fn) progress_fn(delta, total, progress_user); }
    void generate_random_block(DeviceHamiltonian const& h, int lanes, int R, void* dst);
    void seed_stream(int64_t skip_vectors);

    void spectral_density_device(const cd* moments, int M, int cols, int64_t col_stride, int64_t n_stride, const double* energy, int ne,
                                 Scale s, double* out);
    void greens_device(const cd* moments, int M, int cols, const double* energy, int ne, Scale s, cd* out);

    void shard(int total, int* first, int* count) const;
    void allreduce(double* dev, int64_t count);

    double moments_wall0 = 0;
    int64_t launches = 0;
};

/// units [first, first + count) of `total` owned by `rank` (engine.cu)
void shard_range(int total, int world, int rank, int* first, int* count);

/// host-only test hook: the scaled ELL that would be uploaded (engine.cu)
int host_scaled_ell(int dtype, int64_t n, const int32_t* indptr, const int32_t* indices, const void* data, double min_energy,
                    double max_energy, const int32_t* order, int32_t* k_out, int64_t* pitch_out, void* val, int32_t* col);

/// truncated breadth-first walk from `src` (engine.cu); `mark` holds -1 for every site and is restored on return
Cone light_cone(const int32_t* indptr, const int32_t* indices, int32_t src, int depth, std::vector<int32_t>& mark);

/// locality relabelling of the full-system layout (engine.cu)
void cluster_order(int64_t n, const int32_t* indptr, const int32_t* indices, int64_t tile,
                   std::vector<int32_t>& queue, std::vector<int32_t>& rmap, int64_t macro_tiles = 0, int64_t coarse = 16);

// kernels (src/kpm/Kernel.cpp:6-49)
int round_num_moments(int n);
std::vector<double> damping_coefficients(int kernel, double lambda_value, int n);
int kernel_required_num_moments(int kernel, double lambda_value, double scaled_broadening);

// reconstruction (include/kpm/reconstruct.hpp:16-143), double precision with the reference's float constants
void reconstruct_kubo_bastin(const double* sum_nm_c128, const std::vector<double>& scaled_samples, const double* mu, int nmu,
                             double temperature, Scale s, cd* out);
cudaError_t launch_kubo_gamma_sum(const double* mu_c128, int M, const double* scaled_samples, int np, double* out_c128, cudaStream_t s);

} // namespace pbk
