// engine.cu -- host orchestration of the GPU KPM engine (the counterpart of kpm::Core,
// OptimizedHamiltonian, Starter and DefaultCompute: cppcore/src/kpm/*.cpp of the reference).
//
// Design (B200-first, not a port):
//  * the scaled Hamiltonian lives on the device in slot-major ELL.  Stochastic quantities (DOS, conductivity, moments)
//    use a two-level *locality ordering* of the sites (breadth-first clusters of 256 rows inside macro-blocks of 256
//    clusters): results are permutation invariant, a CTA owns one cluster at a time and the clusters resident together
//    share their halo rows through L2.  Unit-vector quantities use the reference's breadth-first relabelling from the
//    source so that the light cone of the recursion is a row prefix (`SliceMap`) and each step only touches
//    `optimal_size(n)` rows; LDOS does this on a *sub-system* cut out of the resident matrix by a kernel (the host only
//    walks the ball the recursion can reach), Green's relabels on the host;
//  * all R vectors of a batch are advanced by ONE fused kernel launch per Chebyshev step; the moments are produced on
//    the device by the kernel's last block, so a whole recursion is an uninterrupted stream of launches (replayed as a
//    CUDA graph when the system is small enough to be launch-bound) with a single device->host copy at the end;
//  * random starters are the reference's own MT19937 stream, generated on the device (GF(2) jump-ahead);
//  * multi-GPU: vectors / LDOS sites are sharded over ranks, one ncclAllReduce of the moment sums at the end.
#include "engine.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <exception>
#include <thread>

namespace pbk {

using cf = std::complex<float>;

void cuda_check(cudaError_t err, char const* what, char const* file, int line) {
    if (err == cudaSuccess) return;
    char buf[512];
    std::snprintf(buf, sizeof(buf), "CUDA error '%s' (%s) at %s:%d: %s", cudaGetErrorName(err), cudaGetErrorString(err), file, line, what);
    throw Error(PBK_CUDA_ERROR, buf);
}

namespace {
struct DevBlockCache {
    struct Block { int dev; void* ptr; size_t size; };
    std::mutex mutex;
    std::vector<Block> blocks;
    size_t bytes = 0;
    static constexpr size_t MIN_BLOCK = size_t{1} << 20, MAX_BYTES = size_t{12} << 30;
};
DevBlockCache& dev_cache() { static DevBlockCache c; return c; }
} // anonymous namespace

size_t DevBuf::cached_bytes() {
    auto& c = dev_cache();
    std::lock_guard<std::mutex> lock(c.mutex);
    return c.bytes;
}

void DevBuf::flush_cache() {
    auto& c = dev_cache();
    std::vector<DevBlockCache::Block> blocks;
    { std::lock_guard<std::mutex> lock(c.mutex); blocks.swap(c.blocks); c.bytes = 0; }
    for (auto const& b : blocks) cudaFree(b.ptr);
}

void DevBuf::alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaGetDevice(&dev);
    if (bytes >= DevBlockCache::MIN_BLOCK) {
        auto& c = dev_cache();
        std::lock_guard<std::mutex> lock(c.mutex);
        for (size_t i = 0; i < c.blocks.size(); ++i) {
            if (c.blocks[i].dev == dev && c.blocks[i].size == bytes) {
                ptr = c.blocks[i].ptr; size = bytes;
                c.bytes -= bytes;
                c.blocks.erase(c.blocks.begin() + static_cast<std::ptrdiff_t>(i));
                return;
            }
        }
    }
    cudaError_t err = cudaMalloc(&ptr, bytes);
    if (err == cudaErrorMemoryAllocation) {   // give the cached blocks back and try once more
        cudaGetLastError();
        flush_cache();
        err = cudaMalloc(&ptr, bytes);
    }
    if (err != cudaSuccess) ptr = nullptr;
    PBK_CUDA(err);
    size = bytes;
}
void DevBuf::release() {
    if (!ptr) return;
    auto& c = dev_cache();
    bool cached = false;
    if (size >= DevBlockCache::MIN_BLOCK && size <= DevBlockCache::MAX_BYTES / 2) {
        int cur = dev;
        cudaGetDevice(&cur);
        if (cur != dev) cudaSetDevice(dev);
        cudaDeviceSynchronize();   // what cudaFree guarantees: nothing in flight still uses the block
        if (cur != dev) cudaSetDevice(cur);
        std::lock_guard<std::mutex> lock(c.mutex);
        while (!c.blocks.empty() && c.bytes + size > DevBlockCache::MAX_BYTES) {   // bounded: oldest blocks go first
            cudaFree(c.blocks.front().ptr);
            c.bytes -= c.blocks.front().size;
            c.blocks.erase(c.blocks.begin());
        }
        c.blocks.push_back({dev, ptr, size});
        c.bytes += size;
        cached = true;
    }
    if (!cached) cudaFree(ptr);
    ptr = nullptr; size = 0;
}

void* PinnedBuf::ensure(size_t bytes) {
    if (bytes > size) {
        release();
        if (cudaMallocHost(&ptr, bytes) != cudaSuccess) { cudaGetLastError(); ptr = nullptr; size = 0; return nullptr; }  // caller falls back to pageable memory
        size = bytes;
    }
    return ptr;
}
void PinnedBuf::release() {
    if (ptr) { cudaFreeHost(ptr); ptr = nullptr; size = 0; }
}

template<class T> void RawVec<T>::release() {
    if (p) { if (pinned_) cudaFreeHost(p); else delete[] p; }
    p = nullptr; n = cap = 0; pinned_ = false;
}
template<class T> void RawVec<T>::resize_uninit(size_t count, bool pinned) {
    if (count <= cap && (pinned_ || !pinned)) { n = count; return; }
    release();
    if (count == 0) return;
    if (pinned) {
        void* q = nullptr;
        if (cudaMallocHost(&q, count * sizeof(T)) == cudaSuccess) { p = static_cast<T*>(q); pinned_ = true; }
        else cudaGetLastError();
    }
    if (!p) p = new T[count];
    n = cap = count;
}
template class RawVec<int32_t>;
template class RawVec<char>;

static double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/// Runs `fn(t)` on `nt` threads and joins them all before returning, also when one of them (or the thread creation
/// itself) throws: the first exception is re-thrown on the caller's thread after the join, so nothing ever unwinds
/// past a joinable std::thread (which would std::terminate the process across the C ABI).
template<class F> static void run_pool(int nt, F fn) {
    if (nt <= 1) { fn(0); return; }
    std::vector<std::thread> threads;
    std::exception_ptr error;
    std::mutex error_mutex;
    auto guarded = [&](int t) {
        try { fn(t); }
        catch (...) { std::lock_guard<std::mutex> lk(error_mutex); if (!error) error = std::current_exception(); }
    };
    try {
        threads.reserve(static_cast<size_t>(nt));
        for (int t = 0; t < nt; ++t) threads.emplace_back(guarded, t);
    } catch (...) {
        std::lock_guard<std::mutex> lk(error_mutex);
        if (!error) error = std::current_exception();
    }
    for (auto& t : threads) t.join();
    if (error) std::rethrow_exception(error);
}

static int host_threads() {
    int nt = static_cast<int>(std::thread::hardware_concurrency());
    return std::max(1, std::min(nt, 16));
}

template<class F> static void parallel_rows(int64_t n, F fn, int64_t min_items = int64_t{1} << 16, int max_threads = 0) {
    int nt = host_threads();
    if (max_threads > 0) nt = std::min(nt, max_threads);
    if (n < min_items) nt = 1;
    if (nt == 1) { fn(int64_t{0}, n); return; }
    int64_t const chunk = (n + nt - 1) / nt;
    run_pool(nt, [&](int t) {
        int64_t const b = t * chunk, e = std::min<int64_t>(n, b + chunk);
        if (b < e) fn(b, e);
    });
}

/// A std::thread that is always joined when the scope ends; an exception thrown by its body is kept and re-thrown by
/// join_and_rethrow() on the owner's thread.
class ScopedThread {
public:
    ScopedThread() = default;
    template<class F> explicit ScopedThread(F fn) {
        thread = std::thread([this, fn]() mutable { try { fn(); } catch (...) { error = std::current_exception(); } });
    }
    ScopedThread(ScopedThread const&) = delete;
    ScopedThread& operator=(ScopedThread const&) = delete;
    ~ScopedThread() { if (thread.joinable()) thread.join(); }
    void join_and_rethrow() {
        if (thread.joinable()) thread.join();
        if (error) { auto e = error; error = nullptr; std::rethrow_exception(e); }
    }
private:
    std::thread thread;
    std::exception_ptr error;
};

// ------------------------------------------------------------------------------------------------
// Scale, SliceMap, kernels
// ------------------------------------------------------------------------------------------------
Scale::Scale(double min_energy, double max_energy) {  // Bounds.hpp:19-25, float literals on purpose
    constexpr auto tolerance = 0.01f;
    a = 0.5f * (max_energy - min_energy) * (1 + tolerance);
    b = 0.5f * (max_energy + min_energy);
    if (std::abs(b / a) < 0.01f * tolerance) { b = 0; }
}

int SliceMap::index(int n, int num_moments) const {  // OptimizedHamiltonian.hpp:67-77
    int const mid = (num_moments - 1 + dest_offset - src_offset) / 2;
    int const max = std::min(last_index(), mid + src_offset);
    if (n < mid) return std::min(max, n + src_offset);
    return std::min(max, num_moments - 1 - n + dest_offset);
}

int round_num_moments(int n) {
    if (n < 2) return 2;
    while ((n - 2) % 4 != 0) ++n;
    return n;
}

static constexpr float pi_f = 3.14159265358979323846f;  // numeric/constant.hpp:8
static constexpr float kb_f = 8.6173303e-5f;             // numeric/constant.hpp:20

std::vector<double> damping_coefficients(int kernel, double lambda_value, int n) {
    std::vector<double> g(n);
    auto const N = static_cast<double>(n);
    for (int i = 0; i < n; ++i) {
        auto const k = static_cast<double>(i);
        if (kernel == PBK_JACKSON) {
            auto const Np = N + 1;
            constexpr auto pi = double{pi_f};
            g[i] = ((Np - k) * std::cos(pi * k / Np) + std::sin(pi * k / Np) / std::tan(pi / Np)) / Np;
        } else if (kernel == PBK_LORENTZ) {
            g[i] = std::sinh(lambda_value * (1 - k / N)) / std::sinh(lambda_value);
        } else {
            g[i] = 1.0;
        }
    }
    return g;
}

int kernel_required_num_moments(int kernel, double lambda_value, double scaled_broadening) {
    double const num = (kernel == PBK_LORENTZ) ? lambda_value : static_cast<double>(pi_f);
    return round_num_moments(static_cast<int>(num / scaled_broadening) + 1);
}


// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
Engine::Engine(int device_, pbk_config const& cfg) : device(device_), config(cfg) {
    if (config.min_energy > config.max_energy) {
        throw Error(PBK_INVALID_ARGUMENT, "KPM: Invalid energy range specified (min > max).");  // Core.cpp:20-22
    }
    if (config.kernel == PBK_LORENTZ && config.lambda_value <= 0) {
        throw Error(PBK_INVALID_ARGUMENT, "Lorentz kernel: lambda must be positive.");  // Kernel.cpp:25
    }
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        throw Error(PBK_CUDA_ERROR, std::string("pbkpm needs a CUDA device and has no CPU fallback: ")
                                    + (err != cudaSuccess ? cudaGetErrorString(err) : "no device found"));
    }
    if (device < 0 || device >= count) throw Error(PBK_INVALID_ARGUMENT, "invalid CUDA device index");
    PBK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    PBK_CUDA(cudaGetDeviceProperties(&prop, device));
    num_sms = prop.multiProcessorCount;
    // layout / launch tuning: config first, environment overrides for experiments
    auto env_int = [](char const* name, long fallback) { char const* v = std::getenv(name); return v ? std::strtol(v, nullptr, 10) : fallback; };
    locality_tile = env_int("PBK_TILE", config.locality_tile);
    if (locality_tile == 0) locality_tile = 256;
    if (locality_tile < 0) locality_tile = 0;
    mt_sequential = env_int("PBK_MT_SEQUENTIAL", 0) != 0;
    step_tpb = static_cast<int>(env_int("PBK_TPB", 256));
    step_blocks_per_sm = static_cast<int>(env_int("PBK_BPSM", 0));
    step_prefetch = static_cast<int>(env_int("PBK_PF", 4));
    step_prefetch_mask = static_cast<int>(env_int("PBK_PFMASK", 0));
    bulk_stages = static_cast<int>(env_int("PBK_BULK", 4));
    bulk_xstage = env_int("PBK_XS", 1) != 0;
    bulk_release = static_cast<int>(env_int("PBK_RELEASE", 0));
    identity_order = env_int("PBK_IDENTITY_ORDER", 0) != 0;
    coarse_sites = env_int("PBK_COARSE", 16);
    res_mode = static_cast<int>(env_int("PBK_RES", 1));
    res_tile = env_int("PBK_RES_TILE", 384);
    res_buffers = static_cast<int>(env_int("PBK_RES_BUFS", 1));
    res_row_bytes = static_cast<int>(env_int("PBK_RES_ROW", 64));
    res_ctas = static_cast<int>(env_int("PBK_RES_CTAS", 3));
    res_stages = static_cast<int>(env_int("PBK_RES_STAGES", 3));
    if (res_tile < 64 || res_tile % 64 != 0) res_tile = 384;
    if (res_row_bytes < 16 || res_row_bytes % 16 != 0) res_row_bytes = 64;
    dev_build = static_cast<int>(env_int("PBK_DEVBUILD", 1));
    bcast_order = static_cast<int>(env_int("PBK_BCAST_ORDER", 1));
    macro_tiles = env_int("PBK_MACRO", 256);   // 65 k-site macro-blocks: +4 % on configs[1] in short runs, +1.5 % in the power-capped bench (profiles/r01_ab_order_v6.log)
    cone_mode = static_cast<int>(env_int("PBK_CONE", 1));
    cone_group_cap = static_cast<int>(env_int("PBK_CONE_GROUP", 0));
    graph_mode = static_cast<int>(env_int("PBK_GRAPH", 1));
    persist_mode = static_cast<int>(env_int("PBK_PERSIST", 1));
    graph_max_bytes = 1e6 * static_cast<double>(env_int("PBK_GRAPH_MAX_MB", 64));
    PBK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&ev0, &ev1, &ev2, &ev3, &ev_begin, &ev_end}) PBK_CUDA(cudaEventCreate(e));
    counter.alloc(64);
    width_dev.alloc(64);
    PBK_CUDA(cudaMemsetAsync(counter.as(), 0, 64, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
    mt_state.alloc(sizeof(uint32_t) * (MT_N + 8));
}

void Engine::clear_graphs() {
    for (auto& kv : graph_cache) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graph_cache.clear();
}

Engine::~Engine() {
    cudaSetDevice(device);
    clear_graphs();
    comm_destroy();
    for (cudaEvent_t e : {ev0, ev1, ev2, ev3, ev_begin, ev_end}) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

void Engine::require_hamiltonian() const {
    if (!has_h) throw Error(PBK_LOGIC_ERROR, "pbkpm: no Hamiltonian has been set");
}

/// Share of outside rows referenced by a breadth-first ball of `tile` sites grown in the middle of the system: the
/// halo of a locality cluster relative to its size (0.2 - 0.3 for a 2-D lattice at 256 sites, ~0.9 for a cubic one)
static double probe_halo_fraction(int64_t n, const int32_t* indptr, const int32_t* indices, int64_t tile) {
    if (n < 4 * tile) return 0.0;
    std::vector<int32_t> queue;
    std::map<int32_t, int> seen;   // site -> 0: inside the ball, 1: halo
    int32_t const seed = static_cast<int32_t>(n / 2);
    queue.push_back(seed); seen[seed] = 0;
    for (size_t head = 0; head < queue.size() && static_cast<int64_t>(queue.size()) < tile; ++head) {
        int32_t const row = queue[head];
        for (int p = indptr[row]; p < indptr[row + 1] && static_cast<int64_t>(queue.size()) < tile; ++p) {
            int32_t const c = indices[p];
            if (seen.find(c) == seen.end()) { seen[c] = 0; queue.push_back(c); }
        }
    }
    int64_t halo = 0;
    for (int32_t row : queue) for (int p = indptr[row]; p < indptr[row + 1]; ++p) if (seen.find(indices[p]) == seen.end()) { seen[indices[p]] = 1; ++halo; }
    return static_cast<double>(halo) / static_cast<double>(queue.size());
}

void Engine::set_hamiltonian(int dt, int64_t n_, const int32_t* indptr, const int32_t* indices, const void* data) {
    if (dt < 0 || dt > 3) throw Error(PBK_INVALID_ARGUMENT, "invalid dtype");
    if (n_ <= 0 || !indptr || !indices || !data) throw Error(PBK_INVALID_ARGUMENT, "invalid Hamiltonian arrays");
    PBK_CUDA(cudaSetDevice(device));
    bool const timing = std::getenv("PBK_TIMING") != nullptr;
    // the order maps of the previous full-system layout are the scratch arrays of the next ordering (no allocation,
    // no first-touch page faults when a context sees one Hamiltonian after another)
    if (natural.host_order && natural.order_queue.size() >= cluster_queue.size()) {
        cluster_queue.swap(natural.order_queue);
        cluster_rmap.swap(natural.reorder_map);
    }
    has_h = false;
    kubo_l.release(); kubo_r.release(); kubo_ws.release();
    clear_graphs();
    natural = DeviceHamiltonian();
    bfs_ready = BfsOrder();
    optimized = DeviceHamiltonian();
    unscaled = DeviceHamiltonian();
    dtype = dt;
    n = n_;
    int64_t const nnz = indptr[n];
    double const t_set0 = now_seconds();
    // The locality ordering of the full-system layout only needs the sparsity pattern: it runs on the caller's arrays in
    // a second thread while this one mirrors them into page-locked memory and uploads them.  With a communicator
    // attached the Hamiltonian is the same on every rank (the sharding contract), so rank 0 alone computes the
    // ordering and the others receive it by one broadcast over NVLink: set_hamiltonian is then a collective call.
    cluster_tile = 0; cluster_on_device = false; cluster_on_host = false;
    // cluster size of the layout: 256-row clusters for the staged kernel; where such a cluster is mostly surface (3-D
    // lattices) larger ones whose x rows the resident-tile kernel keeps in shared memory
    layout_res = locality_tile > 0 && !identity_order && bulk_stages >= 2 && dev_build &&
                 (res_mode >= 2 || (res_mode == 1 && n_ >= 64 * res_tile && probe_halo_fraction(n_, indptr, indices, locality_tile) > 0.5));
    layout_tile = layout_res ? res_tile : locality_tile;
    int64_t const layout_macro = layout_tile > 0 ? std::max<int64_t>(1, macro_tiles * locality_tile / layout_tile) : macro_tiles;
    bool const want_order = locality_tile > 0 && !identity_order;
    bool const receive_order = want_order && world > 1 && comm && bcast_order && nccl && nccl->Broadcast;
    std::unique_ptr<ScopedThread> ordering;   // joined on every exit path, exceptions of its body re-thrown after the join
    if (want_order && !(receive_order && rank != 0)) {
        cluster_tile = layout_tile;
        ordering = std::make_unique<ScopedThread>([this, n_, indptr, indices, layout_macro] { cluster_order(n_, indptr, indices, cluster_tile, cluster_queue, cluster_rmap, layout_macro, coarse_sites); });
    }
    size_t const sz = dtype_size(dt);
    h_indptr.resize_uninit(static_cast<size_t>(n) + 1, true);
    h_indices.resize_uninit(static_cast<size_t>(nnz), true);
    h_data.resize_uninit(static_cast<size_t>(nnz) * sz, true);
    d_indptr.ensure(sizeof(int32_t) * (static_cast<size_t>(n) + 1));
    d_indices.ensure(sizeof(int32_t) * static_cast<size_t>(std::max<int64_t>(nnz, 1)));
    d_data.ensure(sz * static_cast<size_t>(std::max<int64_t>(nnz, 1)));
    // several ranks share one host: a bandwidth-bound copy gains nothing from 16 threads per rank, and rank 0's ordering
    // threads should not be time-sliced against 8 x 16 of them
    int const mt = world > 1 ? 4 : 0;
    parallel_rows(n + 1, [&](int64_t b, int64_t e) { std::memcpy(h_indptr.data() + b, indptr + b, sizeof(int32_t) * static_cast<size_t>(e - b)); }, int64_t{1} << 16, mt);
    PBK_CUDA(cudaMemcpyAsync(d_indptr.as(), h_indptr.data(), sizeof(int32_t) * (static_cast<size_t>(n) + 1), cudaMemcpyHostToDevice, stream));
    parallel_rows(nnz, [&](int64_t b, int64_t e) { std::memcpy(h_indices.data() + b, indices + b, sizeof(int32_t) * static_cast<size_t>(e - b)); }, int64_t{1} << 16, mt);
    PBK_CUDA(cudaMemcpyAsync(d_indices.as(), h_indices.data(), sizeof(int32_t) * static_cast<size_t>(nnz), cudaMemcpyHostToDevice, stream));
    parallel_rows(nnz, [&](int64_t b, int64_t e) { std::memcpy(h_data.data() + b * sz, static_cast<const char*>(data) + b * sz, sz * static_cast<size_t>(e - b)); }, int64_t{1} << 16, mt);
    PBK_CUDA(cudaMemcpyAsync(d_data.as(), h_data.data(), sz * static_cast<size_t>(nnz), cudaMemcpyHostToDevice, stream));
    dev_csr = true;
    double const t_copy = now_seconds();
    if (ordering) { ordering->join_and_rethrow(); cluster_on_host = true; }
    double const t_order = now_seconds();
    if (want_order && dev_build) {   // the ordering on the device: uploaded by its owner, broadcast to the other ranks
        cluster_queue_dev.ensure(sizeof(int32_t) * static_cast<size_t>(n));
        if (cluster_on_host) PBK_CUDA(cudaMemcpyAsync(cluster_queue_dev.as(), cluster_queue.data(), sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
        if (receive_order) { broadcast(cluster_queue_dev.as(), n, 0); cluster_tile = layout_tile; }
        cluster_on_device = true;
    }
    PBK_CUDA(cudaStreamSynchronize(stream));
    stats = pbk_stats{};
    stats.h2d_bytes = static_cast<int64_t>(sizeof(int32_t) * (static_cast<size_t>(n) + 1 + nnz) + sz * nnz + (cluster_on_device && cluster_on_host ? sizeof(int32_t) * n : 0));
    if (timing) std::fprintf(stderr, "[pbkpm] set_hamiltonian: mirror + upload issued %.3f s, + wait for the ordering %.3f s, + order upload / broadcast + sync %.3f s\n",
                             t_copy - t_set0, t_order - t_copy, now_seconds() - t_order);
    has_h = true;
    cone_gmap_rows = 0;
    have_bounds = false;
    lanczos_loops = 0;
    bounds_seconds = 0;
    if (config.min_energy != config.max_energy) {
        bounds_min = config.min_energy; bounds_max = config.max_energy;
        have_bounds = true;
    }
}

// ------------------------------------------------------------------------------------------------
// Host-side construction of the device Hamiltonian: scale (+ BFS relabel) + ELL, then one upload.
// Semantics follow OptimizedHamiltonian::create_scaled / create_reordered (src/kpm/OptimizedHamiltonian.cpp:55-152)
// and csr_to_ell (numeric/ellmatrix.hpp:65-82); padding uses value 0 and the row's own index.
// ------------------------------------------------------------------------------------------------
namespace {

template<class T> struct real_of { using type = T; };
template<class R> struct real_of<std::complex<R>> { using type = R; };

struct HostEll {
    int k = 0;
    int64_t pitch = 0;
    char* val = nullptr;        // k * pitch scalars
    int32_t* col = nullptr;     // k * pitch
    size_t val_bytes = 0, col_count = 0;
    RawVec<char> val_own;       // storage when no staging buffer was supplied
    RawVec<int32_t> col_own;
};

template<class T>
HostEll build_ell_host(int64_t n, const int32_t* indptr, const int32_t* indices, const T* data, bool scaled, Scale s,
                       const int32_t* queue /*new->old or null*/, const int32_t* rmap /*old->new or null*/,
                       PinnedBuf* stage_val = nullptr, PinnedBuf* stage_col = nullptr) {
    using R = typename real_of<T>::type;
    R const sa = static_cast<R>(s.a), sb = scaled ? static_cast<R>(s.b) : R{0};
    R const f = scaled ? R{2} / sa : R{1};
    bool const reordered = queue != nullptr;

    // ELL width: row length plus one where a diagonal has to be created by the b-offset
    std::vector<int> kmax_part(64, 0);
    int kmax = 0;
    {
        std::mutex m;
        parallel_rows(n, [&](int64_t b, int64_t e) {
            int local = 0;
            for (int64_t row = b; row < e; ++row) {
                int cnt = indptr[row + 1] - indptr[row];
                if (sb != R{0}) {
                    bool has_diag = false;
                    for (int p = indptr[row]; p < indptr[row + 1]; ++p) has_diag |= (indices[p] == row);
                    if (!has_diag) ++cnt;
                }
                local = std::max(local, cnt);
            }
            std::lock_guard<std::mutex> lk(m);
            kmax = std::max(kmax, local);
        });
    }
    HostEll ell;
    ell.k = std::max(kmax, 1);
    ell.pitch = (n + 31) / 32 * 32;
    ell.val_bytes = static_cast<size_t>(ell.k) * ell.pitch * sizeof(T);
    ell.col_count = static_cast<size_t>(ell.k) * ell.pitch;
    // every element is written below (entries, padding, tail rows): no zero-fill pass; page-locked staging when offered
    ell.val = stage_val ? static_cast<char*>(stage_val->ensure(ell.val_bytes)) : nullptr;
    ell.col = stage_col ? static_cast<int32_t*>(stage_col->ensure(ell.col_count * sizeof(int32_t))) : nullptr;
    if (!ell.val) { ell.val_own.resize_uninit(ell.val_bytes); ell.val = ell.val_own.data(); }
    if (!ell.col) { ell.col_own.resize_uninit(ell.col_count); ell.col = ell.col_own.data(); }
    T* val = reinterpret_cast<T*>(ell.val);
    int32_t* col = ell.col;
    int const k = ell.k;
    int64_t const pitch = ell.pitch;

    parallel_rows(n, [&](int64_t b, int64_t e) {
        std::vector<std::pair<int32_t, T>> buf;
        for (int64_t new_row = b; new_row < e; ++new_row) {
            int64_t const row = reordered ? queue[new_row] : new_row;
            buf.clear();
            bool diag_done = (sb == R{0});
            for (int p = indptr[row]; p < indptr[row + 1]; ++p) {
                int32_t const c = indices[p];
                T v = data[p];
                if (scaled) {
                    if (c == row && sb != R{0}) {
                        v = reordered ? v * f - sb * f : (v - sb) * f;  // :124-127 vs :61-66
                        diag_done = true;
                    } else {
                        v = v * f;
                    }
                }
                buf.emplace_back(reordered ? rmap[c] : c, v);
            }
            if (!diag_done) buf.emplace_back(static_cast<int32_t>(new_row), reordered ? T{-sb * f} : (T{0} - T{sb}) * f);
            std::sort(buf.begin(), buf.end(), [](auto const& l, auto const& r) { return l.first < r.first; });
            int sidx = 0;
            for (auto const& en : buf) { val[sidx * pitch + new_row] = en.second; col[sidx * pitch + new_row] = en.first; ++sidx; }
            for (; sidx < k; ++sidx) { val[sidx * pitch + new_row] = T{0}; col[sidx * pitch + new_row] = static_cast<int32_t>(new_row); }
        }
        (void)k;
    });
    for (int sidx = 0; sidx < k; ++sidx) for (int64_t r = n; r < pitch; ++r) { val[sidx * pitch + r] = T{0}; col[sidx * pitch + r] = 0; }
    return ell;
}

} // anonymous namespace

/// Host-only test hook (pbk_host_ell): the scaled slot-major ELL exactly as build_device_hamiltonian uploads it, for the
/// caller's order (`order` == nullptr: create_scaled semantics) or a relabelled one (`order[new] = old`: create_reordered).
int host_scaled_ell(int dtype, int64_t n, const int32_t* indptr, const int32_t* indices, const void* data, double min_energy,
                    double max_energy, const int32_t* order, int32_t* k_out, int64_t* pitch_out, void* val, int32_t* col) {
    // the energy range reaches the engine as floats (pbk_config / kpm::Config): round the same way
    Scale const s(static_cast<float>(min_energy), static_cast<float>(max_energy));
    std::vector<int32_t> rmap;
    if (order) {
        rmap.assign(static_cast<size_t>(n), -1);
        for (int64_t i = 0; i < n; ++i) {
            if (order[i] < 0 || order[i] >= n || rmap[order[i]] >= 0) return PBK_INVALID_ARGUMENT;   // not a permutation
            rmap[order[i]] = static_cast<int32_t>(i);
        }
    }
    const int32_t* rm = order ? rmap.data() : nullptr;
    HostEll ell;
    switch (dtype) {
        case F32: ell = build_ell_host<float>(n, indptr, indices, static_cast<const float*>(data), true, s, order, rm); break;
        case C64: ell = build_ell_host<cf>(n, indptr, indices, static_cast<const cf*>(data), true, s, order, rm); break;
        case F64: ell = build_ell_host<double>(n, indptr, indices, static_cast<const double*>(data), true, s, order, rm); break;
        case C128: ell = build_ell_host<cd>(n, indptr, indices, static_cast<const cd*>(data), true, s, order, rm); break;
        default: return PBK_INVALID_ARGUMENT;
    }
    *k_out = ell.k;
    *pitch_out = ell.pitch;
    if (val && col) {
        std::memcpy(val, ell.val, ell.val_bytes);
        std::memcpy(col, ell.col, ell.col_count * sizeof(int32_t));
    }
    return PBK_OK;
}

/// Locality ordering for the stochastic (full-system) quantities: the sites are relabelled cluster by cluster,
/// each cluster a breadth-first ball of at most `tile` sites grown from a seed on the frontier of the clusters
/// made so far.  A CTA of the step kernel works through one tile of consecutive rows at a time, so the x-rows
/// it gathers (the ball and a thin halo) stay in that SM's L1 and every x element comes from HBM once.
/// KPM results are invariant under a relabelling of the sites; the starters are generated per *original* site
/// index and scattered through the map, exactly like the reference does with its own reorder map
/// (cppcore/src/kpm/Starter.cpp:68,80).  Fills queue (new -> old) and rmap (old -> new).
void cluster_order_flat(int64_t n, const int32_t* indptr, const int32_t* indices, int64_t tile,
                        std::vector<int32_t>& queue, std::vector<int32_t>& rmap) {
    queue.clear();
    queue.reserve(n);
    rmap.assign(n, -1);
    std::vector<int32_t> seeds;
    size_t seed_head = 0;
    int64_t next_unvisited = 0;
    while (static_cast<int64_t>(queue.size()) < n) {
        int32_t seed = -1;
        while (seed_head < seeds.size()) {
            int32_t const c = seeds[seed_head++];
            if (rmap[c] < 0) { seed = c; break; }
        }
        if (seed_head > (size_t{1} << 22) && seed_head * 2 > seeds.size()) {  // drop the consumed part of the FIFO
            seeds.erase(seeds.begin(), seeds.begin() + static_cast<std::ptrdiff_t>(seed_head));
            seed_head = 0;
        }
        if (seed < 0) {
            while (rmap[next_unvisited] >= 0) ++next_unvisited;
            seed = static_cast<int32_t>(next_unvisited);
        }
        size_t const begin = queue.size();
        size_t const limit = begin + static_cast<size_t>(tile);
        rmap[seed] = static_cast<int32_t>(queue.size());
        queue.push_back(seed);
        size_t head = begin;
        bool full = false;
        while (head < queue.size() && !full) {
            int32_t const row = queue[head];
            if (head + 4 < queue.size()) {   // the queue is the access pattern: pull the adjacency of a later row into cache
                int32_t const ahead = queue[head + 4];
                __builtin_prefetch(indices + indptr[ahead]);
            }
            int const pend = indptr[row + 1];
            for (int p = indptr[row]; p < pend; ++p) {
                int32_t const c = indices[p];
                if (rmap[c] >= 0) continue;
                if (queue.size() >= limit) { full = true; break; }
                rmap[c] = static_cast<int32_t>(queue.size());
                queue.push_back(c);
                __builtin_prefetch(indptr + c);
            }
            if (!full) ++head;
        }
        for (size_t q = head; q < queue.size(); ++q) {  // unvisited neighbours of the ball's surface seed later balls
            int32_t const row = queue[q];
            for (int p = indptr[row]; p < indptr[row + 1]; ++p) if (rmap[indices[p]] < 0) seeds.push_back(indices[p]);
        }
    }
}

/// Two-level locality ordering: macro-blocks of `macro_tiles` tiles, then the clusters inside each block, blocks in
/// parallel.  The tiles of a macro-block are consecutive rows, so the CTAs that are resident together work on
/// neighbouring clusters and a halo row fetched by one of them is an L2 hit for the others; only the halo of the
/// macro-block boundary is read from DRAM twice.  macro_tiles <= 1: one level.
///
/// The macro-blocks are breadth-first balls too, but of a *coarsened* graph on large systems: `coarse` consecutive sites
/// of the caller's order form one super-node (lattice generators number neighbouring cells consecutively; for an
/// arbitrary order the blocks are merely less compact, the clusters inside them are still grown on the real graph).
/// The coarse graph is built by a parallel pass over the CSR and its ball growing is `coarse` times cheaper than the
/// serial pass over all sites, which was the largest host cost of `set_hamiltonian`.  Every step is deterministic.
void cluster_order(int64_t n, const int32_t* indptr, const int32_t* indices, int64_t tile,
                   std::vector<int32_t>& queue, std::vector<int32_t>& rmap, int64_t macro_tiles, int64_t coarse) {
    int64_t const macro = macro_tiles * tile;
    if (macro_tiles <= 1 || macro >= n) { cluster_order_flat(n, indptr, indices, tile, queue, rmap); return; }
    bool const timing = std::getenv("PBK_TIMING") != nullptr;
    double t_mark = now_seconds();
    auto mark = [&](char const* what) {
        if (!timing) return;
        double const t = now_seconds();
        std::fprintf(stderr, "[pbkpm] cluster_order: %-24s %.3f s\n", what, t - t_mark);
        t_mark = t;
    };

    // ---- level 1: sites in block order (lvl1), block borders (bstart), block of every site (blk) ----
    // (large work arrays are left uninitialised: a value-initialising std::vector would zero -- and first-touch -- 150 MB
    // each on this one thread; every entry is written by the parallel passes below)
    std::vector<int32_t> lvl1_vec;
    RawVec<int32_t> lvl1_raw, blk_raw;
    blk_raw.resize_uninit(static_cast<size_t>(n));
    int32_t* const blk = blk_raw.data();
    const int32_t* lvl1 = nullptr;
    std::vector<int64_t> bstart;
    bool const use_coarse = coarse > 1 && macro % coarse == 0 && macro / coarse >= 1024 && n >= 8 * macro;   // blocks of >= 1024 super-nodes stay compact
    if (!use_coarse) {
        std::vector<int32_t> r1;
        cluster_order_flat(n, indptr, indices, macro, lvl1_vec, r1);
        lvl1 = lvl1_vec.data();
        for (int64_t b = 0; b * macro < n; ++b) bstart.push_back(b * macro);
        bstart.push_back(n);
        parallel_rows(n, [&](int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) blk[i] = static_cast<int32_t>(r1[i] / macro); });
    } else {
        int64_t const ns = (n + coarse - 1) / coarse;       // super-node s = sites [s * coarse, (s + 1) * coarse)
        int64_t const mc = macro / coarse;                  // super-nodes per block
        // coarse adjacency with a fixed stride: at most CAP distinct neighbouring super-nodes, in first-occurrence order
        // (deterministic), unused slots point to the node itself (ignored by the ball growing)
        constexpr int CAP = 16;
        if (ns * CAP >= (int64_t{1} << 31)) { cluster_order(n, indptr, indices, tile, queue, rmap, macro_tiles, 1); return; }
        RawVec<int32_t> cptr, cidx;
        cptr.resize_uninit(static_cast<size_t>(ns) + 1);
        cidx.resize_uninit(static_cast<size_t>(ns) * CAP);
        std::atomic<bool> overflow{false};
        int cshift = -1;   // super-nodes of 2^cshift sites: a shift instead of a 64-bit division per matrix element
        if ((coarse & (coarse - 1)) == 0) { cshift = 0; while ((int64_t{1} << cshift) < coarse) ++cshift; }
        parallel_rows(ns, [&](int64_t b, int64_t e) {
            for (int64_t sn = b; sn < e; ++sn) {
                int32_t* const out = cidx.data() + sn * CAP;
                int cnt = 0;
                int64_t const lo = sn * coarse, hi = std::min<int64_t>(n, lo + coarse);
                for (int p = indptr[lo]; p < indptr[hi]; ++p) {
                    int32_t const c = cshift >= 0 ? static_cast<int32_t>(static_cast<uint32_t>(indices[p]) >> cshift)
                                                  : static_cast<int32_t>(indices[p] / coarse);
                    if (c == sn) continue;
                    bool seen = false;
                    for (int q = 0; q < cnt; ++q) seen |= (out[q] == c);
                    if (seen) continue;
                    if (cnt == CAP) { overflow = true; break; }
                    out[cnt++] = c;
                }
                for (int q = cnt; q < CAP; ++q) out[q] = static_cast<int32_t>(sn);
                cptr.data()[sn] = static_cast<int32_t>(sn * CAP);
            }
        });
        cptr.data()[ns] = static_cast<int32_t>(ns * CAP);
        if (overflow) {   // denser than a lattice: grow the macro-blocks on the real graph instead
            cluster_order(n, indptr, indices, tile, queue, rmap, macro_tiles, 1);
            return;
        }
        mark("coarse graph");
        std::vector<int32_t> q1c, r1c;
        cluster_order_flat(ns, cptr.data(), cidx.data(), mc, q1c, r1c);
        mark("macro-blocks (coarse)");
        // blocks of mc super-nodes in q1c order; the block holding the short last super-node goes to the end so that every
        // other block starts on a tile boundary
        int64_t const nb = (ns + mc - 1) / mc;
        std::vector<int64_t> order(nb);
        for (int64_t b = 0; b < nb; ++b) order[b] = b;
        if (n % coarse != 0) {
            int64_t const odd = r1c[ns - 1] / mc;
            order.erase(order.begin() + odd);
            order.push_back(odd);
        }
        lvl1_raw.resize_uninit(static_cast<size_t>(n));
        int32_t* const lvl1_out = lvl1_raw.data();
        lvl1 = lvl1_out;
        bstart.assign(1, 0);
        for (int64_t ob = 0; ob < nb; ++ob) {   // block sizes first, so that the expansion below can run block-parallel
            int64_t const b = order[ob];
            int64_t size = 0;
            for (int64_t j = b * mc; j < std::min<int64_t>(ns, (b + 1) * mc); ++j) {
                int64_t const sn = q1c[j];
                size += std::min<int64_t>(n, (sn + 1) * coarse) - sn * coarse;
            }
            bstart.push_back(bstart.back() + size);
        }
        parallel_rows(nb, [&](int64_t b0, int64_t b1) {
            for (int64_t ob = b0; ob < b1; ++ob) {
                int64_t const b = order[ob];
                int64_t pos = bstart[ob];
                for (int64_t j = b * mc; j < std::min<int64_t>(ns, (b + 1) * mc); ++j) {
                    int64_t const sn = q1c[j];
                    for (int64_t i = sn * coarse; i < std::min<int64_t>(n, (sn + 1) * coarse); ++i) { lvl1_out[pos++] = static_cast<int32_t>(i); blk[i] = static_cast<int32_t>(ob); }
                }
            }
        }, 2);
    }

    mark("level-1 expansion");
    // ---- level 2: clusters inside every block, grown on the real graph; blocks are independent ----
    if (static_cast<int64_t>(queue.size()) != n) queue.assign(n, 0);   // every entry is written below
    if (static_cast<int64_t>(rmap.size()) != n) rmap.resize(n);
    parallel_rows(n, [&](int64_t b, int64_t e) { std::fill(rmap.begin() + b, rmap.begin() + e, -1); });
    mark("output arrays");
    int64_t const nblocks = static_cast<int64_t>(bstart.size()) - 1;
    auto order_block = [&](int64_t m) {
        int64_t const lo = bstart[m], hi = bstart[m + 1];
        int32_t const me = static_cast<int32_t>(m);
        int64_t filled = lo;             // next position of the final order
        std::vector<int32_t> seeds;
        size_t seed_head = 0;
        int64_t scan = lo;               // next unvisited site of the block in level-1 order
        while (filled < hi) {
            int32_t seed = -1;
            while (seed_head < seeds.size()) { int32_t const c = seeds[seed_head++]; if (rmap[c] < 0) { seed = c; break; } }
            if (seed < 0) { while (rmap[lvl1[scan]] >= 0) ++scan; seed = lvl1[scan]; }
            int64_t const begin = filled;
            int64_t const limit = std::min<int64_t>(hi, (begin - lo) / tile * tile + lo + tile);   // tiles stay aligned inside the block
            rmap[seed] = static_cast<int32_t>(filled); queue[filled++] = seed;
            int64_t head = begin;
            bool full = filled >= limit;
            while (head < filled && !full) {
                int32_t const row = queue[head];
                for (int p = indptr[row]; p < indptr[row + 1]; ++p) {
                    int32_t const c = indices[p];
                    if (blk[c] != me || rmap[c] >= 0) continue;
                    if (filled >= limit) { full = true; break; }
                    rmap[c] = static_cast<int32_t>(filled); queue[filled++] = c;
                }
                if (!full) ++head;
            }
            for (int64_t q = head; q < filled; ++q) {
                int32_t const row = queue[q];
                for (int p = indptr[row]; p < indptr[row + 1]; ++p) { int32_t const c = indices[p]; if (blk[c] == me && rmap[c] < 0) seeds.push_back(c); }
            }
        }
    };
    std::atomic<int64_t> next{0};
    run_pool(host_threads(), [&](int) { for (int64_t m = next++; m < nblocks; m = next++) order_block(m); });
    mark("clusters inside blocks");
}

/// BFS relabelling from src[0]; slice k = the k-th shell (OptimizedHamiltonian.cpp:88-143)
BfsOrder Engine::bfs_order(Indices const& target) const {
    BfsOrder b;
    b.target = target;
    auto& queue = b.queue;
    queue.reserve(n);
    queue.push_back(target.src[0]);
    b.reorder_map.assign(n, -1);
    b.reorder_map[target.src[0]] = 0;
    std::vector<int32_t> borders{1};
    for (int64_t h2_row = 0; h2_row < n; ++h2_row) {
        if (h2_row >= static_cast<int64_t>(queue.size())) {
            throw Error(PBK_RUNTIME_ERROR, "KPM: the Hamiltonian graph is not connected; the optimal_size "
                                           "reordering needs a connected system");
        }
        int32_t const row = queue[h2_row];
        for (int p = h_indptr[row]; p < h_indptr[row + 1]; ++p) {
            int32_t const c = h_indices[p];
            if (b.reorder_map[c] < 0) { b.reorder_map[c] = static_cast<int32_t>(queue.size()); queue.push_back(c); }
        }
        if (h2_row == borders.back() - 1) borders.push_back(static_cast<int32_t>(queue.size()));
    }
    borders.pop_back();
    for (int32_t i : target.src) b.idx.src.push_back(b.reorder_map[i]);
    for (int32_t i : target.dest) b.idx.dest.push_back(b.reorder_map[i]);
    auto find_offset = [&](std::vector<int32_t> const& v) {
        int32_t const mx = *std::max_element(v.begin(), v.end());
        auto const it = std::find_if(borders.begin(), borders.end(), [&](int32_t bd) { return bd > mx; });
        return static_cast<int>(it - borders.begin());
    };
    b.map.src_offset = find_offset(b.idx.src);
    b.map.dest_offset = find_offset(b.idx.dest);
    b.map.data = std::move(borders);
    b.valid = true;
    return b;
}

/// Rows of `dh` (order maps already on the device in dh.queue_dev / dh.perm, or none) from the resident CSR by one
/// kernel (build.cu): mode = BUILD_SCALED (H~), BUILD_PLAIN (Lanczos) or BUILD_VELOCITY (operator of Kubo-Bastin).
bool Engine::build_layout_on_device(DeviceHamiltonian& dh, int mode, Scale s, const float* positions_dev) {
    if (!dev_build || !dev_csr) return false;
    using R32 = float;
    bool const single = dtype == F32 || dtype == C64;
    // the scalar map in the Hamiltonian's real type, like build_ell_host: f = 2 / a, sb = b
    double f = 1.0, sb = 0.0;
    if (mode == BUILD_SCALED) {
        if (single) { R32 const sa = static_cast<R32>(s.a); f = static_cast<double>(R32{2} / sa); sb = static_cast<double>(static_cast<R32>(s.b)); }
        else { f = 2.0 / s.a; sb = s.b; }
    }
    PBK_CUDA(launch_row_width(d_indptr.as<int32_t>(), d_indices.as<int32_t>(), n, mode == BUILD_SCALED && sb != 0.0, width_dev.as<int>(), stream));
    int width = 0;
    PBK_CUDA(cudaMemcpyAsync(&width, width_dev.as(), sizeof(int), cudaMemcpyDeviceToHost, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
    int const k = std::max(width, 1);
    if (k > build_max_width()) return false;
    int64_t const pitch = (n + 31) / 32 * 32;
    dh.val.alloc(static_cast<size_t>(k) * pitch * dtype_size(dtype));
    dh.col.alloc(static_cast<size_t>(k) * pitch * sizeof(int32_t));
    BuildArgs a;
    a.indptr = d_indptr.as<int32_t>(); a.indices = d_indices.as<int32_t>(); a.n = n;
    a.queue = dh.reordered ? dh.queue_dev.as<int32_t>() : nullptr;
    a.perm = dh.reordered ? dh.perm.as<int32_t>() : nullptr;
    a.mode = mode; a.f = f; a.sb = sb; a.positions = positions_dev;
    a.k = k; a.pitch = pitch; a.col = dh.col.as<int32_t>();
    PBK_CUDA(launch_csr_to_ell(dtype, a, d_data.as(), dh.val.as(), stream));
    launches += 2;
    dh.ell = EllDev{dh.val.as(), dh.col.as<int32_t>(), n, pitch, k};
    return true;
}

/// Tiles, halo lists and local codes of the resident-tile step kernel for passes of R vectors (kernels_res.cu).  The
/// nominal tiles are the locality clusters; a tile whose own + halo rows exceed the shared-memory capacity is halved
/// (a few passes of the counting kernel), then one kernel writes the sorted halo lists and the 16-bit codes.
bool Engine::ensure_res_meta(DeviceHamiltonian& dh, int R) {
    if (!dh.res_enabled || !dh.valid) return false;
    ResGeometry const geo = res_geometry(dtype, dh.ell.k, R, res_ctas, res_stages, res_buffers);
    if (dh.res_ntiles > 0 && dh.res_geo.row_bytes == geo.row_bytes) return true;
    if (dh.res_failed_row_bytes == geo.row_bytes) return false;
    auto fail = [&]() { dh.res_failed_row_bytes = geo.row_bytes; dh.res_ntiles = 0; return false; };
    int const rpb = geo.rows_per_iteration;
    if (rpb < 1 || geo.cap_rows < 2 * rpb || dh.tile % rpb != 0 || dh.ell.k > 64) return fail();
    double const t0 = now_seconds();
    std::vector<ResTile> tiles;
    for (int64_t r0 = 0; r0 < n; r0 += dh.tile) tiles.push_back(ResTile{static_cast<int32_t>(r0), static_cast<int32_t>(std::min<int64_t>(dh.tile, n - r0)), 0, 0});
    std::vector<int32_t> nh;
    DevBuf tiles_dev, nh_dev;
    for (int pass = 0; pass < 8; ++pass) {
        tiles_dev.ensure(sizeof(ResTile) * tiles.size());
        nh_dev.ensure(sizeof(int32_t) * tiles.size());
        nh.resize(tiles.size());
        PBK_CUDA(cudaMemcpyAsync(tiles_dev.as(), tiles.data(), sizeof(ResTile) * tiles.size(), cudaMemcpyHostToDevice, stream));
        PBK_CUDA(launch_res_count(dh.ell, tiles_dev.as<ResTile>(), static_cast<int>(tiles.size()), nh_dev.as<int32_t>(), stream));
        PBK_CUDA(cudaMemcpyAsync(nh.data(), nh_dev.as(), sizeof(int32_t) * tiles.size(), cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
        ++launches;
        std::vector<ResTile> next;
        next.reserve(tiles.size() + tiles.size() / 4);
        bool split = false;
        for (size_t t = 0; t < tiles.size(); ++t) {
            ResTile tl = tiles[t];
            tl.nh = nh[t];
            if (tl.nrows + tl.nh <= geo.cap_rows && tl.nh <= res_max_halo()) { next.push_back(tl); continue; }
            if (tl.nrows < 2 * rpb) return fail();       // even the smallest tile does not fit
            int32_t const first = (tl.nrows / 2 + rpb - 1) / rpb * rpb;
            next.push_back(ResTile{tl.row0, first, 0, -1});
            next.push_back(ResTile{tl.row0 + first, tl.nrows - first, 0, -1});
            split = true;
        }
        tiles.swap(next);
        if (!split) break;
        if (pass == 7) return fail();
    }
    int64_t halo_total = 0;
    for (auto& tl : tiles) { tl.halo_off = static_cast<int32_t>(halo_total); halo_total += tl.nh; }
    if (halo_total >= (int64_t{1} << 31)) return fail();
    dh.res_tiles.ensure(sizeof(ResTile) * tiles.size());
    dh.res_halo.ensure(sizeof(int32_t) * static_cast<size_t>(std::max<int64_t>(halo_total, 1)));
    dh.res_codes.ensure(static_cast<size_t>(n) * geo.cb);
    dh.res_vals.ensure(static_cast<size_t>(n) * geo.kvb);
    PBK_CUDA(cudaMemcpyAsync(dh.res_tiles.as(), tiles.data(), sizeof(ResTile) * tiles.size(), cudaMemcpyHostToDevice, stream));
    PBK_CUDA(launch_res_fill(dtype, dh.ell, dh.res_tiles.as<ResTile>(), static_cast<int>(tiles.size()), dh.res_halo.as<int32_t>(),
                             dh.res_codes.as(), dh.res_vals.as(), geo, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
    ++launches;
    dh.res_ntiles = static_cast<int>(tiles.size());
    dh.res_geo = geo;
    dh.res_halo_frac = static_cast<double>(halo_total) / static_cast<double>(n);
    if (std::getenv("PBK_TIMING")) std::fprintf(stderr, "[pbkpm] resident-tile metadata: %zu tiles (nominal %lld rows, capacity %d rows of %u bytes), halo %.2f x rows, %.3f s\n",
                                                 tiles.size(), static_cast<long long>(dh.tile), geo.cap_rows, geo.row_bytes, dh.res_halo_frac, now_seconds() - t0);
    return true;
}

/// host copies of a layout's order maps: they exist already when the host computed the order, and are downloaded on
/// first use when the order arrived by broadcast (only index look-ups and host-built operators need them)
void Engine::ensure_host_order(DeviceHamiltonian& dh) {
    if (!dh.reordered || dh.host_order) return;
    PBK_CUDA(cudaSetDevice(device));
    dh.reorder_map.resize(static_cast<size_t>(n));
    dh.order_queue.resize(static_cast<size_t>(n));
    PBK_CUDA(cudaMemcpyAsync(dh.reorder_map.data(), dh.perm.as(), sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream));
    PBK_CUDA(cudaMemcpyAsync(dh.order_queue.data(), dh.queue_dev.as(), sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
    stats.d2h_bytes += static_cast<int64_t>(2 * sizeof(int32_t) * n);
    dh.host_order = true;
}

void Engine::broadcast(void* dev, int64_t count_int32, int root) {
    if (world <= 1 || !comm) return;
    nccl->check(nccl->Broadcast(dev, dev, static_cast<size_t>(count_int32), /*ncclInt32*/ 2, root, comm, stream), "ncclBroadcast");
}

void Engine::build_device_hamiltonian(DeviceHamiltonian& dh, bool scaled, int order, Indices const& target) {
    require_hamiltonian();
    PBK_CUDA(cudaSetDevice(device));
    double const t0 = now_seconds();
    Scale const s = scaled ? scaling_factors() : Scale();
    dh = DeviceHamiltonian();
    bool const reorder = order != ORDER_NATURAL;

    bool const timing = std::getenv("PBK_TIMING") != nullptr;
    double t_mark = now_seconds();
    auto mark = [&](char const* what) {
        if (!timing) return;
        double const t = now_seconds();
        std::fprintf(stderr, "[pbkpm] build_device_hamiltonian: %-28s %.3f s\n", what, t - t_mark);
        t_mark = t;
    };
    // ---- row order: host vectors (queue: new -> old, dh.reorder_map: old -> new) and / or the device copy ----
    std::vector<int32_t> queue;
    bool order_on_device = false;
    if (order == ORDER_CLUSTER) {
        if (identity_order) {   // PBK_IDENTITY_ORDER=1: the caller's site order, cut into tiles of consecutive rows (staged kernel applies)
            queue.resize(n);
            for (int64_t i = 0; i < n; ++i) queue[i] = static_cast<int32_t>(i);
            dh.reorder_map = queue;
        } else if (cluster_tile == layout_tile && (cluster_on_host || cluster_on_device)) {   // computed / received by set_hamiltonian
            if (cluster_on_host) { queue.swap(cluster_queue); dh.reorder_map.swap(cluster_rmap); }
            else dh.host_order = false;
            if (cluster_on_device) { dh.queue_dev = std::move(cluster_queue_dev); order_on_device = true; }
            cluster_tile = 0; cluster_on_host = cluster_on_device = false;
        } else {
            cluster_order(n, h_indptr.data(), h_indices.data(), layout_tile, queue, dh.reorder_map,
                          std::max<int64_t>(1, macro_tiles * locality_tile / layout_tile), coarse_sites);
        }
        dh.idx = target;   // positions in this layout are looked up where they are needed (ensure_host_order)
        if (dh.host_order) { dh.idx = Indices{}; for (int32_t i : target.src) dh.idx.src.push_back(dh.reorder_map[i]); for (int32_t i : target.dest) dh.idx.dest.push_back(dh.reorder_map[i]); }
        dh.map.data = {static_cast<int32_t>(n)};
        dh.reordered = true;
        dh.tile = identity_order ? locality_tile : layout_tile;
        dh.res_enabled = layout_res && !identity_order;
    } else if (order == ORDER_BFS) {
        if (bfs_ready.valid_for(target)) {  // moments_ldos already ran the relabelling to look at the slice map
            queue = std::move(bfs_ready.queue);
            dh.reorder_map = std::move(bfs_ready.reorder_map);
            dh.idx = bfs_ready.idx;
            dh.map = bfs_ready.map;
            bfs_ready = BfsOrder();
        } else {
            BfsOrder b = bfs_order(target);
            queue = std::move(b.queue);
            dh.reorder_map = std::move(b.reorder_map);
            dh.idx = b.idx;
            dh.map = b.map;
        }
        dh.reordered = true;
        dh.sliced = true;
    } else {
        dh.idx = target;
        dh.map.data = {static_cast<int32_t>(n)};
    }
    mark("row order");

    // ---- device path: the rows are written by a kernel from the resident CSR ----
    bool built = false;
    if (dev_build && dev_csr) {
        if (reorder) {
            if (!order_on_device) {
                dh.queue_dev.alloc(sizeof(int32_t) * static_cast<size_t>(n));
                PBK_CUDA(cudaMemcpyAsync(dh.queue_dev.as(), queue.data(), sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
                stats.h2d_bytes += static_cast<int64_t>(sizeof(int32_t) * n);
            }
            dh.perm.alloc(sizeof(int32_t) * static_cast<size_t>(n));
            PBK_CUDA(launch_invert_order(dh.queue_dev.as<int32_t>(), n, dh.perm.as<int32_t>(), stream));
            ++launches;
        }
        built = build_layout_on_device(dh, scaled ? BUILD_SCALED : BUILD_PLAIN, s, nullptr);
        if (built) mark("rows built on the device");
    }
    if (!built) {   // host path: scaled ELL written by the host threads into page-locked staging, one upload
        if (reorder && !dh.host_order) ensure_host_order(dh);
        if (reorder && queue.empty()) queue = dh.order_queue;
        HostEll ell;
        const int32_t* q = reorder ? queue.data() : nullptr;
        const int32_t* rm = reorder ? dh.reorder_map.data() : nullptr;
        // the full-system layout (the big, long-lived one) is staged in the context's page-locked buffers
        PinnedBuf* const sv = (&dh == &natural) ? &stage_val : nullptr;
        PinnedBuf* const sc = (&dh == &natural) ? &stage_col : nullptr;
        switch (dtype) {
            case F32: ell = build_ell_host<float>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const float*>(h_data.data()), scaled, s, q, rm, sv, sc); break;
            case C64: ell = build_ell_host<cf>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const cf*>(h_data.data()), scaled, s, q, rm, sv, sc); break;
            case F64: ell = build_ell_host<double>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const double*>(h_data.data()), scaled, s, q, rm, sv, sc); break;
            default: ell = build_ell_host<cd>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const cd*>(h_data.data()), scaled, s, q, rm, sv, sc); break;
        }
        mark("scaled ELL on the host");
        dh.val.alloc(ell.val_bytes);
        dh.col.alloc(ell.col_count * sizeof(int32_t));
        PBK_CUDA(cudaMemcpyAsync(dh.val.as(), ell.val, ell.val_bytes, cudaMemcpyHostToDevice, stream));
        PBK_CUDA(cudaMemcpyAsync(dh.col.as(), ell.col, ell.col_count * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += static_cast<int64_t>(ell.val_bytes + ell.col_count * sizeof(int32_t));
        if (reorder) {
            if (!dh.perm.bytes()) dh.perm.alloc(sizeof(int32_t) * static_cast<size_t>(n));
            PBK_CUDA(cudaMemcpyAsync(dh.perm.as(), dh.reorder_map.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream));
            stats.h2d_bytes += static_cast<int64_t>(sizeof(int32_t) * n);
        }
        dh.ell = EllDev{dh.val.as(), dh.col.as<int32_t>(), n, ell.pitch, ell.k};
        PBK_CUDA(cudaStreamSynchronize(stream));   // `ell` may live in pageable memory that goes out of scope here
    }
    if (order == ORDER_CLUSTER) dh.order_queue = std::move(queue);  // operators of the same calculation are laid out alike
    if (order == ORDER_CLUSTER && bulk_stages >= 2) {  // granule-packed records for the bulk-copy staged step kernel
        dh.packed.alloc(packed_ell_bytes(dtype, dh.ell));
        PBK_CUDA(launch_pack_ell(dtype, dh.ell, dh.packed.as(), stream));
        ++launches;
    }
    PBK_CUDA(cudaStreamSynchronize(stream));
    mark("upload / pack");
    dh.original_idx = target;
    dh.valid = true;
    dh.seconds = now_seconds() - t0;
}

DeviceHamiltonian& Engine::natural_hamiltonian() {
    if (!natural.valid) build_device_hamiltonian(natural, true, locality_tile > 0 ? ORDER_CLUSTER : ORDER_NATURAL, Indices{{0}, {0}});
    return natural;
}

DeviceHamiltonian& Engine::optimized_for(Indices const& target) {
    if (!config.optimal_size) {  // no light-cone slicing requested: the natural order serves every index
        auto& h = natural_hamiltonian();
        ensure_host_order(h);
        h.idx = Indices{};
        for (int32_t i : target.src) h.idx.src.push_back(h.reordered ? h.reorder_map[i] : i);
        for (int32_t i : target.dest) h.idx.dest.push_back(h.reordered ? h.reorder_map[i] : i);
        return h;
    }
    if (!(optimized.valid && optimized.original_idx == target)) {  // OptimizedHamiltonian.cpp:43-45
        optimized = DeviceHamiltonian();
        build_device_hamiltonian(optimized, true, ORDER_BFS, target);
    }
    return optimized;
}

DeviceHamiltonian& Engine::unscaled_hamiltonian() {
    if (!unscaled.valid) build_device_hamiltonian(unscaled, false, ORDER_NATURAL, Indices{{0}, {0}});
    return unscaled;
}

/// velocity operator V_ij = H_ij * (pos_i - pos_j) on the unscaled H (src/kpm/Moments.cpp:132-156), laid out like `like`
void Engine::upload_operator(DeviceHamiltonian& dh, const float* pos, DeviceHamiltonian& like) {
    dh = DeviceHamiltonian();
    dh.tile = like.tile;
    dh.map.data = {static_cast<int32_t>(n)};
    if (dev_build && dev_csr && (!like.reordered || (like.queue_dev.bytes() && like.perm.bytes()))) {
        // on the device from the resident CSR: only the coordinates go up
        DevBuf pos_dev(sizeof(float) * static_cast<size_t>(n));
        PBK_CUDA(cudaMemcpyAsync(pos_dev.as(), pos, sizeof(float) * static_cast<size_t>(n), cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += static_cast<int64_t>(sizeof(float) * n);
        dh.reordered = like.reordered;
        if (like.reordered) {   // borrow the order maps of the layout (device-to-device copies: the operator is short-lived)
            dh.queue_dev.alloc(like.queue_dev.bytes());
            dh.perm.alloc(like.perm.bytes());
            PBK_CUDA(cudaMemcpyAsync(dh.queue_dev.as(), like.queue_dev.as(), sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyDeviceToDevice, stream));
            PBK_CUDA(cudaMemcpyAsync(dh.perm.as(), like.perm.as(), sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyDeviceToDevice, stream));
        }
        bool const ok = build_layout_on_device(dh, BUILD_VELOCITY, Scale(), pos_dev.as<float>());
        PBK_CUDA(cudaStreamSynchronize(stream));
        dh.reordered = false;   // as before: the operator itself carries no order maps, it is only laid out like `like`
        dh.queue_dev = DevBuf(); dh.perm = DevBuf();
        if (ok) { dh.valid = true; return; }
    }
    ensure_host_order(like);
    int64_t const nnz = h_indptr[n];
    std::vector<char> data(static_cast<size_t>(nnz) * dtype_size(dtype));
    auto fill = [&](auto* out, auto const* in) {
        using T = std::remove_pointer_t<decltype(out)>;
        parallel_rows(n, [&](int64_t b, int64_t e) {
            for (int64_t row = b; row < e; ++row)
                for (int p = h_indptr[row]; p < h_indptr[row + 1]; ++p)
                    out[p] = in[p] * static_cast<T>(pos[row] - pos[h_indices[p]]);
        });
    };
    switch (dtype) {
        case F32: fill(reinterpret_cast<float*>(data.data()), reinterpret_cast<const float*>(h_data.data())); break;
        case C64: fill(reinterpret_cast<cf*>(data.data()), reinterpret_cast<const cf*>(h_data.data())); break;
        case F64: fill(reinterpret_cast<double*>(data.data()), reinterpret_cast<const double*>(h_data.data())); break;
        default: fill(reinterpret_cast<cd*>(data.data()), reinterpret_cast<const cd*>(h_data.data())); break;
    }
    HostEll ell;
    const int32_t* q = like.reordered ? like.order_queue.data() : nullptr;
    const int32_t* rm = like.reordered ? like.reorder_map.data() : nullptr;
    switch (dtype) {
        case F32: ell = build_ell_host<float>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const float*>(data.data()), false, Scale(), q, rm); break;
        case C64: ell = build_ell_host<cf>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const cf*>(data.data()), false, Scale(), q, rm); break;
        case F64: ell = build_ell_host<double>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const double*>(data.data()), false, Scale(), q, rm); break;
        default: ell = build_ell_host<cd>(n, h_indptr.data(), h_indices.data(), reinterpret_cast<const cd*>(data.data()), false, Scale(), q, rm); break;
    }
    dh.val.alloc(ell.val_bytes);
    dh.col.alloc(ell.col_count * sizeof(int32_t));
    PBK_CUDA(cudaMemcpyAsync(dh.val.as(), ell.val, ell.val_bytes, cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaMemcpyAsync(dh.col.as(), ell.col, ell.col_count * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));   // `ell` (pageable host memory) goes out of scope with this function
    stats.h2d_bytes += static_cast<int64_t>(ell.val_bytes + ell.col_count * sizeof(int32_t));
    dh.ell = EllDev{dh.val.as(), dh.col.as<int32_t>(), n, ell.pitch, ell.k};
    dh.valid = true;
}

/// generic operator of KPM.moments(op=...): c128 CSR cast to the Hamiltonian's scalar type (force_cast)
void Engine::upload_csr_operator(DeviceHamiltonian& dh, int64_t rows, const int32_t* indptr, const int32_t* indices, const cd* data,
                                 DeviceHamiltonian& like) {
    ensure_host_order(like);
    int64_t const nnz = indptr[rows];
    HostEll ell;
    const int32_t* q = like.reordered ? like.order_queue.data() : nullptr;
    const int32_t* rm = like.reordered ? like.reorder_map.data() : nullptr;
    switch (dtype) {
        case F32: { std::vector<float> d(nnz); for (int64_t i = 0; i < nnz; ++i) d[i] = static_cast<float>(data[i].real());
                    ell = build_ell_host<float>(rows, indptr, indices, d.data(), false, Scale(), q, rm); break; }
        case C64: { std::vector<cf> d(nnz); for (int64_t i = 0; i < nnz; ++i) d[i] = cf(static_cast<float>(data[i].real()), static_cast<float>(data[i].imag()));
                    ell = build_ell_host<cf>(rows, indptr, indices, d.data(), false, Scale(), q, rm); break; }
        case F64: { std::vector<double> d(nnz); for (int64_t i = 0; i < nnz; ++i) d[i] = data[i].real();
                    ell = build_ell_host<double>(rows, indptr, indices, d.data(), false, Scale(), q, rm); break; }
        default: ell = build_ell_host<cd>(rows, indptr, indices, data, false, Scale(), q, rm); break;
    }
    dh = DeviceHamiltonian();
    dh.tile = like.tile;
    dh.val.alloc(ell.val_bytes);
    dh.col.alloc(ell.col_count * sizeof(int32_t));
    PBK_CUDA(cudaMemcpyAsync(dh.val.as(), ell.val, ell.val_bytes, cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaMemcpyAsync(dh.col.as(), ell.col, ell.col_count * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));   // `ell` (pageable host memory) goes out of scope with this function
    stats.h2d_bytes += static_cast<int64_t>(ell.val_bytes + ell.col_count * sizeof(int32_t));
    dh.ell = EllDev{dh.val.as(), dh.col.as<int32_t>(), rows, ell.pitch, ell.k};
    dh.map.data = {static_cast<int32_t>(rows)};
    dh.valid = true;
}

// ------------------------------------------------------------------------------------------------
// Bounds: Lanczos on the device (compute/lanczos.hpp:101-154); the tiny tridiagonal problem stays on the host
// ------------------------------------------------------------------------------------------------
namespace {

/// extreme eigenvalues of the symmetric tridiagonal (alpha, beta) by Sturm-sequence bisection
void tridiagonal_minmax(std::vector<double> const& alpha, std::vector<double> const& beta, double* mn, double* mx) {
    int const m = static_cast<int>(alpha.size());
    double lo = alpha[0], hi = alpha[0];
    for (int i = 0; i < m; ++i) {
        double const r = (i > 0 ? std::abs(beta[i - 1]) : 0.0) + (i + 1 < m ? std::abs(beta[i]) : 0.0);
        lo = std::min(lo, alpha[i] - r);
        hi = std::max(hi, alpha[i] + r);
    }
    auto count_below = [&](double x) {  // number of eigenvalues < x
        int cnt = 0;
        double q = alpha[0] - x;
        if (q < 0) ++cnt;
        for (int i = 1; i < m; ++i) {
            double const d = (q == 0.0) ? 1e-300 : q;
            q = alpha[i] - x - beta[i - 1] * beta[i - 1] / d;
            if (q < 0) ++cnt;
        }
        return cnt;
    };
    auto kth = [&](int k) {  // k-th smallest eigenvalue (0-based)
        double a = lo, b = hi;
        for (int it = 0; it < 200 && b - a > 1e-15 * std::max(1.0, std::abs(a) + std::abs(b)); ++it) {
            double const mid = 0.5 * (a + b);
            if (count_below(mid) > k) b = mid; else a = mid;
        }
        return 0.5 * (a + b);
    };
    *mn = kth(0);
    *mx = kth(m - 1);
}

} // anonymous namespace

cudaError_t launch_uniform_transform(int dtype, const uint32_t* raw, int64_t n, void* dst, cudaStream_t s);

void Engine::compute_bounds() {
    if (have_bounds) return;
    require_hamiltonian();
    PBK_CUDA(cudaSetDevice(device));
    double const t0 = now_seconds();
    auto& h = unscaled_hamiltonian();
    size_t const vbytes = static_cast<size_t>(n) * dtype_size(dtype);
    DevBuf v0(vbytes), v1(vbytes), t(vbytes);
    int const w = dtype_words(dtype);
    raw.ensure(sizeof(uint32_t) * n * w);
    ensure_moment_buffers(1, 4);
    double* scal = mom.as<double>();  // [0..1]: <t|v1>, [2]: |v0|^2, [4]: |v1|^2
    scratch.ensure(sizeof(double) * 2 * num_sms * 8);

    // start vector: default-seeded uniform [0, 1) reals, normalised (lanczos.hpp:105-107)
    PBK_CUDA(launch_mt_seed(mt_state.as<uint32_t>(), stream));
    PBK_CUDA(launch_mt_generate(mt_state.as<uint32_t>(), raw.as<uint32_t>(), n * w, stream));
    PBK_CUDA(launch_uniform_transform(dtype, raw.as<uint32_t>(), n, v1.as(), stream));
    PBK_CUDA(cudaMemsetAsync(v0.as(), 0, vbytes, stream));
    PBK_CUDA(launch_dot_moment(dtype, v1.as(), v1.as(), n, scal, 2, 1.0, scratch.as<double>(), counter.as<unsigned>(), num_sms, stream));
    PBK_CUDA(launch_scale_inv_sqrt(dtype, n, v1.as(), scal + 4, stream));

    std::vector<double> alpha, beta;
    double previous_min = std::numeric_limits<double>::max();
    double previous_max = std::numeric_limits<double>::lowest();
    double const precision = static_cast<double>(config.lanczos_precision) / 100;
    void* p0 = v0.as();
    void* p1 = v1.as();
    for (int i = 0; i < 1000; ++i) {
        double const b_prev = beta.empty() ? 0.0 : beta.back();
        StepArgs a;
        a.h = h.ell; a.x = p1; a.y = t.as(); a.nrows = n; a.R = 1; a.subtract = false; a.scale = 1.0;
        PBK_CUDA(launch_step(dtype, a, num_sms, stream, nullptr));
        PBK_CUDA(launch_dot_moment(dtype, t.as(), p1, n, scal, 0, 1.0, scratch.as<double>(), counter.as<unsigned>(), num_sms, stream));
        PBK_CUDA(launch_lanczos_update(dtype, n, t.as(), p1, p0, b_prev, scal, scal + 2, scratch.as<double>(), counter.as<unsigned>(), num_sms, stream));
        PBK_CUDA(launch_scale_inv_sqrt(dtype, n, p0, scal + 2, stream));
        double host[3];
        PBK_CUDA(cudaMemcpyAsync(host, scal, sizeof(host), cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
        std::swap(p0, p1);
        alpha.push_back(host[0]);
        beta.push_back(std::sqrt(host[2]));

        double mn, mx;
        tridiagonal_minmax(alpha, beta, &mn, &mx);
        bool const conv_min = std::abs((previous_min - mn) / mn) < precision;
        bool const conv_max = std::abs((previous_max - mx) / mx) < precision;
        if (conv_min && conv_max) {
            bounds_min = mn; bounds_max = mx; lanczos_loops = i;
            have_bounds = true;
            bounds_seconds = now_seconds() - t0;
            return;
        }
        previous_min = mn; previous_max = mx;
    }
    throw Error(PBK_RUNTIME_ERROR, "Lanczos algorithm did not converge for the min/max eigenvalues.");
}

void Engine::bounds(double* mn, double* mx, int32_t* loops) {
    compute_bounds();
    *mn = bounds_min; *mx = bounds_max; *loops = lanczos_loops;
}

Scale Engine::scaling_factors() {
    compute_bounds();
    return Scale(bounds_min, bounds_max);
}

int Engine::required_num_moments(double broadening) {
    auto const s = scaling_factors();
    return kernel_required_num_moments(config.kernel, config.lambda_value, broadening / s.a);
}

// ------------------------------------------------------------------------------------------------
// Recursion drivers
// ------------------------------------------------------------------------------------------------
int Engine::lane_pad(int R) const {
    if (R == 1) return 1;
    int const vmax = 16 / dtype_size(dtype);
    return (R + vmax - 1) / vmax * vmax;
}

int Engine::pick_batch(int vectors, int extra_blocks) const {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += DevBuf::cached_bytes();   // blocks parked in the allocation cache are given back on demand
    // vec_a / vec_b (+ extra) blocks and the raw random words of one lane
    double const per_lane = static_cast<double>(n) * (dtype_size(dtype) * (2 + extra_blocks) + 4 * dtype_words(dtype));
    double const reusable = static_cast<double>(vec_a.bytes() + vec_b.bytes() + raw.bytes());
    int cap = static_cast<int>((0.85 * static_cast<double>(free_b) + reusable) / per_lane);
    int const hard = 4096 / dtype_size(dtype);  // 256 chunks of 16 bytes per row
    cap = std::min(cap, hard);
    cap = std::min(cap, config.max_batch > 0 ? config.max_batch : 64);
    // layouts ordered for the resident-tile kernel advance exactly as many vectors per pass as its row width holds (a
    // ragged last pass is padded with zero lanes): one geometry, one set of tile metadata
    int const res_lanes = std::max(1, res_row_bytes / dtype_size(dtype));
    if (natural.valid && natural.res_enabled && natural.res_failed_row_bytes != static_cast<uint32_t>(res_row_bytes) &&
        res_lanes <= cap && 2 * vectors >= res_lanes) return res_lanes;
    if (cap < 1) throw Error(PBK_RUNTIME_ERROR, "pbkpm: not enough device memory for one KPM vector pair");
    // a pass of more than one vector is padded to whole 16-byte chunks (lane_pad), so the batch itself must be a
    // multiple of the chunk width: buffers are sized for `rb` lanes and every launch uses lane_pad(lanes) <= rb
    int const vmax = 16 / dtype_size(dtype);
    cap = cap >= vmax ? cap / vmax * vmax : 1;
    int const nb = (vectors + cap - 1) / cap;
    int rb = (vectors + nb - 1) / nb;
    rb = std::min(lane_pad(rb), cap);
    return std::max(rb, 1);
}

void Engine::ensure_moment_buffers(int R, int M) {
    mom.ensure(sizeof(double) * 2 * static_cast<size_t>(R) * M + 64);
    m01.ensure(sizeof(double) * 3 * R + 64);
    acc.ensure(sizeof(double) * 2 * M + 64);
    partials.ensure(sizeof(double) * 3 * static_cast<size_t>(R) * max_step_blocks(num_sms));
    scratch.ensure(sizeof(double) * 2 * num_sms * 8);
}

void Engine::step(DeviceHamiltonian const& h, const void* x, void* y, void* y2, int64_t nrows, int R, bool subtract, bool sums,
                  double scale, int M, int nstep, int fin, int64_t y_block_stride, int64_t y2_block_stride) {
    StepArgs a;
    a.h = h.ell; a.x = x; a.y = y; a.y2 = y2; a.nrows = nrows; a.R = R; a.subtract = subtract; a.sums = sums; a.scale = scale;
    a.y_block_stride = y_block_stride; a.y2_block_stride = y2_block_stride;
    a.partials = partials.as<double>(); a.counter = counter.as<unsigned>(); a.mom = mom.as<double>(); a.m01 = m01.as<double>();
    a.M = M; a.n = nstep; a.fin = fin;
    a.tile = h.tile; a.tpb = step_tpb; a.blocks_per_sm = step_blocks_per_sm; a.prefetch = step_prefetch; a.prefetch_mask = step_prefetch_mask;
    a.packed = h.packed.bytes() ? h.packed.as() : nullptr; a.bulk_stages = bulk_stages; a.bulk_xstage = bulk_xstage; a.bulk_release = bulk_release;
    LaunchInfo info;
    bool done = false;
    if (h.res_enabled && subtract && sums && !y2 && nrows == n && &h == &natural && ensure_res_meta(natural, R)) {
        ResArgs r;
        r.tiles = h.res_tiles.as<ResTile>(); r.ntiles = h.res_ntiles; r.halo_rows = h.res_halo.as<int32_t>();
        r.codes = h.res_codes.as(); r.vals = h.res_vals.as(); r.geo = h.res_geo;
        r.x = x; r.y = y; r.nrows = nrows; r.R = R; r.k = h.ell.k;
        r.partials = a.partials; r.counter = a.counter; r.mom = a.mom; r.m01 = a.m01; r.M = M; r.n = nstep; r.fin = fin;
        PBK_CUDA(launch_step_res(dtype, r, num_sms, stream, &info, &done));
    }
    if (!done) PBK_CUDA(launch_step(dtype, a, num_sms, stream, &info));
    ++launches;
    ++stats.step_launches;
    if (info.res) ++stats.res_launches;
    else if (info.bulk) ++stats.bulk_launches;
    int const s = dtype_size(dtype);
    stats.step_bytes += static_cast<double>(nrows) * (h.ell.k * (s + 4.0) + static_cast<double>(R) * s * (2 + (subtract ? 1 : 0) + (y2 ? 1 : 0)));
}

void Engine::run_diagonal(DeviceHamiltonian const& h, int R, int M, bool opt_size) {
    void* r0 = vec_a.as();
    void* r1 = vec_b.as();
    PBK_CUDA(cudaEventRecord(ev2, stream));
    int64_t const nv = h.vec_rows > 0 ? h.vec_rows : n;
    auto emit = [&]() {
        // r1 = 0.5 * H2 * r0, m0 = 0.5 |r0|^2, m1 = <r1|r0>     (make_r1 + collect.initial)
        int64_t init_rows = nv;
        if (opt_size) {
            PBK_CUDA(cudaMemsetAsync(r1, 0, static_cast<size_t>(nv) * R * dtype_size(dtype), stream));
            init_rows = h.map.data[std::min(h.map.last_index(), h.map.src_offset + 1)];
        }
        step(h, r0, r1, nullptr, init_rows, R, false, true, 0.5, M, 0, FIN_INIT);
        for (int k = 2; k <= M / 2; ++k) {  // calc_moments::basic (diagonal), calc_moments.hpp:36-51
            int64_t const rows = opt_size ? h.map.optimal_size(k, M) : nv;
            step(h, r1, r0, nullptr, rows, R, true, true, 1.0, M, k, FIN_STEP);
            std::swap(r0, r1);
        }
    };
    // Small systems are launch-bound (a step is a few microseconds of work): capture the whole sequence once as a
    // CUDA graph and replay it on later runs with the same buffers and row counts.
    // One vector of a system small enough for one resident grid: the whole recursion in ONE persistent launch (grid
    // barrier per step instead of a kernel launch per step), kernels_persist.cu
    if (persist_mode && R == 1 && !opt_size && !h.transient && nv == n && M / 2 >= 2) {
        int const table_ctas = 2 * num_sms;
        persist_table.ensure(sizeof(double) * 3 * static_cast<size_t>(M / 2) * table_ctas);
        persist_barrier.ensure(64);
        PersistArgs p;
        p.h = h.ell; p.nrows = nv; p.buf0 = r0; p.buf1 = r1; p.steps = M / 2;
        p.table = persist_table.as<double>(); p.table_ctas = table_ctas; p.barrier = persist_barrier.as<unsigned>();
        p.mom = mom.as<double>(); p.M = M;
        bool handled = false;
        PBK_CUDA(launch_persistent_diagonal(dtype, p, num_sms, stream, &handled));
        if (handled) {
            launches += 2;
            ++stats.persist_launches;
            stats.step_launches += M / 2;
            int const s = dtype_size(dtype);
            stats.step_bytes += static_cast<double>(M / 2) * static_cast<double>(nv) * (h.ell.k * (s + 4.0) + 3.0 * s) - static_cast<double>(nv) * s;
            PBK_CUDA(cudaEventRecord(ev3, stream));
            PBK_CUDA(cudaEventSynchronize(ev3));
            float ms = 0;
            PBK_CUDA(cudaEventElapsedTime(&ms, ev2, ev3));
            stats.step_ms += ms;
            return;
        }
    }
    if (h.res_enabled && &h == &natural && !opt_size) ensure_res_meta(natural, R);   // before any capture: it synchronises
    bool const graphable = graph_mode && !h.transient && !h.res_enabled && M / 2 >= 8 &&
                           static_cast<double>(nv) * R * dtype_size(dtype) <= graph_max_bytes;
    bool done = false;
    if (graphable) {
        std::vector<int64_t> key = {reinterpret_cast<int64_t>(h.ell.val), reinterpret_cast<int64_t>(h.ell.col), reinterpret_cast<int64_t>(h.packed.as()),
                                    reinterpret_cast<int64_t>(r0), reinterpret_cast<int64_t>(r1), reinterpret_cast<int64_t>(mom.as()),
                                    reinterpret_cast<int64_t>(partials.as()), reinterpret_cast<int64_t>(m01.as()), reinterpret_cast<int64_t>(counter.as()),
                                    h.ell.pitch, h.ell.k, h.tile, R, M, opt_size ? 1 : 0, nv, dtype, step_tpb, step_blocks_per_sm, step_prefetch,
                                    bulk_stages, bulk_xstage ? 1 : 0, h.res_enabled ? 1 : 0, reinterpret_cast<int64_t>(h.res_tiles.as())};
        if (opt_size) for (int k = 1; k <= M / 2; ++k) key.push_back(h.map.optimal_size(k, M));
        auto it = graph_cache.find(key);
        if (it == graph_cache.end()) {
            if (graph_cache.size() >= 16) clear_graphs();
            RecursionGraph g;
            int64_t const l0 = launches, s0 = stats.step_launches, b0 = stats.bulk_launches;
            double const by0 = stats.step_bytes;
            void* const k0 = r0; void* const k1 = r1;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                bool ok = true;
                try { emit(); } catch (Error const&) { ok = false; }
                cudaError_t const end = cudaStreamEndCapture(stream, &graph);
                if (ok && end == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                    g.launches = launches - l0; g.step_launches = stats.step_launches - s0; g.bulk_launches = stats.bulk_launches - b0;
                    g.step_bytes = stats.step_bytes - by0;
                    it = graph_cache.emplace(std::move(key), g).first;
                }
                if (graph) cudaGraphDestroy(graph);
            }
            // the capture only recorded the launches: undo its bookkeeping, the replay below (or the direct run) redoes it
            launches = l0; stats.step_launches = s0; stats.bulk_launches = b0; stats.step_bytes = by0;
            r0 = k0; r1 = k1;
            if (it == graph_cache.end()) { cudaGetLastError(); graph_mode = 0; }   // capture unsupported here: plain launches from now on
        }
        if (it != graph_cache.end()) {
            PBK_CUDA(cudaGraphLaunch(it->second.exec, stream));
            launches += it->second.launches; stats.step_launches += it->second.step_launches; stats.bulk_launches += it->second.bulk_launches;
            stats.step_bytes += it->second.step_bytes;
            ++stats.graph_launches;
            done = true;
        }
    }
    if (!done) emit();
    PBK_CUDA(cudaEventRecord(ev3, stream));
    PBK_CUDA(cudaEventSynchronize(ev3));
    float ms = 0;
    PBK_CUDA(cudaEventElapsedTime(&ms, ev2, ev3));
    stats.step_ms += ms;
}

void Engine::run_offdiagonal(DeviceHamiltonian const& h, int M, bool opt_size, std::function<void(int, void*, double)> const& collect) {
    void* r0 = vec_a.as();
    void* r1 = vec_b.as();
    PBK_CUDA(cudaEventRecord(ev2, stream));
    int64_t const nv = h.vec_rows > 0 ? h.vec_rows : n;   // light-cone sub-systems are shorter than the system
    int64_t init_rows = nv;
    if (opt_size) {
        PBK_CUDA(cudaMemsetAsync(r1, 0, static_cast<size_t>(nv) * dtype_size(dtype), stream));
        init_rows = h.map.data[std::min(h.map.last_index(), h.map.src_offset + 1)];
    }
    step(h, r0, r1, nullptr, init_rows, 1, false, false, 0.5, M, 0, FIN_NONE);
    collect(0, r0, 0.5);
    collect(1, r1, 1.0);
    for (int k = 2; k < M; ++k) {  // calc_moments::basic (off-diagonal), calc_moments.hpp:103-114
        int64_t const rows = opt_size ? h.map.optimal_size(k, M) : nv;
        step(h, r1, r0, nullptr, rows, 1, true, false, 1.0, M, k, FIN_NONE);
        std::swap(r0, r1);
        collect(k, r1, 1.0);
    }
    PBK_CUDA(cudaEventRecord(ev3, stream));
    PBK_CUDA(cudaEventSynchronize(ev3));
    float ms = 0;
    PBK_CUDA(cudaEventElapsedTime(&ms, ev2, ev3));
    stats.step_ms += ms;
}

void Engine::reset_stats(int M, DeviceHamiltonian const& h, bool opt_size, double multiplier) {  // Stats.cpp:31-47
    if (!keep_kubo_buffers) { kubo_l.release(); kubo_r.release(); kubo_ws.release(); }   // a different quantity: the stacks' memory is free again
    int64_t const h2d = stats.h2d_bytes;
    stats = pbk_stats{};
    stats.h2d_bytes = h2d;
    stats.num_moments = M;
    stats.uses_full_system = h.map.uses_full_system(M);
    auto count = [&](bool opt, uint64_t per_row) {
        uint64_t result = 0;
        if (!opt) result = static_cast<uint64_t>(M) * n * per_row;
        else for (int k = 0; k < M; ++k) result += static_cast<uint64_t>(h.map.optimal_size(k, M)) * per_row;
        if (h.idx.is_diagonal()) result /= 2;
        return result;
    };
    stats.nnz = count(false, h.ell.k);
    stats.opt_nnz = count(opt_size, h.ell.k);
    stats.vec = count(false, 1);
    stats.opt_vec = count(opt_size, 1);
    stats.multiplier = multiplier;
    stats.matrix_memory = static_cast<uint64_t>(n) * h.ell.k * (dtype_size(dtype) + 4);
    stats.vector_memory = static_cast<uint64_t>(n) * dtype_size(dtype);
    stats.hamiltonian_time = h.seconds;
    launches = 0;
}

void Engine::begin_moments() {
    PBK_CUDA(cudaSetDevice(device));
    moments_wall0 = now_seconds();
    PBK_CUDA(cudaEventRecord(ev_begin, stream));
}

void Engine::end_moments() {
    PBK_CUDA(cudaEventRecord(ev_end, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
    float device_ms = 0;
    PBK_CUDA(cudaEventElapsedTime(&device_ms, ev_begin, ev_end));
    stats.moments_device_ms += device_ms;
    stats.moments_time += now_seconds() - moments_wall0;
    stats.kernel_launches = launches;
    stats.eps = stats.moments_time > 0 ? stats.multiplier * static_cast<double>(stats.opt_nnz) / stats.moments_time : 0;
}

void shard_range(int total, int world, int rank, int* first, int* count) {
    // contiguous blocks, remainder to the lowest ranks
    int const base = total / world, rem = total % world;
    *count = base + (rank < rem ? 1 : 0);
    *first = rank * base + std::min(rank, rem);
}

void Engine::shard(int total, int* first, int* count) const { shard_range(total, world, rank, first, count); }

void Engine::seed_stream(int64_t skip_vectors) {
    // position in the reference's single default-seeded stream: vector j owns draws [j*N*w, (j+1)*N*w)
    stream_pos = static_cast<uint64_t>(skip_vectors) * static_cast<uint64_t>(n) * dtype_words(dtype);
    if (mt_sequential) {
        PBK_CUDA(launch_mt_seed(mt_state.as<uint32_t>(), stream));
        ++launches;
        if (skip_vectors > 0) {
            PBK_CUDA(launch_mt_generate(mt_state.as<uint32_t>(), nullptr, skip_vectors * n * dtype_words(dtype), stream));
            ++launches;
        }
    }
}

void Engine::generate_random_block(DeviceHamiltonian const& h, int lanes, int R, void* dst) {
    int64_t const words = static_cast<int64_t>(lanes) * n * dtype_words(dtype);
    raw.ensure(sizeof(uint32_t) * words);
    PBK_CUDA(cudaEventRecord(ev0, stream));
    if (mt_sequential) {
        PBK_CUDA(launch_mt_generate(mt_state.as<uint32_t>(), raw.as<uint32_t>(), words, stream));
        ++launches;
    } else {
        int nl = 0;
        mt_states.ensure(sizeof(uint32_t) * mt_stream_scratch_words(MT_MAX_SEGMENTS));
        PBK_CUDA(launch_mt_stream(mt_states.as<uint32_t>(), MT_MAX_SEGMENTS, stream_pos, words, raw.as<uint32_t>(), stream, &nl));
        launches += nl;
    }
    stream_pos += static_cast<uint64_t>(words);
    PBK_CUDA(launch_random_transform(dtype, raw.as<uint32_t>(), n, R, lanes, h.reordered ? h.perm.as<int32_t>() : nullptr, dst, stream));
    PBK_CUDA(cudaEventRecord(ev1, stream));
    launches += 1;
    PBK_CUDA(cudaEventSynchronize(ev1));
    float ms = 0;
    PBK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    stats.starter_ms += ms;
}

void Engine::allreduce(double* dev, int64_t count) {
    if (world <= 1 || !comm) return;
    nccl->check(nccl->AllReduce(dev, dev, static_cast<size_t>(count), /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, stream), "ncclAllReduce");
}

void Engine::comm_init(int world_, int rank_, const char* id) {
    PBK_CUDA(cudaSetDevice(device));
    comm_destroy();
    if (world_ <= 1) { world = 1; rank = 0; return; }
    if (!nccl) nccl = std::make_unique<NcclApi>();
    NcclId uid;
    std::memcpy(uid.bytes, id, sizeof(uid.bytes));
    using InitFn = int (*)(void**, int, NcclId, int);
    nccl->check(reinterpret_cast<InitFn>(nccl->init_rank)(&comm, world_, uid, rank_), "ncclCommInitRank");
    world = world_;
    rank = rank_;
}

void Engine::comm_destroy() {
    if (comm && nccl) { nccl->CommDestroy(comm); }
    comm = nullptr;
    world = 1;
    rank = 0;
}

// ------------------------------------------------------------------------------------------------
// compute-strategy level entry points
// ------------------------------------------------------------------------------------------------
static void check_num_moments(int M) {
    if (M < 2 || M % 2 != 0) throw Error(PBK_INVALID_ARGUMENT, "pbkpm: num_moments must be an even number >= 2 (the reference uses 4k+2)");
}

void Engine::moments_dos(int M, int num_random, cd* out) {
    check_num_moments(M);
    if (num_random < 1) throw Error(PBK_INVALID_ARGUMENT, "num_random must be positive");
    auto& h = natural_hamiltonian();
    h.idx = Indices{{0}, {0}};
    reset_stats(M, h, false, num_random);
    int first = 0, count = 0;
    shard(num_random, &first, &count);
    begin_moments();
    progress(-1, num_random);
    ensure_moment_buffers(1, M);
    PBK_CUDA(cudaMemsetAsync(acc.as(), 0, sizeof(double) * 2 * M, stream));
    if (count > 0) {
        int const rb = pick_batch(count, 0);
        ensure_moment_buffers(rb, M);
        size_t const block_bytes = static_cast<size_t>(n) * rb * dtype_size(dtype);
        double const t_alloc = now_seconds();
        vec_a.ensure(block_bytes);
        vec_b.ensure(block_bytes);
        if (std::getenv("PBK_TIMING")) std::fprintf(stderr, "[pbkpm] moments_dos: vector blocks (2 x %.1f GB) ready after %.3f s\n", block_bytes / 1e9, now_seconds() - t_alloc);
        stats.batch = rb;
        seed_stream(first);
        bool const full_width = h.res_enabled && rb * dtype_size(dtype) == res_row_bytes;   // resident-tile passes keep one row width
        for (int b0 = 0; b0 < count; b0 += rb) {
            int const lanes = std::min(rb, count - b0);
            int const R = full_width ? rb : lane_pad(lanes);
            generate_random_block(h, lanes, R, vec_a.as());
            run_diagonal(h, R, M, false);
            PBK_CUDA(launch_accumulate_lanes(mom.as<double>(), lanes, M, acc.as<double>(), stream));
            ++launches;
            ++stats.num_batches;
            progress(lanes, num_random);
        }
    }
    allreduce(acc.as<double>(), 2 * M);
    PBK_CUDA(cudaMemcpyAsync(out, acc.as(), sizeof(double) * 2 * M, cudaMemcpyDeviceToHost, stream));
    stats.d2h_bytes += sizeof(double) * 2 * M;
    end_moments();
    if (num_random != 1) for (int i = 0; i < M; ++i) out[i] /= static_cast<double>(num_random);  // Moments.cpp:23-27
    progress(num_random, num_random);
}

void Engine::moments_diagonal(int M, const cd* r0, int count, cd* out) {
    check_num_moments(M);
    auto& h = natural_hamiltonian();
    reset_stats(M, h, false, count);
    begin_moments();
    int const rb = pick_batch(count, 0);
    ensure_moment_buffers(rb, M);
    size_t const block_bytes = static_cast<size_t>(n) * rb * dtype_size(dtype);
    vec_a.ensure(block_bytes);
    vec_b.ensure(block_bytes);
    DevBuf staging(sizeof(double) * 2 * n);
    std::vector<cd> host(static_cast<size_t>(rb) * M);
    stats.batch = rb;
    for (int b0 = 0; b0 < count; b0 += rb) {
        int const lanes = std::min(rb, count - b0);
        int const R = lane_pad(lanes);
        PBK_CUDA(cudaMemsetAsync(vec_a.as(), 0, static_cast<size_t>(n) * R * dtype_size(dtype), stream));
        for (int j = 0; j < lanes; ++j) {
            PBK_CUDA(cudaMemcpyAsync(staging.as(), r0 + static_cast<size_t>(b0 + j) * n, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, stream));
            PBK_CUDA(launch_scatter_block(dtype, staging.as<double>(), n, R, j, h.reordered ? h.perm.as<int32_t>() : nullptr, vec_a.as(), stream));
            PBK_CUDA(cudaStreamSynchronize(stream));
            stats.h2d_bytes += sizeof(double) * 2 * n;
        }
        run_diagonal(h, R, M, false);
        PBK_CUDA(cudaMemcpyAsync(host.data(), mom.as(), sizeof(cd) * static_cast<size_t>(lanes) * M, cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
        for (int j = 0; j < lanes; ++j) for (int k = 0; k < M; ++k) out[static_cast<size_t>(k) * count + b0 + j] = host[static_cast<size_t>(j) * M + k];
        ++stats.num_batches;
    }
    end_moments();
}

void Engine::random_vectors(int count, cd* out) {
    auto& h = natural_hamiltonian();
    PBK_CUDA(cudaSetDevice(device));
    size_t const block_bytes = static_cast<size_t>(n) * dtype_size(dtype);
    vec_a.ensure(block_bytes);
    DevBuf staging(sizeof(double) * 2 * n);
    seed_stream(0);
    for (int j = 0; j < count; ++j) {
        generate_random_block(h, 1, 1, vec_a.as());
        PBK_CUDA(launch_extract_lane(dtype, vec_a.as(), n, 1, 0, h.reordered ? h.perm.as<int32_t>() : nullptr, staging.as<double>(), stream));
        PBK_CUDA(cudaMemcpyAsync(out + static_cast<size_t>(j) * n, staging.as(), sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
    }
}

// ------------------------------------------------------------------------------------------------
// LDOS on light-cone sub-systems.  A unit vector at site i spreads one breadth-first shell per Chebyshev step, and the
// diagonal algorithm stops at step M/2, so <i|T_n(H~)|i> for n < M only involves the sites within M/2 bonds of i.  The
// reference relabels the *whole* system from i on the host for every site (OptimizedHamiltonian::create_reordered) and
// then touches only `optimal_size(n)` rows per step.  Here the host only walks the ball itself (truncated BFS, same
// visiting order), the ball's rows are cut out of the device-resident scaled ELL by a kernel, and the recursion runs on
// that sub-system with the same slice map -- identical arithmetic per row, no per-site pass over the full system.
// ------------------------------------------------------------------------------------------------
SliceMap Cone::map() const {
    SliceMap m;
    m.data = borders;
    if (!exhausted && m.data.size() > 1) m.data.pop_back();   // rows of the outermost shell miss neighbours: never processed
    m.src_offset = 0;
    m.dest_offset = 0;
    return m;
}

Cone light_cone(const int32_t* indptr, const int32_t* indices, int32_t src, int depth, std::vector<int32_t>& mark) {
    Cone c;
    c.queue.push_back(src);
    mark[src] = 0;
    c.borders = {1};
    size_t head = 0;
    for (int shell = 0; shell < depth; ++shell) {
        size_t const end = c.queue.size();
        for (; head < end; ++head) {
            int32_t const row = c.queue[head];
            for (int p = indptr[row]; p < indptr[row + 1]; ++p) {
                int32_t const col = indices[p];
                if (mark[col] < 0) { mark[col] = static_cast<int32_t>(c.queue.size()); c.queue.push_back(col); }
            }
        }
        if (c.queue.size() == end) { c.exhausted = true; break; }
        c.borders.push_back(static_cast<int32_t>(c.queue.size()));
    }
    for (int32_t site : c.queue) mark[site] = -1;
    return c;
}

Cone Engine::bfs_cone(int32_t src, int depth, std::vector<int32_t>& mark) const {
    return light_cone(h_indptr.data(), h_indices.data(), src, depth, mark);
}

bool Engine::moments_ldos_cones(int M, Indices const& target, cd* out) {
    int const nidx = static_cast<int>(target.src.size());
    int const depth = M / 2;
    auto& hn = natural_hamiltonian();
    int const s = dtype_size(dtype);
    int const kell = hn.ell.k;
    auto half_rows = [&](SliceMap const& m) {   // rows processed by one diagonal calculation (Stats.cpp:31-47: halved)
        double sum = 0;
        for (int k = 0; k < M; ++k) sum += static_cast<double>(m.optimal_size(k, M));
        return sum / 2;
    };
    // cost model on the first site (every rank looks at the same one): per-site cones against full-system batches
    std::vector<std::vector<int32_t>> marks(1, std::vector<int32_t>(static_cast<size_t>(n), -1));
    Cone first_cone = bfs_cone(target.src[0], depth, marks[0]);
    if (first_cone.exhausted && static_cast<int64_t>(first_cone.queue.size()) < n) {
        throw Error(PBK_RUNTIME_ERROR, "KPM: the Hamiltonian graph is not connected; the optimal_size "
                                       "reordering needs a connected system");
    }
    if (nidx > 1) {
        // one launch per step either way: a launch costs at least the launch latency, else its bytes at the rate the
        // kernel reaches (scalar general kernel on a cone / staged kernel on the full system)
        double const t_launch = 5e-6, bw_cone = 3e12, bw_full = 5e12;
        int const steps = M / 2;
        double const cone_step_bytes = half_rows(first_cone.map()) / steps * (kell * (s + 4.0) + 3.0 * s);
        int const grp = std::min(nidx, 32);   // sub-systems advanced by one launch (bounded by memory later on)
        double const cone_time = std::ceil(static_cast<double>(nidx) / grp) * steps * std::max(t_launch, grp * cone_step_bytes / bw_cone);
        int const rb = std::min(nidx, 64);
        double const full_step_bytes = static_cast<double>(n) * (kell * (s + 4.0) + 3.0 * rb * s);
        double const full_time = std::ceil(nidx / 64.0) * steps * std::max(t_launch, full_step_bytes / bw_full);
        if (cone_time >= full_time) return false;
    }

    int64_t const h2d0 = stats.h2d_bytes;
    stats = pbk_stats{};
    stats.h2d_bytes = h2d0;
    stats.num_moments = M;
    stats.multiplier = nidx;
    stats.nnz = static_cast<uint64_t>(M) * n * kell / 2;
    stats.vec = static_cast<uint64_t>(M) * n / 2;
    stats.matrix_memory = static_cast<uint64_t>(n) * kell * (s + 4);
    stats.vector_memory = static_cast<uint64_t>(n) * s;
    stats.hamiltonian_time = hn.seconds;
    stats.batch = 1;
    launches = 0;

    begin_moments();
    progress(-1, nidx);
    int first = 0, count = 0;
    shard(nidx, &first, &count);
    std::vector<cd> table(static_cast<size_t>(M) * nidx, cd(0, 0));
    if (cone_gmap_rows != n) {
        cone_gmap.ensure(sizeof(int32_t) * n);
        PBK_CUDA(cudaMemsetAsync(cone_gmap.as(), 0xff, sizeof(int32_t) * n, stream));
        cone_gmap_rows = n;
    }
    ensure_moment_buffers(1, M);
    cone_table.ensure(sizeof(cd) * static_cast<size_t>(M) * std::max(count, 1));
    const int32_t* perm = hn.reordered ? hn.perm.as<int32_t>() : nullptr;
    double opt_rows = 0;
    bool full = false;
    int const steps = M / 2;
    constexpr int C_MAX = 3;
    int const blocks_cap = num_sms * 4;   // blocks per sub-system and launch

    // group size: sub-systems advanced together by one launch per step, bounded by device memory
    auto slot_bytes = [&](Cone const& c) {
        int64_t const rows = c.complete_rows();
        int64_t const pitch = (rows + 31) / 32 * 32;
        return static_cast<double>(kell) * pitch * (s + 4.0) + 2.0 * static_cast<double>(c.queue.size()) * s + 4.0 * c.queue.size();
    };
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += DevBuf::cached_bytes();   // blocks parked in the allocation cache are given back on demand
    double const budget = std::min(0.4 * static_cast<double>(free_b) + static_cast<double>(cone_val.bytes() + cone_col.bytes() + vec_a.bytes()), 16e9);
    int group = static_cast<int>(std::max(1.0, std::min(64.0, budget / (1.25 * slot_bytes(first_cone)))));
    if (cone_group_cap > 0) group = std::min(group, cone_group_cap);
    group = std::max(1, std::min(group, count));
    int const nthreads = std::max(1, std::min<int>({static_cast<int>(std::thread::hardware_concurrency()), 16, group}));
    marks.resize(nthreads, std::vector<int32_t>(static_cast<size_t>(n), -1));
    stats.batch = group;
    DevBuf slots_dev(sizeof(ConeSlot) * group), rows_dev(sizeof(int32_t) * static_cast<size_t>(group) * (steps + 1)),
           m01_dev(sizeof(double) * 3 * group), counters_dev(sizeof(unsigned) * group),
           partials_dev(sizeof(double) * C_MAX * static_cast<size_t>(group) * blocks_cap);
    PBK_CUDA(cudaMemsetAsync(counters_dev.as(), 0, sizeof(unsigned) * group, stream));

    for (int c0 = 0; c0 < count; c0 += group) {
        int const nc = std::min(group, count - c0);
        std::vector<Cone> cones(nc);
        {   // host: the balls of this group, one thread per site (overlaps the device work of the previous group)
            int const nt = std::min(nthreads, nc);
            run_pool(nt, [&](int t) {
                for (int j = t; j < nc; j += nt) {
                    if (first + c0 + j == 0) cones[j] = first_cone;
                    else cones[j] = bfs_cone(target.src[first + c0 + j], depth, marks[t]);
                }
            });
        }
        // pooled buffers of the group
        std::vector<int64_t> ell_off(nc + 1, 0), vec_off(nc + 1, 0);
        std::vector<int64_t> pitches(nc), rows_all(nc);
        int64_t max_queue = 0;
        for (int j = 0; j < nc; ++j) {
            rows_all[j] = cones[j].complete_rows();
            pitches[j] = (rows_all[j] + 31) / 32 * 32;
            ell_off[j + 1] = ell_off[j] + static_cast<int64_t>(kell) * pitches[j];
            int64_t const nloc = static_cast<int64_t>(cones[j].queue.size());
            vec_off[j + 1] = vec_off[j] + 2 * ((nloc + 31) / 32 * 32);
            max_queue = std::max(max_queue, nloc);
        }
        // 25 % headroom: later groups (other ball sizes) should not force a re-allocation, which would wait for the device
        auto ensure_room = [](DevBuf& b, size_t bytes) { if (bytes > b.bytes()) b.ensure(bytes + bytes / 4); };
        ensure_room(cone_val, static_cast<size_t>(ell_off[nc]) * s);
        ensure_room(cone_col, static_cast<size_t>(ell_off[nc]) * sizeof(int32_t));
        ensure_room(vec_a, static_cast<size_t>(vec_off[nc]) * s);
        ensure_room(cone_queue, sizeof(int32_t) * static_cast<size_t>(max_queue));
        std::vector<ConeSlot> slots(nc);
        std::vector<int32_t> rows_table(static_cast<size_t>(nc) * (steps + 1), 0);
        std::vector<int64_t> max_rows(steps + 1, 0);
        int64_t max_nvec = 0;
        for (int j = 0; j < nc; ++j) {
            Cone const& cone = cones[j];
            int64_t const nloc = static_cast<int64_t>(cone.queue.size());
            char* const val_j = cone_val.as<char>() + static_cast<size_t>(ell_off[j]) * s;
            int32_t* const col_j = cone_col.as<int32_t>() + ell_off[j];
            PBK_CUDA(cudaMemcpyAsync(cone_queue.as(), cone.queue.data(), sizeof(int32_t) * nloc, cudaMemcpyHostToDevice, stream));
            stats.h2d_bytes += static_cast<int64_t>(sizeof(int32_t) * nloc);
            PBK_CUDA(launch_cone_mark(cone_queue.as<int32_t>(), nloc, perm, cone_gmap.as<int32_t>(), true, stream));
            PBK_CUDA(launch_cone_extract(dtype, hn.ell, cone_queue.as<int32_t>(), perm, cone_gmap.as<int32_t>(), rows_all[j], val_j, col_j, pitches[j], stream));
            PBK_CUDA(launch_cone_mark(cone_queue.as<int32_t>(), nloc, perm, cone_gmap.as<int32_t>(), false, stream));
            launches += 3;
            SliceMap const map = cone.map();
            int32_t* rt = rows_table.data() + static_cast<size_t>(j) * (steps + 1);
            rt[1] = map.data[std::min(map.last_index(), 1)];
            for (int k = 2; k <= steps; ++k) rt[k] = static_cast<int32_t>(map.optimal_size(k, M));
            for (int k = 1; k <= steps; ++k) max_rows[k] = std::max<int64_t>(max_rows[k], rt[k]);
            max_nvec = std::max(max_nvec, nloc);
            ConeSlot& sl = slots[j];
            sl.val = val_j; sl.col = col_j; sl.pitch = pitches[j];
            sl.buf[0] = vec_a.as<char>() + static_cast<size_t>(vec_off[j]) * s;
            sl.buf[1] = vec_a.as<char>() + static_cast<size_t>(vec_off[j] + (nloc + 31) / 32 * 32) * s;
            sl.nvec = nloc;
            sl.rows = rows_dev.as<int32_t>() + static_cast<size_t>(j) * (steps + 1);
            sl.mom = cone_table.as<double>() + 2 * static_cast<size_t>(c0 + j) * M;
            sl.m01 = m01_dev.as<double>() + 3 * j;
            opt_rows += half_rows(map);
            full = full || map.uses_full_system(M);
        }
        PBK_CUDA(cudaMemcpyAsync(slots_dev.as(), slots.data(), sizeof(ConeSlot) * nc, cudaMemcpyHostToDevice, stream));
        PBK_CUDA(cudaMemcpyAsync(rows_dev.as(), rows_table.data(), sizeof(int32_t) * rows_table.size(), cudaMemcpyHostToDevice, stream));
        if (c0 == 0) PBK_CUDA(cudaEventRecord(ev2, stream));
        PBK_CUDA(launch_cone_group_start(dtype, slots_dev.as<ConeSlot>(), nc, max_nvec, stream));
        ++launches;
        for (int k = 1; k <= steps; ++k) {
            PBK_CUDA(launch_cone_group_step(dtype, slots_dev.as<ConeSlot>(), nc, k, kell, M, max_rows[k], partials_dev.as<double>(),
                                            counters_dev.as<unsigned>(), blocks_cap, stream));
            ++launches;
            ++stats.step_launches;
            for (int j = 0; j < nc; ++j) {
                int64_t const r = rows_table[static_cast<size_t>(j) * (steps + 1) + k];
                stats.step_bytes += static_cast<double>(r) * (kell * (s + 4.0) + (k == 1 ? 2.0 : 3.0) * s);
            }
        }
        // no synchronisation here: the host walks the balls of the next group while the device advances this one
        // (the pageable uploads above were staged when they were issued)
        stats.num_batches += 1;
        progress(nc, nidx);
    }
    if (count > 0) {
        PBK_CUDA(cudaEventRecord(ev3, stream));
        PBK_CUDA(cudaEventSynchronize(ev3));
        float ms = 0;
        PBK_CUDA(cudaEventElapsedTime(&ms, ev2, ev3));
        stats.step_ms += ms;   // first step of the first group .. last step of the last group (extraction of later groups included)
    }
    if (count > 0) {
        std::vector<cd> host(static_cast<size_t>(count) * M);
        PBK_CUDA(cudaMemcpyAsync(host.data(), cone_table.as(), sizeof(cd) * host.size(), cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
        stats.d2h_bytes += static_cast<int64_t>(sizeof(cd) * host.size());
        for (int j = 0; j < count; ++j) for (int k = 0; k < M; ++k) table[static_cast<size_t>(k) * nidx + first + j] = host[static_cast<size_t>(j) * M + k];
    }
    if (world > 1) {
        DevBuf t(sizeof(cd) * table.size());
        PBK_CUDA(cudaMemcpyAsync(t.as(), table.data(), sizeof(cd) * table.size(), cudaMemcpyHostToDevice, stream));
        allreduce(t.as<double>(), static_cast<int64_t>(2 * table.size()));
        PBK_CUDA(cudaMemcpyAsync(table.data(), t.as(), sizeof(cd) * table.size(), cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
    }
    std::copy(table.begin(), table.end(), out);
    stats.opt_vec = static_cast<uint64_t>(opt_rows / std::max(count, 1));
    stats.opt_nnz = stats.opt_vec * kell;
    stats.uses_full_system = full;
    end_moments();
    progress(nidx, nidx);
    return true;
}

void Engine::moments_ldos(int M, const int32_t* idx, int nidx, cd* out) {
    check_num_moments(M);
    if (nidx < 1) throw Error(PBK_INVALID_ARGUMENT, "at least one index is required");
    for (int i = 0; i < nidx; ++i) if (idx[i] < 0 || idx[i] >= n) throw Error(PBK_INVALID_ARGUMENT, "LDOS index out of range");
    Indices target{std::vector<int32_t>(idx, idx + nidx), std::vector<int32_t>(idx, idx + nidx)};
    // per-site light-cone sub-systems cut out of the resident Hamiltonian, unless many small launches would cost more
    // than advancing all the unit vectors together (then: the reference's relabelling from src[0], or the full-system batch)
    if (config.optimal_size && cone_mode && moments_ldos_cones(M, target, out)) return;
    // Light-cone slicing pays when the sources sit together (one site, the orbitals of a site, a small region).
    // For sources spread over the sample the union of the light cones is the whole system from the first steps on:
    // then the full-system locality layout with the staged kernel is the faster way to advance the unit vectors.
    bool spread = false;
    if (config.optimal_size && nidx > 1 && !(optimized.valid && optimized.original_idx == target)) {
        if (!bfs_ready.valid_for(target)) bfs_ready = bfs_order(target);
        double sliced_rows = 0;
        for (int k = 0; k < M; ++k) sliced_rows += static_cast<double>(bfs_ready.map.optimal_size(k, M));
        spread = sliced_rows > 0.6 * static_cast<double>(M) * static_cast<double>(n);
        if (spread) bfs_ready = BfsOrder();
    }
    DeviceHamiltonian* hp = nullptr;
    if (spread) {
        auto& hn = natural_hamiltonian();
        ensure_host_order(hn);
        hn.idx = Indices{};
        for (int32_t i : target.src) hn.idx.src.push_back(hn.reordered ? hn.reorder_map[i] : i);
        hn.idx.dest = hn.idx.src;
        hp = &hn;
    } else {
        hp = &optimized_for(target);
    }
    auto& h = *hp;
    bool const opt = config.optimal_size != 0 && h.sliced;
    reset_stats(M, h, opt, nidx);
    begin_moments();
    progress(-1, nidx);
    // sources are sharded over ranks in contiguous blocks; every rank returns the full table after the allreduce
    int first = 0, count = 0;
    shard(nidx, &first, &count);
    std::vector<cd> table(static_cast<size_t>(M) * nidx, cd(0, 0));
    if (count > 0) {
        int const rb = pick_batch(count, 0);
        ensure_moment_buffers(rb, M);
        size_t const block_bytes = static_cast<size_t>(n) * rb * dtype_size(dtype);
        vec_a.ensure(block_bytes);
        vec_b.ensure(block_bytes);
        idx_buf.ensure(sizeof(int32_t) * rb);
        std::vector<cd> host(static_cast<size_t>(rb) * M);
        stats.batch = rb;
        for (int b0 = 0; b0 < count; b0 += rb) {
            int const lanes = std::min(rb, count - b0);
            int const R = lane_pad(lanes);
            PBK_CUDA(cudaMemcpyAsync(idx_buf.as(), h.idx.src.data() + first + b0, sizeof(int32_t) * lanes, cudaMemcpyHostToDevice, stream));
            PBK_CUDA(launch_unit_starter(dtype, vec_a.as(), n, R, idx_buf.as<int32_t>(), lanes, stream));
            launches += 1;
            run_diagonal(h, R, M, opt);
            PBK_CUDA(cudaMemcpyAsync(host.data(), mom.as(), sizeof(cd) * static_cast<size_t>(lanes) * M, cudaMemcpyDeviceToHost, stream));
            PBK_CUDA(cudaStreamSynchronize(stream));
            stats.d2h_bytes += sizeof(cd) * static_cast<size_t>(lanes) * M;
            for (int j = 0; j < lanes; ++j) for (int k = 0; k < M; ++k) table[static_cast<size_t>(k) * nidx + first + b0 + j] = host[static_cast<size_t>(j) * M + k];
            ++stats.num_batches;
            progress(lanes, nidx);
        }
    }
    if (world > 1) {
        DevBuf t(sizeof(cd) * table.size());
        PBK_CUDA(cudaMemcpyAsync(t.as(), table.data(), sizeof(cd) * table.size(), cudaMemcpyHostToDevice, stream));
        allreduce(t.as<double>(), static_cast<int64_t>(2 * table.size()));
        PBK_CUDA(cudaMemcpyAsync(table.data(), t.as(), sizeof(cd) * table.size(), cudaMemcpyDeviceToHost, stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
    }
    std::copy(table.begin(), table.end(), out);
    end_moments();
    progress(nidx, nidx);
}

/// Light cone of an off-diagonal Green's function <dest_j| T_n(H~) |src>, n < M: the breadth-first ball around `src`
/// up to shell (M - 1 + dest_offset) / 2 + 1, where dest_offset is the shell of the farthest destination -- step n of the
/// recursion only needs the rows within min(n, M - 1 - n + dest_offset) bonds of the source (SliceMap::index with
/// src_offset = 0, OptimizedHamiltonian.hpp:67-77), and the ball has one more shell so that these rows are complete.
/// Returns false when the ball would be most of the system or a destination is out of reach (then: full relabelling).
static bool greens_cone(const int32_t* indptr, const int32_t* indices, int64_t n, int32_t src, std::vector<int32_t> const& dests, int M,
                        std::vector<int32_t>& mark, Cone& c, std::vector<int32_t>& dest_pos, int& dest_offset) {
    c = Cone();
    c.queue.push_back(src);
    mark[src] = 0;
    c.borders = {1};
    std::vector<int> dest_shell(dests.size(), -1);
    size_t found = 0;
    for (size_t j = 0; j < dests.size(); ++j) if (dests[j] == src) { dest_shell[j] = 0; ++found; }
    size_t head = 0;
    bool ok = true;
    for (int shell = 1; shell <= M; ++shell) {
        if (found == dests.size()) {
            int far = 0;
            for (int d : dest_shell) far = std::max(far, d);
            if (shell > (M - 1 + far) / 2 + 1) break;       // the previous shell was the last one needed
        }
        if (static_cast<int64_t>(c.queue.size()) > n / 2) { ok = false; break; }
        size_t const end = c.queue.size();
        for (; head < end; ++head) {
            int32_t const row = c.queue[head];
            for (int p = indptr[row]; p < indptr[row + 1]; ++p) {
                int32_t const col = indices[p];
                if (mark[col] < 0) { mark[col] = static_cast<int32_t>(c.queue.size()); c.queue.push_back(col); }
            }
        }
        if (c.queue.size() == end) { c.exhausted = true; break; }
        c.borders.push_back(static_cast<int32_t>(c.queue.size()));
        for (size_t j = 0; j < dests.size(); ++j) {
            if (dest_shell[j] < 0 && mark[dests[j]] >= 0) { dest_shell[j] = shell; ++found; }
        }
    }
    if (found != dests.size()) ok = false;                  // a destination outside the reach of M - 1 steps (or of the component)
    dest_pos.clear();
    dest_offset = 0;
    if (ok) {
        for (size_t j = 0; j < dests.size(); ++j) { dest_pos.push_back(mark[dests[j]]); dest_offset = std::max(dest_offset, dest_shell[j]); }
    }
    for (int32_t site : c.queue) mark[site] = -1;
    return ok;
}

void Engine::moments_greens(int M, int row, const int32_t* cols, int ncols, cd* out) {
    check_num_moments(M);
    if (ncols < 1) throw Error(PBK_INVALID_ARGUMENT, "at least one column index is required");
    require_hamiltonian();
    if (row < 0 || row >= n) throw Error(PBK_LOGIC_ERROR, "KPM::calc_greens(i,j): invalid value for i or j.");   // KPM.cpp:105-107
    for (int i = 0; i < ncols; ++i) if (cols[i] < 0 || cols[i] >= n) throw Error(PBK_LOGIC_ERROR, "KPM::calc_greens(i,j): invalid value for i or j.");
    Indices target{{row}, std::vector<int32_t>(cols, cols + ncols)};
    if (config.optimal_size && cone_mode && target.is_diagonal()) {   // <i|T_n|i>: the LDOS recursion on the light cone of i
        int32_t const site = row;
        moments_ldos(M, &site, 1, out);
        return;
    }
    // ---- off-diagonal on a light-cone sub-system cut out of the resident Hamiltonian (no relabelling of the full system) ----
    if (config.optimal_size && cone_mode && dev_build) {
        std::vector<int32_t> mark(static_cast<size_t>(n), -1), dest_pos;
        Cone cone;
        int dest_offset = 0;
        if (greens_cone(h_indptr.data(), h_indices.data(), n, row, target.dest, M, mark, cone, dest_pos, dest_offset)) {
            auto& hn = natural_hamiltonian();
            int const s = dtype_size(dtype);
            int const kell = hn.ell.k;
            int64_t const nloc = static_cast<int64_t>(cone.queue.size());
            int64_t const rows = cone.complete_rows();
            int64_t const pitch = (rows + 31) / 32 * 32;
            DeviceHamiltonian sub;
            sub.transient = true;
            sub.sliced = true;
            sub.vec_rows = nloc;
            sub.map = cone.map();
            sub.map.dest_offset = std::min(dest_offset, sub.map.last_index());
            sub.idx = Indices{{0}, dest_pos};
            sub.original_idx = target;
            sub.seconds = hn.seconds;
            sub.ell.k = kell;
            reset_stats(M, sub, true, 1);
            // Stats of the reference count the rows of the *system* for the unoptimised figure
            begin_moments();
            auto ensure_room = [](DevBuf& b, size_t bytes) { if (bytes > b.bytes()) b.ensure(bytes + bytes / 4); };
            ensure_room(cone_val, static_cast<size_t>(kell) * pitch * s);
            ensure_room(cone_col, static_cast<size_t>(kell) * pitch * sizeof(int32_t));
            ensure_room(cone_queue, sizeof(int32_t) * static_cast<size_t>(nloc));
            ensure_room(vec_a, static_cast<size_t>(nloc) * s);
            ensure_room(vec_b, static_cast<size_t>(nloc) * s);
            if (cone_gmap_rows != n) {
                cone_gmap.ensure(sizeof(int32_t) * n);
                PBK_CUDA(cudaMemsetAsync(cone_gmap.as(), 0xff, sizeof(int32_t) * n, stream));
                cone_gmap_rows = n;
            }
            const int32_t* perm = hn.reordered ? hn.perm.as<int32_t>() : nullptr;
            PBK_CUDA(cudaMemcpyAsync(cone_queue.as(), cone.queue.data(), sizeof(int32_t) * nloc, cudaMemcpyHostToDevice, stream));
            stats.h2d_bytes += static_cast<int64_t>(sizeof(int32_t) * nloc);
            PBK_CUDA(launch_cone_mark(cone_queue.as<int32_t>(), nloc, perm, cone_gmap.as<int32_t>(), true, stream));
            PBK_CUDA(launch_cone_extract(dtype, hn.ell, cone_queue.as<int32_t>(), perm, cone_gmap.as<int32_t>(), rows, cone_val.as(), cone_col.as<int32_t>(), pitch, stream));
            PBK_CUDA(launch_cone_mark(cone_queue.as<int32_t>(), nloc, perm, cone_gmap.as<int32_t>(), false, stream));
            launches += 3;
            sub.ell = EllDev{cone_val.as(), cone_col.as<int32_t>(), rows, pitch, kell};
            stats.batch = 1;
            stats.num_batches = 1;
            ensure_moment_buffers(std::max(ncols, 1), M);
            idx_buf.ensure(sizeof(int32_t) * std::max(ncols, 1));
            int32_t const zero = 0;
            PBK_CUDA(cudaMemcpyAsync(idx_buf.as(), &zero, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
            PBK_CUDA(launch_unit_starter(dtype, vec_a.as(), nloc, 1, idx_buf.as<int32_t>(), 1, stream));
            PBK_CUDA(cudaMemcpyAsync(idx_buf.as(), dest_pos.data(), sizeof(int32_t) * ncols, cudaMemcpyHostToDevice, stream));
            ++launches;
            run_offdiagonal(sub, M, true, [&](int k, void* r, double scale) {
                PBK_CUDA(launch_gather_moment(dtype, r, 1, idx_buf.as<int32_t>(), ncols, mom.as<double>(), M, k, scale, stream));
                ++launches;
            });
            PBK_CUDA(cudaMemcpyAsync(out, mom.as(), sizeof(cd) * static_cast<size_t>(ncols) * M, cudaMemcpyDeviceToHost, stream));
            stats.d2h_bytes += sizeof(cd) * static_cast<size_t>(ncols) * M;
            end_moments();
            return;
        }
    }
    auto& h = optimized_for(target);
    bool const opt = config.optimal_size != 0 && h.sliced;
    reset_stats(M, h, opt, 1);
    begin_moments();
    size_t const vbytes = static_cast<size_t>(n) * lane_pad(1) * dtype_size(dtype);
    vec_a.ensure(vbytes);
    vec_b.ensure(vbytes);
    stats.batch = 1;
    stats.num_batches = 1;
    idx_buf.ensure(sizeof(int32_t) * std::max(ncols, 1));
    if (h.idx.is_diagonal()) {  // Core.cpp:106-110
        ensure_moment_buffers(lane_pad(1), M);
        int const R = lane_pad(1);
        PBK_CUDA(cudaMemcpyAsync(idx_buf.as(), h.idx.src.data(), sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        PBK_CUDA(launch_unit_starter(dtype, vec_a.as(), n, R, idx_buf.as<int32_t>(), 1, stream));
        ++launches;
        run_diagonal(h, R, M, opt);
        PBK_CUDA(cudaMemcpyAsync(out, mom.as(), sizeof(cd) * M, cudaMemcpyDeviceToHost, stream));
    } else {                    // Core.cpp:111-115, MultiUnitCollector
        ensure_moment_buffers(std::max(ncols, 1), M);
        PBK_CUDA(cudaMemcpyAsync(idx_buf.as(), h.idx.src.data(), sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        PBK_CUDA(launch_unit_starter(dtype, vec_a.as(), n, 1, idx_buf.as<int32_t>(), 1, stream));
        PBK_CUDA(cudaMemcpyAsync(idx_buf.as(), h.idx.dest.data(), sizeof(int32_t) * ncols, cudaMemcpyHostToDevice, stream));
        ++launches;
        run_offdiagonal(h, M, opt, [&](int k, void* r, double scale) {
            PBK_CUDA(launch_gather_moment(dtype, r, 1, idx_buf.as<int32_t>(), ncols, mom.as<double>(), M, k, scale, stream));
            ++launches;
        });
        PBK_CUDA(cudaMemcpyAsync(out, mom.as(), sizeof(cd) * static_cast<size_t>(ncols) * M, cudaMemcpyDeviceToHost, stream));
    }
    stats.d2h_bytes += sizeof(cd) * static_cast<size_t>(ncols) * M;
    end_moments();
}

void Engine::moments_kubo(int M, const float* left, const float* right, int num_random, cd* out) {
    check_num_moments(M);
    if (num_random < 1) throw Error(PBK_INVALID_ARGUMENT, "num_random must be positive");
    double const t_kubo0 = now_seconds();
    auto& h = natural_hamiltonian();
    keep_kubo_buffers = true;
    reset_stats(M, h, false, num_random);
    keep_kubo_buffers = false;
    DeviceHamiltonian vl, vr;
    upload_operator(vl, left, h);
    upload_operator(vr, right, h);
    int const s = dtype_size(dtype);
    int first = 0, count = 0;
    shard(num_random, &first, &count);
    // All random vectors of this rank advance together as the lanes of one N x R block (one pass over H and over the
    // velocity operators serves them all), and the two M x (N R) stacks are contracted by ONE GEMM whose inner dimension
    // runs over sites and lanes: mu = sum_r L_r R_r^H (Core.cpp:140-144 does the same vector by vector).  The lanes per
    // pass are bounded by the memory of the two stacks.
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += DevBuf::cached_bytes();   // blocks parked in the allocation cache are given back on demand
    KuboStackLayout const probe = kubo_stack_layout(dtype, M, 256);      // row stride (payload + pad) of this scalar type
    double const pad = static_cast<double>(probe.row_stride) / 256.0;
    double const per_lane = (2.0 * M * pad + 5.0) * static_cast<double>(n) * s + 8.0 * static_cast<double>(n);
    int lanes_cap = static_cast<int>(0.85 * static_cast<double>(free_b) / per_lane);
    lanes_cap = std::min(lanes_cap, config.max_batch > 0 ? config.max_batch : 16);
    int const vmax = 16 / s;
    if (lanes_cap >= vmax) lanes_cap = lanes_cap / vmax * vmax;
    if (lanes_cap < 1) throw Error(PBK_RUNTIME_ERROR, "pbkpm: the two num_moments x system_size Kubo-Bastin stacks do not fit in device memory");
    int const nb = std::max(1, (std::max(count, 1) + lanes_cap - 1) / lanes_cap);
    int rb = (std::max(count, 1) + nb - 1) / nb;
    rb = std::min(lane_pad(rb), lanes_cap);
    size_t const vbytes = static_cast<size_t>(n) * rb * s;                 // one N x rb block
    // the stacks are k-blocked (kubo.cu): the step kernel writes row m of a stack through a blocked destination
    KuboStackLayout const layout = kubo_stack_layout(dtype, M, vbytes);
    int64_t const bstride = layout.block_stride;
    double const t_alloc = now_seconds();
    begin_moments();
    progress(-1, num_random);
    // the stacks stay with the context between calls (sigma_xx then sigma_xy on one object): releasing and re-mapping
    // 2 x 28 GB costs up to a second of host time per call; any other moments entry point gives them back (reset_stats)
    kubo_l.ensure(layout.bytes);
    kubo_r.ensure(layout.bytes);
    DevBuf& lstack = kubo_l; DevBuf& rstack = kubo_r;
    DevBuf u(vbytes), mu(sizeof(cd) * static_cast<size_t>(M) * M);
    if (std::getenv("PBK_TIMING")) std::fprintf(stderr, "[pbkpm] moments_kubo: operators + sizing %.3f s, stacks (2 x %.1f GB) allocated in %.3f s\n",
                                                 t_alloc - t_kubo0, layout.bytes / 1e9, now_seconds() - t_alloc);
    kubo_ws.ensure(kubo_gemm_workspace_bytes(M, layout.blocks, num_sms));
    DevBuf& gemm_ws = kubo_ws;
    vec_a.ensure(vbytes);
    vec_b.ensure(vbytes);
    ensure_moment_buffers(rb, M);
    PBK_CUDA(cudaMemsetAsync(mu.as(), 0, mu.bytes(), stream));
    seed_stream(first);
    stats.batch = rb;
    auto row_of = [&](DevBuf& st, int k) { return static_cast<void*>(st.as<char>() + static_cast<size_t>(k) * layout.row_stride); };
    for (int b0 = 0; b0 < count; b0 += rb) {
        int const lanes = std::min(rb, count - b0);
        int const R = lane_pad(lanes);                     // padded lanes are zero vectors: they add nothing to mu
        // the blocks this pass fills; the tail of the last one (past the end of the vectors) must read as zero
        KuboStackLayout const used = kubo_stack_layout(dtype, M, static_cast<size_t>(n) * R * s);
        PBK_CUDA(cudaMemsetAsync(lstack.as<char>() + static_cast<size_t>(used.blocks - 1) * bstride, 0, static_cast<size_t>(bstride), stream));
        PBK_CUDA(cudaMemsetAsync(rstack.as<char>() + static_cast<size_t>(used.blocks - 1) * bstride, 0, static_cast<size_t>(bstride), stream));
        generate_random_block(h, lanes, R, u.as());
        PBK_CUDA(cudaEventRecord(ev2, stream));
        // left: starter v_l|r>, rows are T_n(H) v_l |r>            (Core.cpp:131-133, DenseMatrixCollector)
        void* r0 = vec_a.as(); void* r1 = vec_b.as();
        step(vl, u.as(), r0, nullptr, n, R, false, false, 1.0, M, 0, FIN_NONE);
        step(vl, u.as(), row_of(lstack, 0), nullptr, n, R, false, false, 0.5, M, 0, FIN_NONE, bstride, 0);
        step(h, r0, r1, row_of(lstack, 1), n, R, false, false, 0.5, M, 0, FIN_NONE, 0, bstride);
        for (int k = 2; k < M; ++k) { step(h, r1, r0, row_of(lstack, k), n, R, true, false, 1.0, M, k, FIN_NONE, 0, bstride); std::swap(r0, r1); }
        // right: starter |r>, rows are v_r T_n(H)|r>                 (Core.cpp:135-137)
        r0 = u.as(); r1 = vec_b.as();
        step(vr, r0, row_of(rstack, 0), nullptr, n, R, false, false, 0.5, M, 0, FIN_NONE, bstride, 0);
        step(h, r0, r1, nullptr, n, R, false, false, 0.5, M, 0, FIN_NONE);
        step(vr, r1, row_of(rstack, 1), nullptr, n, R, false, false, 1.0, M, 0, FIN_NONE, bstride, 0);
        // r0 (= u) is overwritten from here on; its content is no longer needed
        for (int k = 2; k < M; ++k) {
            step(h, r1, r0, nullptr, n, R, true, false, 1.0, M, k, FIN_NONE);
            std::swap(r0, r1);
            step(vr, r1, row_of(rstack, k), nullptr, n, R, false, false, 1.0, M, k, FIN_NONE, bstride, 0);
        }
        PBK_CUDA(cudaEventRecord(ev3, stream));
        double flops = 0;
        PBK_CUDA(launch_kubo_gemm(dtype, lstack.as(), rstack.as(), M, static_cast<int64_t>(n) * R, used, mu.as<double>(),
                                  gemm_ws.as<double>(), gemm_ws.bytes(), num_sms, stream, &flops));
        launches += dtype_complex(dtype) ? 4 : 2;
        PBK_CUDA(cudaEventRecord(ev1, stream));
        PBK_CUDA(cudaEventSynchronize(ev1));
        float ms = 0;
        PBK_CUDA(cudaEventElapsedTime(&ms, ev2, ev3));
        stats.step_ms += ms;
        PBK_CUDA(cudaEventElapsedTime(&ms, ev3, ev1));
        stats.gemm_ms += ms;
        stats.gemm_flops += flops;
        ++stats.num_batches;
        progress(lanes, num_random);
    }
    allreduce(mu.as<double>(), 2LL * M * M);
    PBK_CUDA(cudaMemcpyAsync(out, mu.as(), mu.bytes(), cudaMemcpyDeviceToHost, stream));
    stats.d2h_bytes += static_cast<int64_t>(mu.bytes());
    end_moments();
    for (size_t i = 0; i < static_cast<size_t>(M) * M; ++i) out[i] /= static_cast<double>(num_random);  // Moments.cpp:127-130
    progress(num_random, num_random);
}

// ------------------------------------------------------------------------------------------------
// kpm::Core level
// ------------------------------------------------------------------------------------------------
void Engine::core_moments(int num_moments, const cd* alpha, const cd* beta, int64_t op_rows, const int32_t* op_indptr,
                          const int32_t* op_indices, const cd* op_data, cd* out) {
    if (num_moments < 1) throw Error(PBK_INVALID_ARGUMENT, "num_moments must be positive");
    int const M = round_num_moments(num_moments);
    auto& h = natural_hamiltonian();
    std::vector<cd> m(M);
    if (!beta && op_rows == 0) {  // Core.cpp:45-49
        moments_diagonal(M, alpha, 1, m.data());
    } else {                      // Core.cpp:50-55, GenericCollector
        reset_stats(M, h, false, 1);
        DeviceHamiltonian op;
        if (op_rows != 0) upload_csr_operator(op, op_rows, op_indptr, op_indices, op_data, h);
        begin_moments();
        size_t const vbytes = static_cast<size_t>(n) * dtype_size(dtype);
        vec_a.ensure(vbytes);
        vec_b.ensure(vbytes);
        vec_t.ensure(vbytes);
        ensure_moment_buffers(1, M);
        DevBuf staging(sizeof(double) * 2 * n), beta_dev(vbytes);
        PBK_CUDA(cudaMemcpyAsync(staging.as(), beta ? beta : alpha, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, stream));
        PBK_CUDA(launch_scatter_block(dtype, staging.as<double>(), n, 1, 0, h.reordered ? h.perm.as<int32_t>() : nullptr, beta_dev.as(), stream));
        PBK_CUDA(cudaStreamSynchronize(stream));
        PBK_CUDA(cudaMemcpyAsync(staging.as(), alpha, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, stream));
        PBK_CUDA(launch_scatter_block(dtype, staging.as<double>(), n, 1, 0, h.reordered ? h.perm.as<int32_t>() : nullptr, vec_a.as(), stream));
        stats.h2d_bytes += 2 * sizeof(double) * 2 * n;
        run_offdiagonal(h, M, false, [&](int k, void* r, double scale) {
            const void* v = r;
            if (op.valid) {
                step(op, r, vec_t.as(), nullptr, n, 1, false, false, 1.0, M, k, FIN_NONE);
                v = vec_t.as();
            }
            PBK_CUDA(launch_dot_moment(dtype, beta_dev.as(), v, n, mom.as<double>(), k, scale, scratch.as<double>(), counter.as<unsigned>(), num_sms, stream));
            ++launches;
        });
        PBK_CUDA(cudaMemcpyAsync(m.data(), mom.as(), sizeof(cd) * M, cudaMemcpyDeviceToHost, stream));
        end_moments();
    }
    auto const g = damping_coefficients(config.kernel, config.lambda_value, M);
    for (int i = 0; i < num_moments; ++i) out[i] = m[i] * g[i];
}

void Engine::calc_dos(const double* energy, int ne, double broadening, int num_random, double* out) {
    double const t0 = now_seconds();
    auto const s = scaling_factors();
    int const M = required_num_moments(broadening);
    std::vector<cd> m(M);
    double const t1 = now_seconds();
    moments_dos(M, num_random, m.data());
    double const t2 = now_seconds();
    auto const g = damping_coefficients(config.kernel, config.lambda_value, M);
    for (int i = 0; i < M; ++i) m[i] *= g[i];
    spectral_density_device(m.data(), M, 1, 0, 1, energy, ne, s, out);
    last_total_seconds = now_seconds() - t0;
    if (std::getenv("PBK_TIMING")) std::fprintf(stderr, "[pbkpm] calc_dos: setup %.3f s, moments_dos %.3f s (layout %.3f, moments phase wall %.3f, device %.3f), reconstruction %.3f s\n",
                                                 t1 - t0, t2 - t1, stats.hamiltonian_time, stats.moments_time, stats.moments_device_ms * 1e-3, now_seconds() - t2);
}

void Engine::calc_ldos(const double* energy, int ne, double broadening, const int32_t* idx, int nidx, double* out) {
    double const t0 = now_seconds();
    auto const s = scaling_factors();
    int const M = required_num_moments(broadening);
    std::vector<cd> m(static_cast<size_t>(M) * nidx);
    moments_ldos(M, idx, nidx, m.data());
    auto const g = damping_coefficients(config.kernel, config.lambda_value, M);
    for (int k = 0; k < M; ++k) for (int i = 0; i < nidx; ++i) m[static_cast<size_t>(k) * nidx + i] *= g[k];
    spectral_density_device(m.data(), M, nidx, 1, nidx, energy, ne, s, out);
    last_total_seconds = now_seconds() - t0;
}

void Engine::calc_greens(int row, const int32_t* cols, int ncols, const double* energy, int ne, double broadening, cd* out) {
    double const t0 = now_seconds();
    bool const timing = std::getenv("PBK_TIMING") != nullptr;
    auto const s = scaling_factors();
    int const M = required_num_moments(broadening);
    std::vector<cd> m(static_cast<size_t>(M) * ncols);
    double const t1 = now_seconds();
    moments_greens(M, row, cols, ncols, m.data());
    double const t2 = now_seconds();
    auto const g = damping_coefficients(config.kernel, config.lambda_value, M);
    for (int i = 0; i < ncols; ++i) for (int k = 0; k < M; ++k) m[static_cast<size_t>(i) * M + k] *= g[k];
    greens_device(m.data(), M, ncols, energy, ne, s, out);
    last_total_seconds = now_seconds() - t0;
    if (timing) std::fprintf(stderr, "[pbkpm] calc_greens: setup %.3f s, moments_greens %.3f s (hamiltonian %.3f, moments phase %.3f), reconstruction %.3f s\n",
                             t1 - t0, t2 - t1, stats.hamiltonian_time, stats.moments_time, now_seconds() - t2);
}

void Engine::calc_conductivity(const float* left, const float* right, const double* mu, int nmu, double broadening,
                               double temperature, int num_random, int num_points, cd* out) {
    double const t0 = now_seconds();
    auto const s = scaling_factors();
    int const M = required_num_moments(broadening);
    std::vector<cd> m(static_cast<size_t>(M) * M);
    double const t1 = now_seconds();
    moments_kubo(M, left, right, num_random, m.data());
    double const t2 = now_seconds();
    auto const g = damping_coefficients(config.kernel, config.lambda_value, M);
    for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) m[static_cast<size_t>(i) * M + j] *= g[i] * g[j];  // Kernel.hpp:49-56

    // energy samples: linspace over the *unscaled* bounds (Bounds.hpp:54), then scaled
    std::vector<double> samples(num_points);
    for (int i = 0; i < num_points; ++i) {
        double const e = (i == num_points - 1) ? bounds_max : bounds_min + i * ((bounds_max - bounds_min) / std::max(1, num_points - 1));
        samples[i] = (e - s.b) / s.a;
    }
    // sum_nm(E) = 1/(1-E^2)^2 * sum_{m,n} mu_mn * Gamma_mn(E): O(points * M^2) -> on the device
    DevBuf mu_dev(sizeof(cd) * m.size()), samples_dev(sizeof(double) * num_points), sum_dev(sizeof(cd) * num_points);
    PBK_CUDA(cudaMemcpyAsync(mu_dev.as(), m.data(), sizeof(cd) * m.size(), cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaMemcpyAsync(samples_dev.as(), samples.data(), sizeof(double) * num_points, cudaMemcpyHostToDevice, stream));
    PBK_CUDA(launch_kubo_gamma_sum(mu_dev.as<double>(), M, samples_dev.as<double>(), num_points, sum_dev.as<double>(), stream));
    std::vector<double> sum_nm(2 * static_cast<size_t>(num_points));
    PBK_CUDA(cudaMemcpyAsync(sum_nm.data(), sum_dev.as(), sizeof(cd) * num_points, cudaMemcpyDeviceToHost, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
    double const t3 = now_seconds();
    reconstruct_kubo_bastin(sum_nm.data(), samples, mu, nmu, temperature, s, out);
    last_total_seconds = now_seconds() - t0;
    if (std::getenv("PBK_TIMING")) std::fprintf(stderr, "[pbkpm] calc_conductivity: setup %.3f s, moments_kubo %.3f s (device %.3f: recursion %.3f, GEMM %.3f), damping + Gamma sum %.3f s, Fermi integration %.3f s\n",
                                                 t1 - t0, t2 - t1, stats.moments_device_ms * 1e-3, stats.step_ms * 1e-3, stats.gemm_ms * 1e-3, t3 - t2, now_seconds() - t3);
}

// ------------------------------------------------------------------------------------------------
// Reconstruction on the device (reconstruct.cu): damped moments go up, curves come back
// ------------------------------------------------------------------------------------------------
void Engine::spectral_density_device(const cd* moments, int M, int cols, int64_t col_stride, int64_t n_stride, const double* energy, int ne,
                                     Scale s, double* out) {
    if (ne <= 0) return;
    PBK_CUDA(cudaSetDevice(device));
    size_t const count = static_cast<size_t>(M) * std::max<int64_t>(n_stride, 1);
    std::vector<double> scaled(ne);
    for (int i = 0; i < ne; ++i) scaled[i] = (energy[i] - s.b) / s.a;
    DevBuf mu_dev(sizeof(cd) * count), e_dev(sizeof(double) * ne), out_dev(sizeof(double) * static_cast<size_t>(ne) * cols);
    PBK_CUDA(cudaMemcpyAsync(mu_dev.as(), moments, sizeof(cd) * count, cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaMemcpyAsync(e_dev.as(), scaled.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, stream));
    double const k = static_cast<double>(2 / pi_f) / s.a;  // real_t{2 / constant::pi}
    PBK_CUDA(launch_spectral_density(mu_dev.as<double>(), M, cols, col_stride, n_stride, e_dev.as<double>(), ne, k, out_dev.as<double>(), stream));
    PBK_CUDA(cudaMemcpyAsync(out, out_dev.as(), sizeof(double) * static_cast<size_t>(ne) * cols, cudaMemcpyDeviceToHost, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
}

void Engine::greens_device(const cd* moments, int M, int cols, const double* energy, int ne, Scale s, cd* out) {
    if (ne <= 0) return;
    PBK_CUDA(cudaSetDevice(device));
    std::vector<double> scaled(ne);
    for (int i = 0; i < ne; ++i) scaled[i] = (energy[i] - s.b) / s.a;
    size_t const count = static_cast<size_t>(M) * cols;
    DevBuf mu_dev(sizeof(cd) * count), e_dev(sizeof(double) * ne), out_dev(sizeof(cd) * static_cast<size_t>(ne) * cols);
    PBK_CUDA(cudaMemcpyAsync(mu_dev.as(), moments, sizeof(cd) * count, cudaMemcpyHostToDevice, stream));
    PBK_CUDA(cudaMemcpyAsync(e_dev.as(), scaled.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, stream));
    PBK_CUDA(launch_greens(mu_dev.as<double>(), M, cols, e_dev.as<double>(), ne, 1.0 / s.a, out_dev.as<double>(), stream));
    PBK_CUDA(cudaMemcpyAsync(out, out_dev.as(), sizeof(cd) * static_cast<size_t>(ne) * cols, cudaMemcpyDeviceToHost, stream));
    PBK_CUDA(cudaStreamSynchronize(stream));
}

// ------------------------------------------------------------------------------------------------
// Reporting (Core::report, Bounds::report, Stats::report)
// ------------------------------------------------------------------------------------------------
static std::string with_suffix(double v) {
    char const* suffix[] = {"", "k", "M", "G", "T", "P"};
    int i = 0;
    while (std::abs(v) >= 1000 && i < 5) { v /= 1000; ++i; }
    char buf[64];
    std::snprintf(buf, sizeof(buf), (i == 0 || v >= 100) ? "%.0f%s" : (v >= 10 ? "%.1f%s" : "%.2f%s"), v, suffix[i]);
    return buf;
}
static std::string duration(double s) {
    char buf[64];
    if (s < 1e-3) std::snprintf(buf, sizeof(buf), "%.0fus", s * 1e6);
    else if (s < 1) std::snprintf(buf, sizeof(buf), "%.2fms", s * 1e3);
    else std::snprintf(buf, sizeof(buf), "%.2fs", s);
    return buf;
}

std::string Engine::report(bool shortform) const {
    char buf[1024];
    double const removed = stats.nnz > 0 ? 100.0 * static_cast<double>(stats.nnz - stats.opt_nnz) / static_cast<double>(stats.nnz) : 0.0;
    char const* star = stats.uses_full_system ? "" : "*";
    if (shortform) {
        std::snprintf(buf, sizeof(buf), "%.2f, %.2f, %d [%s] %.0f%%%s [%s] %s @ %seps [%s] | %s", bounds_min, bounds_max, lanczos_loops,
                      duration(bounds_seconds).c_str(), removed, star, duration(stats.hamiltonian_time).c_str(),
                      with_suffix(static_cast<double>(stats.num_moments)).c_str(), with_suffix(stats.eps).c_str(),
                      duration(stats.moments_time).c_str(), duration(last_total_seconds).c_str());
    } else {
        std::snprintf(buf, sizeof(buf),
                      "- Spectrum bounds found (%.2f, %.2f eV) using Lanczos procedure with %d loops | %s\n"
                      "- The reordering optimization was able to remove %.0f%%%s of the workload | %s\n"
                      "- KPM calculated %s moments at %s non-zero elements per second (B200, %d vectors/pass) | %s\n"
                      "Total time: %s",
                      bounds_min, bounds_max, lanczos_loops, duration(bounds_seconds).c_str(), removed, star,
                      duration(stats.hamiltonian_time).c_str(), with_suffix(static_cast<double>(stats.num_moments)).c_str(),
                      with_suffix(stats.eps).c_str(), stats.batch, duration(stats.moments_time).c_str(),
                      duration(last_total_seconds).c_str());
    }
    return buf;
}

// ------------------------------------------------------------------------------------------------
// Kubo-Bastin: Fermi-weighted integration over the energy samples (kpm/reconstruct.hpp:133-143); the O(points * M^2)
// Gamma-matrix sum runs on the device (kubo.cu).  Double precision; the reference's float constants are kept
// ------------------------------------------------------------------------------------------------
void reconstruct_kubo_bastin(const double* sum_nm, const std::vector<double>& en, const double* mu, int nmu,
                             double temperature, Scale s, cd* out) {
    int const np = static_cast<int>(en.size());
    double const inv_kbt_sc = s.a / (kb_f * temperature);
    double const en_max = *std::max_element(en.begin(), en.end());
    double const en_min = *std::min_element(en.begin(), en.end());
    double const coeff = (en_max - en_min) / static_cast<double>(2 * np);
    cd const prefix = cd(4.0) / (s.a * s.a);
    for (int j = 0; j < nmu; ++j) {
        double const mi = (mu[j] - s.b) / s.a;
        cd total(0, 0), first(0, 0), last(0, 0);
        for (int p = 0; p < np; ++p) {
            double const fd = 1.0 / (1.0 + std::exp((en[p] - mi) * inv_kbt_sc));
            cd const f = fd * cd(sum_nm[2 * p], sum_nm[2 * p + 1]);
            total += f;
            if (p == 0) first = f;
            if (p == np - 1) last = f;
        }
        out[j] = prefix * (coeff * (2.0 * total - first - last));
    }
}

} // namespace pbk
