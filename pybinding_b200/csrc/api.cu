// api.cu -- the extern "C" surface declared in include/pbkpm.h.  Exceptions never cross the ABI: every
// entry point returns a pbk_status and stores the message in the context (pbk_last_error).
#include "engine.hpp"

#include <cstdlib>
#include <cstring>

using namespace pbk;

struct pbk_ctx {
    std::unique_ptr<Engine> engine;
    std::string error;
};

namespace {
thread_local std::string create_error;

template<class F> int guarded(pbk_ctx* ctx, F fn) {
    if (!ctx || !ctx->engine) return PBK_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(ctx->engine->mutex);
    try {
        fn(*ctx->engine);
        return PBK_OK;
    } catch (Error const& e) {
        ctx->error = e.what();
        return e.code;
    } catch (std::bad_alloc const&) {
        ctx->error = "pbkpm: out of host memory";
        return PBK_RUNTIME_ERROR;
    } catch (std::exception const& e) {
        ctx->error = e.what();
        return PBK_RUNTIME_ERROR;
    }
}
} // anonymous namespace

extern "C" {

int pbk_version(void) { return PBK_VERSION; }

int pbk_device_count(int* count) {
    int c = 0;
    cudaError_t const err = cudaGetDeviceCount(&c);
    if (count) *count = (err == cudaSuccess) ? c : 0;
    return err == cudaSuccess ? PBK_OK : PBK_CUDA_ERROR;
}

int pbk_create(pbk_ctx** out, int device, const pbk_config* config) {
    if (!out) return PBK_INVALID_ARGUMENT;
    *out = nullptr;
    pbk_config cfg{};
    cfg.kernel = PBK_JACKSON;
    cfg.lambda_value = 4.0;
    cfg.optimal_size = 1;
    cfg.interleaved = 1;
    cfg.matrix_format = 1;
    cfg.lanczos_precision = 0.002f;
    if (config) cfg = *config;
    try {
        auto ctx = std::make_unique<pbk_ctx>();
        ctx->engine = std::make_unique<Engine>(device, cfg);
        *out = ctx.release();
        return PBK_OK;
    } catch (Error const& e) {
        create_error = e.what();
        return e.code;
    } catch (std::exception const& e) {
        create_error = e.what();
        return PBK_RUNTIME_ERROR;
    }
}

void pbk_destroy(pbk_ctx* ctx) { delete ctx; }

const char* pbk_last_error(const pbk_ctx* ctx) { return ctx ? ctx->error.c_str() : create_error.c_str(); }

int pbk_set_progress_callback(pbk_ctx* ctx, pbk_progress_fn fn, void* user) {
    return guarded(ctx, [&](Engine& e) { e.set_progress(fn, user); });
}

int pbk_set_hamiltonian(pbk_ctx* ctx, int dtype, int64_t n, const int32_t* indptr, const int32_t* indices, const void* data) {
    return guarded(ctx, [&](Engine& e) { e.set_hamiltonian(dtype, n, indptr, indices, data); });
}

int pbk_bounds(pbk_ctx* ctx, double* mn, double* mx, int32_t* loops) {
    return guarded(ctx, [&](Engine& e) { double a, b; int32_t l; e.bounds(&a, &b, &l); if (mn) *mn = a; if (mx) *mx = b; if (loops) *loops = l; });
}

int pbk_scaling_factors(pbk_ctx* ctx, double* a, double* b) {
    return guarded(ctx, [&](Engine& e) { auto const s = e.scaling_factors(); if (a) *a = s.a; if (b) *b = s.b; });
}

int pbk_required_num_moments(pbk_ctx* ctx, double broadening, int32_t* num_moments) {
    return guarded(ctx, [&](Engine& e) { *num_moments = e.required_num_moments(broadening); });
}

int pbk_kernel_damping(int kernel, double lambda_value, int32_t n, double* out) {
    if (n < 0 || !out || kernel < 0 || kernel > 2) return PBK_INVALID_ARGUMENT;
    auto const g = damping_coefficients(kernel, lambda_value, n);
    std::memcpy(out, g.data(), sizeof(double) * n);
    return PBK_OK;
}

int pbk_kernel_required_num_moments(int kernel, double lambda_value, double scaled_broadening, int32_t* out) {
    if (!out || kernel < 0 || kernel > 2) return PBK_INVALID_ARGUMENT;
    *out = kernel_required_num_moments(kernel, lambda_value, scaled_broadening);
    return PBK_OK;
}

int pbk_locality_order(int64_t n, const int32_t* indptr, const int32_t* indices, int32_t tile, int32_t* order) {
    if (n <= 0 || !indptr || !indices || !order || tile < 1) return PBK_INVALID_ARGUMENT;
    std::vector<int32_t> queue, rmap;
    cluster_order(n, indptr, indices, tile, queue, rmap);
    std::memcpy(order, queue.data(), sizeof(int32_t) * static_cast<size_t>(n));
    return PBK_OK;
}

int pbk_locality_order2(int64_t n, const int32_t* indptr, const int32_t* indices, int32_t tile, int32_t macro_tiles, int32_t* order) {
    if (n <= 0 || !indptr || !indices || !order || tile < 1 || macro_tiles < 0) return PBK_INVALID_ARGUMENT;
    std::vector<int32_t> queue, rmap;
    char const* coarse_env = std::getenv("PBK_COARSE");   // same knob as the engine (default 16 sites per super-node)
    cluster_order(n, indptr, indices, tile, queue, rmap, macro_tiles, coarse_env ? std::strtol(coarse_env, nullptr, 10) : 16);
    std::memcpy(order, queue.data(), sizeof(int32_t) * static_cast<size_t>(n));
    return PBK_OK;
}

int pbk_host_ell(int dtype, int64_t n, const int32_t* indptr, const int32_t* indices, const void* data, double min_energy,
                 double max_energy, const int32_t* order, int32_t* k, int64_t* pitch, void* val, int32_t* col) {
    if (n <= 0 || !indptr || !indices || !data || !k || !pitch || !(min_energy < max_energy)) return PBK_INVALID_ARGUMENT;
    try {
        return host_scaled_ell(dtype, n, indptr, indices, data, min_energy, max_energy, order, k, pitch, val, col);
    } catch (...) {
        return PBK_RUNTIME_ERROR;
    }
}

int pbk_light_cone(int64_t n, const int32_t* indptr, const int32_t* indices, int32_t src, int32_t depth,
                   int32_t* queue, int64_t* queue_size, int32_t* borders, int32_t* num_borders, int32_t* exhausted) {
    if (n <= 0 || !indptr || !indices || src < 0 || src >= n || depth < 0 || !queue || !queue_size || !borders || !num_borders || !exhausted) {
        return PBK_INVALID_ARGUMENT;
    }
    std::vector<int32_t> mark(static_cast<size_t>(n), -1);
    Cone const c = light_cone(indptr, indices, src, depth, mark);
    std::memcpy(queue, c.queue.data(), sizeof(int32_t) * c.queue.size());
    std::memcpy(borders, c.borders.data(), sizeof(int32_t) * c.borders.size());
    *queue_size = static_cast<int64_t>(c.queue.size());
    *num_borders = static_cast<int32_t>(c.borders.size());
    *exhausted = c.exhausted ? 1 : 0;
    return PBK_OK;
}

int pbk_shard(int32_t total, int32_t world_size, int32_t rank, int32_t* first, int32_t* count) {
    if (total < 0 || world_size < 1 || rank < 0 || rank >= world_size || !first || !count) return PBK_INVALID_ARGUMENT;
    int f = 0, c = 0;
    shard_range(total, world_size, rank, &f, &c);
    *first = f; *count = c;
    return PBK_OK;
}

int pbk_mt_jump_window(uint64_t position, uint32_t* window) {
    if (!window) return PBK_INVALID_ARGUMENT;
    return mt_jump_window_host(position, window) ? PBK_OK : PBK_RUNTIME_ERROR;
}

int pbk_moments_dos(pbk_ctx* ctx, int32_t num_moments, int32_t num_random, void* out) {
    return guarded(ctx, [&](Engine& e) { e.moments_dos(num_moments, num_random, static_cast<cd*>(out)); });
}

int pbk_moments_ldos(pbk_ctx* ctx, int32_t num_moments, const int32_t* idx, int32_t nidx, void* out) {
    return guarded(ctx, [&](Engine& e) { e.moments_ldos(num_moments, idx, nidx, static_cast<cd*>(out)); });
}

int pbk_moments_greens(pbk_ctx* ctx, int32_t num_moments, int32_t row, const int32_t* cols, int32_t ncols, void* out) {
    return guarded(ctx, [&](Engine& e) { e.moments_greens(num_moments, row, cols, ncols, static_cast<cd*>(out)); });
}

int pbk_moments_kubo(pbk_ctx* ctx, int32_t num_moments, const float* left, const float* right, int32_t num_random, void* out) {
    return guarded(ctx, [&](Engine& e) { e.moments_kubo(num_moments, left, right, num_random, static_cast<cd*>(out)); });
}

int pbk_moments_diagonal(pbk_ctx* ctx, int32_t num_moments, const void* r0, int32_t count, void* out) {
    return guarded(ctx, [&](Engine& e) { e.moments_diagonal(num_moments, static_cast<const cd*>(r0), count, static_cast<cd*>(out)); });
}

int pbk_random_vectors(pbk_ctx* ctx, int32_t count, void* out) {
    return guarded(ctx, [&](Engine& e) { e.random_vectors(count, static_cast<cd*>(out)); });
}

int pbk_moments(pbk_ctx* ctx, int32_t num_moments, const void* alpha, const void* beta, int64_t op_rows,
                const int32_t* op_indptr, const int32_t* op_indices, const void* op_data, void* out) {
    return guarded(ctx, [&](Engine& e) {
        e.core_moments(num_moments, static_cast<const cd*>(alpha), static_cast<const cd*>(beta), op_rows, op_indptr, op_indices,
                       static_cast<const cd*>(op_data), static_cast<cd*>(out));
    });
}

int pbk_calc_dos(pbk_ctx* ctx, const double* energy, int32_t ne, double broadening, int32_t num_random, double* out) {
    return guarded(ctx, [&](Engine& e) { e.calc_dos(energy, ne, broadening, num_random, out); });
}

int pbk_calc_ldos(pbk_ctx* ctx, const double* energy, int32_t ne, double broadening, const int32_t* idx, int32_t nidx, double* out) {
    return guarded(ctx, [&](Engine& e) { e.calc_ldos(energy, ne, broadening, idx, nidx, out); });
}

int pbk_calc_greens(pbk_ctx* ctx, int32_t row, const int32_t* cols, int32_t ncols, const double* energy, int32_t ne,
                    double broadening, void* out) {
    return guarded(ctx, [&](Engine& e) { e.calc_greens(row, cols, ncols, energy, ne, broadening, static_cast<cd*>(out)); });
}

int pbk_calc_conductivity(pbk_ctx* ctx, const float* left, const float* right, const double* mu, int32_t nmu, double broadening,
                          double temperature, int32_t num_random, int32_t num_points, void* out) {
    return guarded(ctx, [&](Engine& e) {
        e.calc_conductivity(left, right, mu, nmu, broadening, temperature, num_random, num_points, static_cast<cd*>(out));
    });
}

int pbk_get_stats(pbk_ctx* ctx, pbk_stats* out) {
    return guarded(ctx, [&](Engine& e) { *out = e.get_stats(); });
}

int pbk_report(pbk_ctx* ctx, int shortform, char* buffer, int64_t size) {
    return guarded(ctx, [&](Engine& e) {
        auto const r = e.report(shortform != 0);
        if (size > 0) { std::strncpy(buffer, r.c_str(), static_cast<size_t>(size) - 1); buffer[size - 1] = '\0'; }
    });
}

int pbk_comm_unique_id(char id[128]) {
    try {
        NcclApi* api = nullptr;
        static std::unique_ptr<NcclApi> holder;
        if (!holder) holder = std::make_unique<NcclApi>();
        api = holder.get();
        api->check(api->GetUniqueId(id), "ncclGetUniqueId");
        return PBK_OK;
    } catch (std::exception const& e) {
        create_error = e.what();
        return PBK_NCCL_ERROR;
    }
}

int pbk_comm_init(pbk_ctx* ctx, int32_t world_size, int32_t rank, const char id[128]) {
    return guarded(ctx, [&](Engine& e) { e.comm_init(world_size, rank, id); });
}

int pbk_comm_destroy(pbk_ctx* ctx) {
    return guarded(ctx, [&](Engine& e) { e.comm_destroy(); });
}

} // extern "C"
