// kernels_persist.cu -- K1 for launch-bound systems: the WHOLE diagonal recursion of one vector in ONE persistent kernel.
//
// A 40 x 40 nm graphene sheet (BASELINE configs[0]: 61 k sites, one random vector, 1026 moments) moves 2 MB per
// Chebyshev step -- a few hundred nanoseconds of memory traffic against several microseconds of launch latency per
// step, even when the 512 launches are replayed as one CUDA graph.  Here the grid is launched once (cooperatively: every
// CTA is resident) and keeps everything that belongs to a row in registers for the whole run:
//   * the row's matrix elements (col[K], val[K]) are loaded once,
//   * r_{n-1}[row] and r_{n-2}[row] never leave the owning thread; only the new value is published to a global buffer
//     (L2) for the neighbours, which gather it with ld.global.cg after a grid-wide barrier,
//   * the two sums of a step go to per-CTA slots of a [step][CTA] table (warp shuffles + one shared-memory pass, no
//     atomics on data, fixed order), and a small second kernel folds the table into the moments.
// One grid barrier per step (sense-free monotonic counter: arrive with atomicAdd, spin on a volatile load) replaces one
// kernel launch per step.  Same arithmetic per row as `cheb_step` (FMA order over the slots, f64 sums).
//
// Replaces, from the reference (cppcore/): calc_moments::basic (diagonal) + compute::kpm_spmv_diagonal for a single
// vector (include/kpm/calc_moments.hpp:36-51, include/compute/kernel_polynomial.hpp:61-81) and the DiagonalCollector.
#include "kernels.cuh"
#include "step_common.cuh"

#include <cooperative_groups.h>

namespace pbk {
namespace {

constexpr int PERSIST_TPB = 256;

struct PersistDev {
    const void* val; const int32_t* col; int64_t pitch; int k;
    int nrows;
    void* buf[2];             // buf[0] holds r0 on entry; r_n is published to buf[n & 1]
    int steps;                // M / 2: step 1 is r1 = H r0 / 2, steps 2 .. M / 2 the recursion
    double* table;            // [steps][gridDim.x][C] partial sums
    unsigned* barrier;        // zeroed before the launch
};

template<class T> __device__ __forceinline__ T ldcg_(const T* p);
template<> __device__ __forceinline__ float ldcg_(const float* p) { return __ldcg(p); }
template<> __device__ __forceinline__ double ldcg_(const double* p) { return __ldcg(p); }
template<> __device__ __forceinline__ float2 ldcg_(const float2* p) { return __ldcg(p); }
template<> __device__ __forceinline__ double2 ldcg_(const double2* p) { return __ldcg(p); }

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                       // the CTA's published values, observed through the barrier above, before the arrival
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

/// block sum of C doubles per thread -> table slot of this CTA (fixed order: lanes by shuffle, warps in index order)
template<int C>
__device__ __forceinline__ void block_sums(double (&acc)[C], double* slot) {
    __shared__ double sm[PERSIST_TPB / 32][C];
#pragma unroll
    for (int q = 0; q < C; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PERSIST_TPB / 32; ++w) s += sm[w][threadIdx.x];
        slot[threadIdx.x] = s;
    }
    // (the grid barrier that follows starts with __syncthreads: sm is free again before its next use)
}

template<class T, int K, int RPT>
__global__ void __launch_bounds__(PERSIST_TPB) cheb_persistent(PersistDev a) {
    constexpr int C = ST<T>::C;
    const T* __restrict__ val = static_cast<const T*>(a.val);
    int const stride = gridDim.x * PERSIST_TPB;
    int const first = blockIdx.x * PERSIST_TPB + threadIdx.x;

    int32_t c[RPT][K]; T v[RPT][K];
    T xo[RPT], yo[RPT];                        // r_{n-1}[row], r_{n-2}[row]
    bool own[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        int const row = first + j * stride;
        own[j] = row < a.nrows;
        xo[j] = zero_(T{}); yo[j] = zero_(T{});
#pragma unroll
        for (int s = 0; s < K; ++s) { c[j][s] = 0; v[j][s] = zero_(T{}); }
        if (own[j]) {
#pragma unroll
            for (int s = 0; s < K; ++s) { c[j][s] = a.col[s * a.pitch + row]; v[j][s] = val[s * a.pitch + row]; }
            xo[j] = static_cast<const T*>(a.buf[0])[row];
        }
    }
    unsigned arrivals = 0;
    for (int n = 1; n <= a.steps; ++n) {
        const T* __restrict__ x = static_cast<const T*>(a.buf[(n - 1) & 1]);
        T* __restrict__ y = static_cast<T*>(a.buf[n & 1]);
        double acc[C];
#pragma unroll
        for (int q = 0; q < C; ++q) acc[q] = 0.0;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            if (!own[j]) continue;
            T xg[K];
#pragma unroll
            for (int s = 0; s < K; ++s) xg[s] = ldcg_(x + c[j][s]);
            T r = n == 1 ? zero_(T{}) : neg_(yo[j]);
#pragma unroll
            for (int s = 0; s < K; ++s) r = fma_(v[j][s], xg[s], r);
            if (n == 1) r = scale_(r, 0.5);
            sums_(acc, xo[j], r);
            y[first + j * stride] = r;
            yo[j] = xo[j]; xo[j] = r;
        }
        block_sums<C>(acc, a.table + (static_cast<int64_t>(n - 1) * gridDim.x + blockIdx.x) * C);
        if (n < a.steps) { arrivals += gridDim.x; grid_barrier(a.barrier, arrivals); }
    }
}

/// moments from the [step][CTA] table: fixed-order sum over the CTAs, then the bookkeeping of the Diagonal collector
/// (src/kpm/default/collectors.cpp:6-34): mu_0 = s0 / 2, mu_1 = s1, mu_{2(n-1)} = 2 (s0_n - mu_0), mu_{2n-1} = 2 s1_n - mu_1
template<int C>
__global__ void persist_finish_kernel(const double* __restrict__ table, int grid, int steps, double* __restrict__ mom, int M) {
    __shared__ double first[3];
    int const n = blockIdx.x + 1;
    __shared__ double sums[2][3];
    for (int pass = 0; pass < 2; ++pass) {
        int const step = pass == 0 ? 1 : n;
        if (threadIdx.x < C) {
            double s = 0.0;
            for (int b = 0; b < grid; ++b) s += table[(static_cast<int64_t>(step - 1) * grid + b) * C + threadIdx.x];
            sums[pass][threadIdx.x] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        first[0] = 0.5 * sums[0][0]; first[1] = sums[0][1]; first[2] = C == 3 ? sums[0][2] : 0.0;
        if (n == 1) {
            mom[0] = first[0]; mom[1] = 0.0; mom[2] = first[1]; mom[3] = first[2];
        } else {
            int64_t const i0 = 2 * (n - 1);
            mom[i0 * 2] = 2.0 * (sums[1][0] - first[0]); mom[i0 * 2 + 1] = 0.0;
            mom[(i0 + 1) * 2] = 2.0 * sums[1][1] - first[1];
            mom[(i0 + 1) * 2 + 1] = C == 3 ? 2.0 * sums[1][2] - first[2] : 0.0;
        }
    }
    (void)steps; (void)M;
}

using PersistKernel = void (*)(PersistDev);

template<class T, int K>
PersistKernel persist_kernel_rpt(int rpt) {
    switch (rpt) {
        case 1: return cheb_persistent<T, K, 1>;
        case 2: return cheb_persistent<T, K, 2>;
        case 4: return cheb_persistent<T, K, 4>;
        default: return nullptr;
    }
}
template<class T>
PersistKernel persist_kernel(int k, int rpt) {
    switch (k) {
        case 3: return persist_kernel_rpt<T, 3>(rpt);
        case 4: return persist_kernel_rpt<T, 4>(rpt);
        case 7: return persist_kernel_rpt<T, 7>(rpt);
        default: return nullptr;
    }
}

template<class T>
cudaError_t launch_persist_t(PersistArgs const& a, int num_sms, cudaStream_t stream, bool* handled) {
    constexpr int C = ST<T>::C;
    *handled = false;
    int const max_grid = num_sms * 2;
    int rpt = 0;
    for (int cand : {1, 2, 4}) { if (static_cast<int64_t>(max_grid) * PERSIST_TPB * cand >= a.nrows) { rpt = cand; break; } }
    if (rpt == 0 || a.steps < 1) return cudaSuccess;
    PersistKernel const fn = persist_kernel<T>(a.h.k, rpt);
    if (!fn) return cudaSuccess;
    int per_sm = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PERSIST_TPB, 0);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) return cudaSuccess;
    int const resident = num_sms * (per_sm < 2 ? per_sm : 2);
    int grid = static_cast<int>((a.nrows + static_cast<int64_t>(PERSIST_TPB) * rpt - 1) / (static_cast<int64_t>(PERSIST_TPB) * rpt));
    if (grid > resident) return cudaSuccess;          // every CTA must be resident for the grid barrier
    if (grid > a.table_ctas) return cudaSuccess;
    PersistDev d{};
    d.val = a.h.val; d.col = a.h.col; d.pitch = a.h.pitch; d.k = a.h.k; d.nrows = static_cast<int>(a.nrows);
    d.buf[0] = a.buf0; d.buf[1] = a.buf1; d.steps = a.steps; d.table = a.table; d.barrier = a.barrier;
    err = cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), stream);
    if (err != cudaSuccess) return err;
    void* params[] = {&d};
    err = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(fn), dim3(grid), dim3(PERSIST_TPB), params, 0, stream);
    if (err != cudaSuccess) return err;
    persist_finish_kernel<C><<<a.steps, 32, 0, stream>>>(a.table, grid, a.steps, a.mom, a.M);
    *handled = true;
    return cudaGetLastError();
}

} // anonymous namespace

cudaError_t launch_persistent_diagonal(int dtype, PersistArgs const& a, int num_sms, cudaStream_t stream, bool* handled) {
    *handled = false;
    if (a.nrows <= 0 || a.nrows >= (int64_t{1} << 30)) return cudaSuccess;
    switch (dtype) {
        case F32: return launch_persist_t<float>(a, num_sms, stream, handled);
        case C64: return launch_persist_t<float2>(a, num_sms, stream, handled);
        case F64: return launch_persist_t<double>(a, num_sms, stream, handled);
        case C128: return launch_persist_t<double2>(a, num_sms, stream, handled);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
