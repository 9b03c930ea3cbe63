// step_common.cuh -- device helpers shared by the fused Chebyshev step kernels (kernels.cu, kernels_bulk.cu).
#pragma once
#include "kernels.cuh"

#include <cstdint>

namespace pbk {
namespace {

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
template<class T> struct ST;
template<> struct ST<float>   { using real = float;  static constexpr bool cplx = false; static constexpr int C = 2; };
template<> struct ST<double>  { using real = double; static constexpr bool cplx = false; static constexpr int C = 2; };
template<> struct ST<float2>  { using real = float;  static constexpr bool cplx = true;  static constexpr int C = 3; };
template<> struct ST<double2> { using real = double; static constexpr bool cplx = true;  static constexpr int C = 3; };

__device__ __forceinline__ float zero_(float) { return 0.f; }
__device__ __forceinline__ double zero_(double) { return 0.0; }
__device__ __forceinline__ float2 zero_(float2) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ double2 zero_(double2) { return make_double2(0.0, 0.0); }

__device__ __forceinline__ float neg_(float a) { return -a; }
__device__ __forceinline__ double neg_(double a) { return -a; }
__device__ __forceinline__ float2 neg_(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ double2 neg_(double2 a) { return make_double2(-a.x, -a.y); }

// acc + a * x  (complex: 4 FMAs, same association as compute::detail::mul + add)
__device__ __forceinline__ float fma_(float a, float x, float acc) { return fmaf(a, x, acc); }
__device__ __forceinline__ double fma_(double a, double x, double acc) { return fma(a, x, acc); }
__device__ __forceinline__ float2 fma_(float2 a, float2 x, float2 acc) {
    acc.x = fmaf(a.x, x.x, acc.x); acc.x = fmaf(-a.y, x.y, acc.x);
    acc.y = fmaf(a.x, x.y, acc.y); acc.y = fmaf(a.y, x.x, acc.y);
    return acc;
}
__device__ __forceinline__ double2 fma_(double2 a, double2 x, double2 acc) {
    acc.x = fma(a.x, x.x, acc.x); acc.x = fma(-a.y, x.y, acc.x);
    acc.y = fma(a.x, x.y, acc.y); acc.y = fma(a.y, x.x, acc.y);
    return acc;
}
__device__ __forceinline__ float scale_(float a, double s) { return a * static_cast<float>(s); }
__device__ __forceinline__ double scale_(double a, double s) { return a * s; }
__device__ __forceinline__ float2 scale_(float2 a, double s) { float f = static_cast<float>(s); return make_float2(a.x * f, a.y * f); }
__device__ __forceinline__ double2 scale_(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

// acc[0] += |x|^2 ; acc[1] (+ acc[2]) += conj(y) * x   -- in double
__device__ __forceinline__ void sums_(double* acc, float x, float y) {
    double const xd = x, yd = y; acc[0] = fma(xd, xd, acc[0]); acc[1] = fma(yd, xd, acc[1]);
}
__device__ __forceinline__ void sums_(double* acc, double x, double y) { acc[0] = fma(x, x, acc[0]); acc[1] = fma(y, x, acc[1]); }
__device__ __forceinline__ void sums_(double* acc, float2 x, float2 y) {
    double const xr = x.x, xi = x.y, yr = y.x, yi = y.y;
    acc[0] = fma(xr, xr, acc[0]); acc[0] = fma(xi, xi, acc[0]);
    acc[1] = fma(yr, xr, acc[1]); acc[1] = fma(yi, xi, acc[1]);
    acc[2] = fma(yr, xi, acc[2]); acc[2] = fma(-yi, xr, acc[2]);
}
__device__ __forceinline__ void sums_(double* acc, double2 x, double2 y) {
    acc[0] = fma(x.x, x.x, acc[0]); acc[0] = fma(x.y, x.y, acc[0]);
    acc[1] = fma(y.x, x.x, acc[1]); acc[1] = fma(y.y, x.y, acc[1]);
    acc[2] = fma(y.x, x.y, acc[2]); acc[2] = fma(-y.y, x.x, acc[2]);
}

/// V elements of T moved as one vector access (16 bytes when V * sizeof(T) == 16)
template<class T, int V> struct alignas(V * sizeof(T)) Chunk { T e[V]; };

template<class CH> __device__ __forceinline__ CH load_nc(const CH* p) {  // read-only path (ld.global.nc)
    CH r;
    if constexpr (sizeof(CH) == 16) { int4 t = __ldg(reinterpret_cast<const int4*>(p)); r = *reinterpret_cast<CH*>(&t); }
    else if constexpr (sizeof(CH) == 8) { int2 t = __ldg(reinterpret_cast<const int2*>(p)); r = *reinterpret_cast<CH*>(&t); }
    else { int t = __ldg(reinterpret_cast<const int*>(p)); r = *reinterpret_cast<CH*>(&t); }
    return r;
}
template<class CH> __device__ __forceinline__ CH load_cs(const CH* p) {  // streaming: read once
    CH r;
    if constexpr (sizeof(CH) == 16) { int4 t = __ldcs(reinterpret_cast<const int4*>(p)); r = *reinterpret_cast<CH*>(&t); }
    else if constexpr (sizeof(CH) == 8) { int2 t = __ldcs(reinterpret_cast<const int2*>(p)); r = *reinterpret_cast<CH*>(&t); }
    else { int t = __ldcs(reinterpret_cast<const int*>(p)); r = *reinterpret_cast<CH*>(&t); }
    return r;
}
template<class CH> __device__ __forceinline__ void store_(CH* p, CH const& v) {
    if constexpr (sizeof(CH) == 16) { *reinterpret_cast<int4*>(p) = *reinterpret_cast<const int4*>(&v); }
    else if constexpr (sizeof(CH) == 8) { *reinterpret_cast<int2*>(p) = *reinterpret_cast<const int2*>(&v); }
    else { *reinterpret_cast<int*>(p) = *reinterpret_cast<const int*>(&v); }
}
template<class CH> __device__ __forceinline__ void store_cs(CH* p, CH const& v) {  // streaming store: not re-read by this launch
    if constexpr (sizeof(CH) == 16) { __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&v)); }
    else if constexpr (sizeof(CH) == 8) { __stcs(reinterpret_cast<int2*>(p), *reinterpret_cast<const int2*>(&v)); }
    else { __stcs(reinterpret_cast<int*>(p), *reinterpret_cast<const int*>(&v)); }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template<class T> __device__ __forceinline__ T ldg_scalar(const T* p) {
    Chunk<T, 1> c = load_nc(reinterpret_cast<const Chunk<T, 1>*>(p));
    return c.e[0];
}

struct StepDev {  // POD copy of StepArgs for the kernel
    const void* val; const int32_t* col; int64_t pitch; int k;
    const void* x; void* y; void* y2;
    int64_t nrows; int R; int cpr; int rpb;
    int ipt; int64_t tile_jump;   // block-iterations per tile; rows to skip to reach this block's next tile
    int pf;                       // L2 prefetch distance of the streamed operands, in block-iterations (0: off)
    int pfmask;                   // only chunks with (tx & pfmask) == 0 issue the prefetch (one request per line)
    double scale;
    double* partials; unsigned* counter; double* mom; double* m01; int M; int n; int fin;
    // destinations inside a k-blocked Kubo-Bastin stack (kubo.cu): byte stride between consecutive 256-byte blocks of the
    // vector, 0 = an ordinary contiguous N x R block
    int64_t y_bs; int64_t y2_bs;
};

/// address of chunk `ci` of a vector whose consecutive 256-byte blocks lie `bs` bytes apart (row m of a k-blocked stack)
template<class CH> __device__ __forceinline__ CH* blocked_dst(void* base, int64_t ci, int64_t bs) {
    static_assert(256 % sizeof(CH) == 0, "chunks never straddle a block");
    int64_t const o = ci * static_cast<int64_t>(sizeof(CH));
    return reinterpret_cast<CH*>(static_cast<char*>(base) + (o >> 8) * bs + (o & 255));
}

// ------------------------------------------------------------------------------------------------
// Tail of the fused step kernels: block reduction of the per-thread sums (fixed-order tree over the rows of
// the block, one value at a time through a TPB-sized buffer -- a small static footprint keeps the SM's L1
// carve-out large), per-block partials, and the last block to finish reduces over blocks in a fixed order
// and writes the moments (Diagonal / BatchDiagonal collectors, src/kpm/default/collectors.cpp:6-34).
// ------------------------------------------------------------------------------------------------
template<int C, int NACC, int TPB>
__device__ __forceinline__ void finish_sums(StepDev const& a, double (&acc)[NACC], int tx, int ty) {
    __shared__ double sm[TPB];
    __shared__ bool is_last;
    int const tid = threadIdx.x;
    int const cpr = a.cpr, rpb = a.rpb;
    int p2 = 1;
    while (p2 < rpb) p2 <<= 1;
    int const RC = a.R * C;
    if (cpr >= 32 || 32 % cpr == 0) {
        // rows that share a warp are folded with warp shuffles (fixed order), then thread tx adds the per-warp (per-row)
        // partials of its column in warp order: two CTA barriers per accumulator instead of log2(rows) + 2
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            double v = acc[q];                                     // threads without a row hold zeros
            if (cpr < 32) { for (int off = 16; off >= cpr; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off); }
            sm[tid] = v;
            __syncthreads();
            if (tid < cpr) {
                double s = 0.0;
                if (cpr >= 32) { for (int r = 0; r < rpb; ++r) s += sm[tid + r * cpr]; }
                else { for (int w = 0; w < TPB / 32; ++w) s += sm[w * 32 + tid]; }   // lane tx of every warp holds the warp's partial
                a.partials[static_cast<int64_t>(blockIdx.x) * RC + tid * NACC + q] = s;
            }
            __syncthreads();
        }
    } else {
        // chunk counts that do not divide a warp (R = 12, 20, ...): shared-memory tree over the rows
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            sm[tid] = acc[q];
            for (int s = p2 >> 1; s > 0; s >>= 1) {
                __syncthreads();
                if (ty < s && ty + s < rpb) sm[tid] += sm[tid + s * cpr];
            }
            __syncthreads();
            if (ty == 0) a.partials[static_cast<int64_t>(blockIdx.x) * RC + tx * NACC + q] = sm[tx];
            __syncthreads();
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) { is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1); }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- last block: reduce over blocks in a fixed order and write the moments ----
    int const nt = blockDim.x;
    int const nb = gridDim.x;
    auto finalize = [&](int o, double s) {
        int const lane = o / C, comp = o % C;
        if (a.fin == FIN_INIT) {
            if (comp == 0) { double m0 = 0.5 * s; a.m01[lane * 3 + 0] = m0; a.mom[(static_cast<int64_t>(lane) * a.M + 0) * 2] = m0; a.mom[(static_cast<int64_t>(lane) * a.M + 0) * 2 + 1] = 0.0; }
            else if (comp == 1) { a.m01[lane * 3 + 1] = s; a.mom[(static_cast<int64_t>(lane) * a.M + 1) * 2] = s; if (C == 2) { a.m01[lane * 3 + 2] = 0.0; a.mom[(static_cast<int64_t>(lane) * a.M + 1) * 2 + 1] = 0.0; } }
            else { a.m01[lane * 3 + 2] = s; a.mom[(static_cast<int64_t>(lane) * a.M + 1) * 2 + 1] = s; }
        } else if (a.fin == FIN_STEP) {
            int64_t const i0 = static_cast<int64_t>(lane) * a.M + 2 * (a.n - 1);
            if (comp == 0) { a.mom[i0 * 2] = 2.0 * (s - a.m01[lane * 3 + 0]); a.mom[i0 * 2 + 1] = 0.0; }
            else if (comp == 1) { a.mom[(i0 + 1) * 2] = 2.0 * s - a.m01[lane * 3 + 1]; if (C == 2) a.mom[(i0 + 1) * 2 + 1] = 0.0; }
            else { a.mom[(i0 + 1) * 2 + 1] = 2.0 * s - a.m01[lane * 3 + 2]; }
        }
    };
    if (RC >= nt) {
        for (int o = tid; o < RC; o += nt) {
            double s = 0.0;
            for (int b = 0; b < nb; ++b) s += a.partials[static_cast<int64_t>(b) * RC + o];
            finalize(o, s);
        }
    } else {
        int const G = nt / RC;
        int const g = tid / RC, o = tid % RC;
        double s = 0.0;
        if (g < G) { for (int b = g; b < nb; b += G) s += a.partials[static_cast<int64_t>(b) * RC + o]; }
        __syncthreads();
        if (g < G) sm[g * RC + o] = s;
        __syncthreads();
        if (tid < RC) {
            double t = 0.0;
            for (int gg = 0; gg < G; ++gg) t += sm[gg * RC + tid];
            finalize(tid, t);
        }
    }
    if (tid == 0) *a.counter = 0u;
}


using StepKernel = void (*)(StepDev);

} // anonymous namespace
} // namespace pbk
