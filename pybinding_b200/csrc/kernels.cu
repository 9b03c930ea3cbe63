// kernels.cu -- hand-written sm_100a kernels of the KPM hot path.
//
// K1/K2/K3 `cheb_step`: one fused Chebyshev step for a block of R vectors,
//     y[row, :] = sum_s val[s][row] * x[col[s][row], :]  -  y[row, :]          (rows < nrows)
// optionally fused with the two per-vector reductions of the diagonal algorithm
//     m2[r] = sum_row |x[row, r]|^2 ,   m3[r] = sum_row conj(y_new[row, r]) * x[row, r]
// and with the moment bookkeeping (the last block to finish reduces the per-block partial sums in
// a fixed order and writes mu_{2(n-1)} = 2 (m2 - m0), mu_{2n-1} = 2 m3 - m1 straight into the device
// moment array), so one launch per step is all a diagonal KPM run needs.
//
// Replaces, from the reference (cppcore/): compute::kpm_spmv / kpm_spmv_diagonal
// (include/compute/kernel_polynomial.hpp:19-326), make_r1 (include/kpm/Starter.hpp:55-118), the
// Diagonal/BatchDiagonal collectors (src/kpm/default/collectors.cpp:6-34) and the thread-pool
// batching of DefaultCompute (src/kpm/default/Compute.cpp:52-88).
//
// Layout: H in slot-major ELL (val[s * pitch + row], col[s * pitch + row]); vectors as an N x R
// row-major block so one matrix element feeds R contiguous lanes.  Each thread owns one 16-byte
// chunk (V lanes) of one row: a warp touches 512 contiguous bytes of y / x[row] and gathers whole
// 16-byte chunks of x[col] -- every access is a full-sector vector load.  The (col, val) pair of a
// thread's *next* row is prefetched while the gathers of the current row are in flight, which
// removes the index -> gather dependency from the critical path.  Dot products are accumulated in
// double regardless of the vector type.
#include "kernels.cuh"

#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "step_common.cuh"

namespace pbk {

namespace {

// ------------------------------------------------------------------------------------------------
// The fused step kernel.  K > 0: ELL width known at compile time (unrolled + prefetched); K == 0: generic.
//
// Row -> CTA mapping: the rows are cut into tiles of ipt * rpb consecutive rows and tile t belongs to CTA
// t mod gridDim.x, which walks through it rpb rows at a time.  With the engine's locality ordering a tile is
// a breadth-first ball of the lattice, so the x-rows gathered by a CTA are re-used from the SM's L1 (each
// x-row is needed by its own row and by every neighbour) instead of being re-read from L2 by other SMs.
// ipt == 1 degenerates to a plain grid-stride loop (used for the light-cone-sliced runs).
// ------------------------------------------------------------------------------------------------
template<class T, int V, int K, bool SUB, bool SUMS, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) cheb_step(StepDev a) {
    using CH = Chunk<T, V>;
    constexpr int C = ST<T>::C;
    constexpr int NACC = V * C;
    constexpr int KK = K > 0 ? K : 1;

    const T* __restrict__ val = static_cast<const T*>(a.val);
    const int32_t* __restrict__ col = a.col;
    const CH* __restrict__ x = static_cast<const CH*>(a.x);
    CH* __restrict__ y = static_cast<CH*>(a.y);
    CH* __restrict__ y2 = static_cast<CH*>(a.y2);

    int const cpr = a.cpr;                       // 16-byte chunks per row
    int const tx = threadIdx.x % cpr;            // chunk within the row
    int const ty = threadIdx.x / cpr;            // row within the block
    int const rpb = a.rpb;
    int const ipt = a.ipt;
    int64_t const tile_jump = a.tile_jump;
    int64_t row = static_cast<int64_t>(blockIdx.x) * ipt * rpb + ty;
    int w = 0;                                   // block-iteration within the current tile
    auto advance = [&](int64_t r) {
        r += rpb;
        if (++w == ipt) { w = 0; r += tile_jump; }
        return r;
    };

    // L2 prefetch cursor: runs `pf` block-iterations ahead of `row` along the same row sequence.  The streamed
    // operands (y and the tile's own x rows) are then L2 hits when the demand loads arrive, so the DRAM
    // queue depth no longer depends on how many loads the register file can keep in flight.
    int64_t prow = row;
    int pw = 0;
    auto advance_pf = [&](int64_t r) {
        r += rpb;
        if (++pw == ipt) { pw = 0; r += tile_jump; }
        return r;
    };
    bool const pf_lane = (tx & a.pfmask) == 0;
    if (a.pf > 0) {
        for (int i = 0; i < a.pf; ++i) {
            if (prow < a.nrows && pf_lane) {
                prefetch_l2(x + prow * cpr + tx);
                if constexpr (SUB) prefetch_l2(y + prow * cpr + tx);
            }
            prow = advance_pf(prow);
        }
    }

    double acc[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = 0.0;

    // y and the optional second destination; either may be a row of a k-blocked Kubo-Bastin stack (launch-uniform)
    int64_t const y_bs = a.y_bs, y2_bs = a.y2_bs;
    auto store_out = [&](int64_t ci, CH const& out) {
        if (y_bs == 0) store_cs(y + ci, out); else store_cs(blocked_dst<CH>(a.y, ci, y_bs), out);
        if (y2) { if (y2_bs == 0) store_cs(y2 + ci, out); else store_cs(blocked_dst<CH>(a.y2, ci, y2_bs), out); }
    };

    if constexpr (K > 0) {
        int32_t c[KK]; T v[KK];
        if (row < a.nrows) {
#pragma unroll
            for (int s = 0; s < KK; ++s) { c[s] = __ldg(col + s * a.pitch + row); v[s] = ldg_scalar(val + s * a.pitch + row); }
        }
        while (row < a.nrows) {
            int64_t const next = advance(row);
            if (a.pf > 0) {
                if (prow < a.nrows && pf_lane) {
                    prefetch_l2(x + prow * cpr + tx);
                    if constexpr (SUB) prefetch_l2(y + prow * cpr + tx);
                }
                prow = advance_pf(prow);
            }
            // gathers of the current row (independent loads, all in flight together)
            CH xg[KK];
#pragma unroll
            for (int s = 0; s < KK; ++s) xg[s] = load_nc(x + static_cast<int64_t>(c[s]) * cpr + tx);
            CH yv, xr;
            if constexpr (SUB) yv = load_cs(y + row * cpr + tx);
            if constexpr (SUMS) xr = load_nc(x + row * cpr + tx);
            // prefetch the next row's matrix entries while the gathers fly
            int32_t cn[KK]; T vn[KK];
            if (next < a.nrows) {
#pragma unroll
                for (int s = 0; s < KK; ++s) { cn[s] = __ldg(col + s * a.pitch + next); vn[s] = ldg_scalar(val + s * a.pitch + next); }
            }
            CH out;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                T r = SUB ? neg_(yv.e[e]) : zero_(T{});
#pragma unroll
                for (int s = 0; s < KK; ++s) r = fma_(v[s], xg[s].e[e], r);
                if constexpr (!SUB) r = scale_(r, a.scale);
                out.e[e] = r;
                if constexpr (SUMS) sums_(acc + e * C, xr.e[e], r);
            }
            store_out(row * cpr + tx, out);
#pragma unroll
            for (int s = 0; s < KK; ++s) { c[s] = cn[s]; v[s] = vn[s]; }
            row = next;
        }
    } else {
        for (; row < a.nrows; row = advance(row)) {
            CH out;
#pragma unroll
            for (int e = 0; e < V; ++e) out.e[e] = zero_(T{});
            if constexpr (SUB) {
                CH yv = load_cs(y + row * cpr + tx);
#pragma unroll
                for (int e = 0; e < V; ++e) out.e[e] = neg_(yv.e[e]);
            }
            for (int s = 0; s < a.k; ++s) {
                int32_t const cc = __ldg(col + s * a.pitch + row);
                T const vv = ldg_scalar(val + s * a.pitch + row);
                CH const xg = load_nc(x + static_cast<int64_t>(cc) * cpr + tx);
#pragma unroll
                for (int e = 0; e < V; ++e) out.e[e] = fma_(vv, xg.e[e], out.e[e]);
            }
            if constexpr (!SUB) {
#pragma unroll
                for (int e = 0; e < V; ++e) out.e[e] = scale_(out.e[e], a.scale);
            }
            if constexpr (SUMS) {
                CH const xr = load_nc(x + row * cpr + tx);
#pragma unroll
                for (int e = 0; e < V; ++e) sums_(acc + e * C, xr.e[e], out.e[e]);
            }
            store_out(row * cpr + tx, out);
        }
    }

    if constexpr (SUMS) finish_sums<C, NACC, TPB>(a, acc, tx, ty);
}


template<class T, int V, int K, int TPB, int MINB>
StepKernel step_kernel_vk(bool sub, bool sums) {
    if (sub && sums) return cheb_step<T, V, K, true, true, TPB, MINB>;
    if (sub) return cheb_step<T, V, K, true, false, TPB, MINB>;
    if (sums) return cheb_step<T, V, K, false, true, TPB, MINB>;
    return cheb_step<T, V, K, false, false, TPB, MINB>;
}

template<class T, int V, int TPB, int MINB>
StepKernel step_kernel_v(int k, bool sub, bool sums, int* kused) {
    switch (k) {
        case 3: *kused = 3; return step_kernel_vk<T, V, 3, TPB, MINB>(sub, sums);
        case 4: *kused = 4; return step_kernel_vk<T, V, 4, TPB, MINB>(sub, sums);
        case 7: *kused = 7; return step_kernel_vk<T, V, 7, TPB, MINB>(sub, sums);
        default: *kused = 0; return step_kernel_vk<T, V, 0, TPB, MINB>(sub, sums);
    }
}

/// resident CTAs per SM of a kernel variant (queried once per variant and block size)
int resident_blocks(StepKernel fn, int block, int dyn_smem = 0) {
    static std::mutex mutex;
    static std::map<std::tuple<StepKernel, int, int>, int> cache;
    std::lock_guard<std::mutex> lock(mutex);
    auto const key = std::make_tuple(fn, block, dyn_smem);
    auto const it = cache.find(key);
    if (it != cache.end()) return it->second;
    if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, block, dyn_smem) != cudaSuccess || nb < 1) nb = 1;
    cache[key] = nb;
    return nb;
}

template<class T, int TPB, int MINB>
cudaError_t launch_step_t(StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info) {
    constexpr int VMAX = 16 / sizeof(T);
    int V = VMAX;
    while (V > 1 && a.R % V != 0) V >>= 1;
    if (VMAX == 4 && V == 2) V = 1;  // only the full-width and the scalar variants are instantiated
    int const cpr = a.R / V;
    if (cpr > TPB) return cudaErrorInvalidValue;
    int const rpb = TPB / cpr;
    int const block = rpb * cpr;
    int kused = 0;
    StepKernel const fn = (V == VMAX && VMAX > 1) ? step_kernel_v<T, VMAX, TPB, MINB>(a.h.k, a.subtract, a.sums, &kused)
                                                  : step_kernel_v<T, 1, TPB, MINB>(a.h.k, a.subtract, a.sums, &kused);
    int ipt = 1;
    if (a.tile > rpb) ipt = static_cast<int>((a.tile + rpb - 1) / rpb);
    int64_t const tile_rows = static_cast<int64_t>(ipt) * rpb;
    int64_t const need = (a.nrows + tile_rows - 1) / tile_rows;

    // persistent grid: exactly the CTAs that are resident at once, so all of them sweep the rows together
    int const cap = num_sms * (a.blocks_per_sm > 0 ? a.blocks_per_sm : resident_blocks(fn, block));
    int grid = static_cast<int>(need < static_cast<int64_t>(cap) ? need : cap);
    if (grid < 1) grid = 1;
    if (grid > max_step_blocks(num_sms)) grid = max_step_blocks(num_sms);

    StepDev d{a.h.val, a.h.col, a.h.pitch, a.h.k, a.x, a.y, a.y2, a.nrows, a.R, cpr, rpb,
              ipt, static_cast<int64_t>(grid - 1) * tile_rows, a.prefetch, a.prefetch_mask, a.scale,
              a.partials, a.counter, a.mom, a.m01, a.M, a.n, a.fin, a.y_block_stride, a.y2_block_stride};
    if (a.y_block_stride != 0 && a.subtract) return cudaErrorInvalidValue;   // a blocked y is write-only
    fn<<<grid, block, 0, stream>>>(d);
    if (info) { info->grid = grid; info->block = block; info->V = V; info->K = kused; info->bulk = 0; }
    return cudaGetLastError();
}

template<class T>
cudaError_t launch_step_tpb(StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info) {
    switch (a.tpb) {
        case 512: return launch_step_t<T, 512, 2>(a, num_sms, stream, info);
        case 1024: return launch_step_t<T, 1024, 1>(a, num_sms, stream, info);
        case 2565: return launch_step_t<T, 256, 5>(a, num_sms, stream, info);  // experiments: 256 threads, >= 5/6/8 CTAs per SM
        case 2566: return launch_step_t<T, 256, 6>(a, num_sms, stream, info);
        case 2568: return launch_step_t<T, 256, 8>(a, num_sms, stream, info);
        case 2563: return launch_step_t<T, 256, 3>(a, num_sms, stream, info);
        default: return launch_step_t<T, 256, 4>(a, num_sms, stream, info);
    }
}

} // anonymous namespace

int max_step_blocks(int num_sms) { return num_sms * 8; }

cudaError_t launch_step(int dtype, StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info) {
    if (a.nrows <= 0) return cudaSuccess;
    if (a.bulk_stages >= 2 && a.packed) {  // main variant where it applies (kernels_bulk.cu)
        bool handled = false;
        cudaError_t const err = launch_step_bulk(dtype, a, num_sms, stream, info, &handled);
        if (handled || err != cudaSuccess) return err;
    }
    switch (dtype) {
        case F32: return launch_step_tpb<float>(a, num_sms, stream, info);
        case C64: return launch_step_tpb<float2>(a, num_sms, stream, info);
        case F64: return launch_step_tpb<double>(a, num_sms, stream, info);
        case C128: return launch_step_tpb<double2>(a, num_sms, stream, info);
        default: return cudaErrorInvalidValue;
    }
}

// ================================================================================================
// small helper kernels
// ================================================================================================
namespace {

template<class T> __device__ __forceinline__ T from_c128(double re, double im);
template<> __device__ __forceinline__ float from_c128<float>(double re, double) { return static_cast<float>(re); }
template<> __device__ __forceinline__ double from_c128<double>(double re, double) { return re; }
template<> __device__ __forceinline__ float2 from_c128<float2>(double re, double im) { return make_float2(static_cast<float>(re), static_cast<float>(im)); }
template<> __device__ __forceinline__ double2 from_c128<double2>(double re, double im) { return make_double2(re, im); }
__device__ __forceinline__ double re_(float a) { return a; }
__device__ __forceinline__ double re_(double a) { return a; }
__device__ __forceinline__ double re_(float2 a) { return a.x; }
__device__ __forceinline__ double re_(double2 a) { return a.x; }
__device__ __forceinline__ double im_(float) { return 0.0; }
__device__ __forceinline__ double im_(double) { return 0.0; }
__device__ __forceinline__ double im_(float2 a) { return a.y; }
__device__ __forceinline__ double im_(double2 a) { return a.y; }

template<class T>
__global__ void unit_starter_kernel(T* dst, int R, const int32_t* src, int nsrc) {
    int const lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane < nsrc) dst[static_cast<int64_t>(src[lane]) * R + lane] = from_c128<T>(1.0, 0.0);
}

template<class T>
__global__ void gather_moment_kernel(const T* v, int R, const int32_t* idx, int nidx, double* mom, int64_t M, int n, double scale) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nidx) return;
    T const e = v[static_cast<int64_t>(idx[i]) * R];
    mom[(static_cast<int64_t>(i) * M + n) * 2] = re_(e) * scale;
    mom[(static_cast<int64_t>(i) * M + n) * 2 + 1] = im_(e) * scale;
}

/// Deterministic grid reduction of NV doubles per thread: block tree, then the last block sums the partials in order.
template<int NV, class Fin>
__device__ void grid_reduce(double (&v)[NV], double* scratch, unsigned* counter, Fin fin) {
    __shared__ double sm[256 * NV];
    __shared__ bool is_last;
    int const tid = threadIdx.x;
#pragma unroll
    for (int q = 0; q < NV; ++q) sm[tid * NV + q] = v[q];
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        __syncthreads();
        if (tid < s) {
#pragma unroll
            for (int q = 0; q < NV; ++q) sm[tid * NV + q] += sm[(tid + s) * NV + q];
        }
    }
    __syncthreads();
    if (tid < NV) scratch[static_cast<int64_t>(blockIdx.x) * NV + tid] = sm[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) t[q] = 0.0;
    for (int b = tid; b < static_cast<int>(gridDim.x); b += blockDim.x) {
#pragma unroll
        for (int q = 0; q < NV; ++q) t[q] += scratch[static_cast<int64_t>(b) * NV + q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) sm[tid * NV + q] = t[q];
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        __syncthreads();
        if (tid < s) {
#pragma unroll
            for (int q = 0; q < NV; ++q) sm[tid * NV + q] += sm[(tid + s) * NV + q];
        }
    }
    __syncthreads();
    if (tid == 0) { fin(sm); *counter = 0u; }
}

template<class T>
__global__ void __launch_bounds__(256) dot_moment_kernel(const T* beta, const T* v, int64_t n_rows, double* mom, int n, double scale,
                                                         double* scratch, unsigned* counter) {
    double acc[2] = {0.0, 0.0};
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_rows; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double const br = re_(beta[i]), bi = im_(beta[i]), vr = re_(v[i]), vi = im_(v[i]);
        acc[0] += br * vr + bi * vi;   // conj(beta) * v
        acc[1] += br * vi - bi * vr;
    }
    grid_reduce<2>(acc, scratch, counter, [&](double* s) { mom[2 * n] = s[0] * scale; mom[2 * n + 1] = s[1] * scale; });
}

__global__ void accumulate_lanes_kernel(const double* mom, int R, int M, double* acc) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;  // over 2*M doubles
    if (i >= 2 * M) return;
    double s = 0.0;
    for (int lane = 0; lane < R; ++lane) s += mom[static_cast<int64_t>(lane) * M * 2 + i];
    acc[i] += s;
}

template<class T>
__global__ void scatter_block_kernel(const double* src, int64_t n, int R, int lane, const int32_t* perm, T* dst) {
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t const row = perm ? perm[i] : i;
    dst[row * R + lane] = from_c128<T>(src[2 * i], src[2 * i + 1]);
}

template<class T>
__global__ void extract_lane_kernel(const T* v, int64_t n, int R, int lane, const int32_t* perm, double* out) {
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t const row = perm ? perm[i] : i;
    T const e = v[row * R + lane];
    out[2 * i] = re_(e); out[2 * i + 1] = im_(e);
}

// ---- light-cone sub-systems ("cones") of the unit-vector quantities ---------------------------
/// gmap[row of site queue[i] in the resident layout] = set ? i : -1
__global__ void cone_mark_kernel(const int32_t* __restrict__ queue, int64_t count, const int32_t* __restrict__ perm, int32_t* gmap, int set) {
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int32_t const site = queue[i];
    gmap[perm ? perm[site] : site] = set ? static_cast<int32_t>(i) : -1;
}

/// rows [0, rows) of the sub-system: the resident ELL row of site queue[i] with its columns relabelled through gmap.
/// Columns outside the cone cannot occur for rows whose whole neighbourhood was visited; they are neutralised anyway.
template<class T>
__global__ void cone_extract_kernel(const T* __restrict__ val, const int32_t* __restrict__ col, int64_t pitch, int k,
                                    const int32_t* __restrict__ queue, const int32_t* __restrict__ perm, const int32_t* __restrict__ gmap,
                                    int64_t rows, T* __restrict__ out_val, int32_t* __restrict__ out_col, int64_t out_pitch) {
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= out_pitch) return;
    if (i >= rows) {
        for (int s = 0; s < k; ++s) { out_val[s * out_pitch + i] = zero_(T{}); out_col[s * out_pitch + i] = 0; }
        return;
    }
    int32_t const site = queue[i];
    int64_t const g = perm ? perm[site] : site;
    for (int s = 0; s < k; ++s) {
        T v = val[s * pitch + g];
        int32_t lc = gmap[col[s * pitch + g]];
        if (lc < 0) { lc = static_cast<int32_t>(i); v = zero_(T{}); }
        out_val[s * out_pitch + i] = v;
        out_col[s * out_pitch + i] = lc;
    }
}

/// One Chebyshev step of the diagonal recursion for a GROUP of independent light-cone sub-systems (one unit vector
/// each): blockIdx.y selects the sub-system, blockIdx.x strides over its rows of this step.  Step k reads r_{k-1} from
/// buf[(k-1)&1], overwrites r_{k-2} in buf[k&1] with r_k, and the last block of each sub-system writes its two moments
/// (same arithmetic and reductions as `cheb_step` with R = 1).  k == 1 is the initial step r_1 = H~ r_0 / 2.
template<class T, int K>   // K > 0: ELL width known at compile time (all index / value / gather loads of a row in flight together)
__global__ void __launch_bounds__(256) cone_group_step_kernel(const ConeSlot* __restrict__ slots, int k, int kell, int M,
                                                              double* partials, unsigned* counters) {
    constexpr int C = ST<T>::C;
    ConeSlot const sl = slots[blockIdx.y];
    const T* __restrict__ val = static_cast<const T*>(sl.val);
    const int32_t* __restrict__ col = sl.col;
    const T* __restrict__ x = static_cast<const T*>(sl.buf[(k - 1) & 1]);
    T* __restrict__ y = static_cast<T*>(sl.buf[k & 1]);
    int64_t const nrows = sl.rows[k];
    bool const init = k == 1;
    double acc[C];
#pragma unroll
    for (int q = 0; q < C; ++q) acc[q] = 0.0;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < nrows; row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        T r = init ? zero_(T{}) : neg_(y[row]);
        if constexpr (K > 0) {
            int32_t c[K]; T v[K], xg[K];
#pragma unroll
            for (int s = 0; s < K; ++s) { c[s] = __ldg(col + s * sl.pitch + row); v[s] = ldg_scalar(val + s * sl.pitch + row); }
#pragma unroll
            for (int s = 0; s < K; ++s) xg[s] = ldg_scalar(x + c[s]);
#pragma unroll
            for (int s = 0; s < K; ++s) r = fma_(v[s], xg[s], r);
        } else {
            for (int s = 0; s < kell; ++s) {
                int32_t const c = __ldg(col + s * sl.pitch + row);
                T const v = ldg_scalar(val + s * sl.pitch + row);
                r = fma_(v, ldg_scalar(x + c), r);
            }
        }
        if (init) r = scale_(r, 0.5);
        sums_(acc, ldg_scalar(x + row), r);
        y[row] = r;
    }
    StepDev fin{};
    fin.R = 1; fin.cpr = 1; fin.rpb = 256;
    fin.partials = partials + static_cast<int64_t>(blockIdx.y) * gridDim.x * C;
    fin.counter = counters + blockIdx.y;
    fin.mom = sl.mom; fin.m01 = sl.m01; fin.M = M; fin.n = k; fin.fin = init ? FIN_INIT : FIN_STEP;
    finish_sums<C, C, 256>(fin, acc, 0, static_cast<int>(threadIdx.x));
}

/// buf0 = unit vector at position 0, buf1 = 0 for every sub-system of a group
template<class T>
__global__ void cone_group_start_kernel(const ConeSlot* __restrict__ slots) {
    ConeSlot const sl = slots[blockIdx.y];
    T* b0 = static_cast<T*>(sl.buf[0]);
    T* b1 = static_cast<T*>(sl.buf[1]);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < sl.nvec; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        b0[i] = i == 0 ? from_c128<T>(1.0, 0.0) : zero_(T{});
        b1[i] = zero_(T{});
    }
}

// ---- Lanczos -----------------------------------------------------------------------------------
template<class T> __device__ __forceinline__ T axpy_(double a, T x, T y);  // y - a*x
template<> __device__ __forceinline__ float axpy_(double a, float x, float y) { return y - static_cast<float>(a) * x; }
template<> __device__ __forceinline__ double axpy_(double a, double x, double y) { return y - a * x; }
template<> __device__ __forceinline__ float2 axpy_(double a, float2 x, float2 y) { float f = static_cast<float>(a); return make_float2(y.x - f * x.x, y.y - f * x.y); }
template<> __device__ __forceinline__ double2 axpy_(double a, double2 x, double2 y) { return make_double2(y.x - a * x.x, y.y - a * x.y); }

/// v0 = t - b_prev*v0 - a*v1 ; out = |v0|^2     (lanczos_spmv + lanczos_axpy of compute/lanczos.hpp:26-98)
template<class T>
__global__ void __launch_bounds__(256) lanczos_update_kernel(const T* t, const T* v1, T* v0, int64_t n, double b_prev, const double* a_dev,
                                                             double* out, double* scratch, unsigned* counter) {
    double const a = a_dev[0];
    double acc[1] = {0.0};
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        T w = axpy_<T>(b_prev, v0[i], t[i]);
        w = axpy_<T>(a, v1[i], w);
        v0[i] = w;
        acc[0] += re_(w) * re_(w) + im_(w) * im_(w);
    }
    grid_reduce<1>(acc, scratch, counter, [&](double* s) { out[0] = s[0]; });
}

template<class T>
__global__ void scale_inv_sqrt_kernel(T* v, int64_t n, const double* norm2) {
    double const f = 1.0 / sqrt(norm2[0]);
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) v[i] = scale_(v[i], f);
}

inline int blocks_for(int64_t n, int tpb, int cap) {
    int64_t b = (n + tpb - 1) / tpb;
    if (b < 1) b = 1;
    return static_cast<int>(b < cap ? b : cap);
}

} // anonymous namespace

#define PBK_DISPATCH(dtype, CALL)                                   \
    switch (dtype) {                                                \
        case F32: { using T = float; CALL; break; }                 \
        case C64: { using T = float2; CALL; break; }                \
        case F64: { using T = double; CALL; break; }                \
        case C128: { using T = double2; CALL; break; }              \
        default: return cudaErrorInvalidValue;                      \
    }

cudaError_t launch_unit_starter(int dtype, void* dst, int64_t n, int R, const int32_t* src_dev, int nsrc, cudaStream_t s) {
    cudaError_t err = cudaMemsetAsync(dst, 0, static_cast<size_t>(n) * R * dtype_size(dtype), s);
    if (err != cudaSuccess || nsrc == 0) return err;
    PBK_DISPATCH(dtype, (unit_starter_kernel<T><<<(nsrc + 127) / 128, 128, 0, s>>>(static_cast<T*>(dst), R, src_dev, nsrc)));
    return cudaGetLastError();
}

cudaError_t launch_gather_moment(int dtype, const void* v, int R, const int32_t* idx_dev, int nidx, double* mom, int64_t M, int n,
                                 double scale, cudaStream_t s) {
    PBK_DISPATCH(dtype, (gather_moment_kernel<T><<<(nidx + 127) / 128, 128, 0, s>>>(static_cast<const T*>(v), R, idx_dev, nidx, mom, M, n, scale)));
    return cudaGetLastError();
}

cudaError_t launch_dot_moment(int dtype, const void* beta, const void* v, int64_t n_rows, double* mom, int n, double scale,
                              double* scratch, unsigned* counter, int num_sms, cudaStream_t s) {
    int const grid = blocks_for(n_rows, 256, num_sms * 4);
    PBK_DISPATCH(dtype, (dot_moment_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(beta), static_cast<const T*>(v), n_rows, mom, n, scale, scratch, counter)));
    return cudaGetLastError();
}

cudaError_t launch_accumulate_lanes(const double* mom, int R, int M, double* acc, cudaStream_t s) {
    accumulate_lanes_kernel<<<(2 * M + 255) / 256, 256, 0, s>>>(mom, R, M, acc);
    return cudaGetLastError();
}

cudaError_t launch_scatter_block(int dtype, const double* src_c128, int64_t n, int R, int lane, const int32_t* perm_dev, void* dst, cudaStream_t s) {
    int const grid = static_cast<int>((n + 255) / 256);
    PBK_DISPATCH(dtype, (scatter_block_kernel<T><<<grid, 256, 0, s>>>(src_c128, n, R, lane, perm_dev, static_cast<T*>(dst))));
    return cudaGetLastError();
}

cudaError_t launch_extract_lane(int dtype, const void* v, int64_t n, int R, int lane, const int32_t* perm_dev, double* out_c128, cudaStream_t s) {
    int const grid = static_cast<int>((n + 255) / 256);
    PBK_DISPATCH(dtype, (extract_lane_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(v), n, R, lane, perm_dev, out_c128)));
    return cudaGetLastError();
}

cudaError_t launch_cone_group_start(int dtype, const ConeSlot* slots_dev, int nslots, int64_t max_nvec, cudaStream_t s) {
    int gx = static_cast<int>((max_nvec + 255) / 256);
    if (gx > 512) gx = 512;
    if (gx < 1) gx = 1;
    dim3 const grid(gx, nslots);
    PBK_DISPATCH(dtype, (cone_group_start_kernel<T><<<grid, 256, 0, s>>>(slots_dev)));
    return cudaGetLastError();
}

cudaError_t launch_cone_group_step(int dtype, const ConeSlot* slots_dev, int nslots, int k, int kell, int M, int64_t max_rows,
                                   double* partials, unsigned* counters, int blocks_per_slot_cap, cudaStream_t s) {
    int64_t need = (max_rows + 255) / 256;
    if (need < 1) need = 1;
    int const gx = static_cast<int>(need < blocks_per_slot_cap ? need : blocks_per_slot_cap);
    dim3 const grid(gx, nslots);
    switch (kell) {
        case 3: PBK_DISPATCH(dtype, (cone_group_step_kernel<T, 3><<<grid, 256, 0, s>>>(slots_dev, k, kell, M, partials, counters))); break;
        case 4: PBK_DISPATCH(dtype, (cone_group_step_kernel<T, 4><<<grid, 256, 0, s>>>(slots_dev, k, kell, M, partials, counters))); break;
        case 7: PBK_DISPATCH(dtype, (cone_group_step_kernel<T, 7><<<grid, 256, 0, s>>>(slots_dev, k, kell, M, partials, counters))); break;
        default: PBK_DISPATCH(dtype, (cone_group_step_kernel<T, 0><<<grid, 256, 0, s>>>(slots_dev, k, kell, M, partials, counters))); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_cone_mark(const int32_t* queue_dev, int64_t count, const int32_t* perm_dev, int32_t* gmap, bool set, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    cone_mark_kernel<<<static_cast<int>((count + 255) / 256), 256, 0, s>>>(queue_dev, count, perm_dev, gmap, set ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_cone_extract(int dtype, EllDev const& h, const int32_t* queue_dev, const int32_t* perm_dev, const int32_t* gmap,
                                int64_t rows, void* out_val, int32_t* out_col, int64_t out_pitch, cudaStream_t s) {
    int const grid = static_cast<int>((out_pitch + 255) / 256);
    PBK_DISPATCH(dtype, (cone_extract_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(h.val), h.col, h.pitch, h.k, queue_dev, perm_dev, gmap,
                                                                      rows, static_cast<T*>(out_val), out_col, out_pitch)));
    return cudaGetLastError();
}

cudaError_t launch_lanczos_update(int dtype, int64_t n, const void* t, const void* v1, void* v0, double b_prev, const double* a_dev,
                                  double* out, double* scratch, unsigned* counter, int num_sms, cudaStream_t s) {
    int const grid = blocks_for(n, 256, num_sms * 4);
    PBK_DISPATCH(dtype, (lanczos_update_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(t), static_cast<const T*>(v1), static_cast<T*>(v0), n, b_prev, a_dev, out, scratch, counter)));
    return cudaGetLastError();
}

cudaError_t launch_scale_inv_sqrt(int dtype, int64_t n, void* v, const double* norm2_dev, cudaStream_t s) {
    int const grid = static_cast<int>((n + 255) / 256);
    PBK_DISPATCH(dtype, (scale_inv_sqrt_kernel<T><<<grid, 256, 0, s>>>(static_cast<T*>(v), n, norm2_dev)));
    return cudaGetLastError();
}

} // namespace pbk
