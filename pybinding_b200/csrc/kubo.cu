// kubo.cu -- K4: the Kubo-Bastin moment matrix  mu (M x M) += L (M x N) * R^H (N x M).
//
// Replaces MomentMultiplication::matrix_mul_add (cppcore/src/kpm/Moments.cpp:92-101,123-126), which
// the reference evaluates as a single-threaded Eigen GEMM on two host-resident M x N stacks.
// This is the one dense, tensor-core-shaped piece of the KPM path.  tcgen05/UMMA has no fp64 kind, so the
// fp64 tensor pipe is driven with mma.sync f64 -- on sm_100a every f64 shape compiles to DMMA.8x8x4, measured
// peak 37.1 TFLOP/s (tools/probe/dmma_probe.cu, profiles/r01_fp64_mma_probe.jsonl).  f32/c64 stacks are widened
// to fp64 when the fragments are read from shared memory, so every scalar type accumulates in double.
//
// Both stacks are K-major (each moment row is contiguous over the N sites), i.e. the "TN" case where
// A and B fragments use the same access pattern.  Complex stacks are treated as real M x 2N matrices:
//   Re(mu) = A' * B'^T                      with A' = [.. ar_k, ai_k ..], B' = [.. br_k, bi_k ..]
//   Im(mu) = A'' * B'^T                     with A'' = [.. ai_k, -ar_k ..]   (pair swap + negate on fragment load)
//
// Kernel: 128 x 128 tile per CTA (8 warps, 32 x 64 per warp = 32 DMMAs per k4-step out of 12 fragment loads),
// operands streamed global -> shared by a 3-stage cp.async ring of 128-byte rows (16 doubles / 32 floats per
// stage, zero-filled past the edges), padded row stride so that fragment loads are bank-conflict free.
// M is rarely a multiple of the tile (the reference's num_moments is 4k+2): 8 x 8 blocks that lie entirely
// outside the matrix are skipped, so the cost follows ceil(M/8)^2 blocks, not the padded tile area.
// Split-K over N with per-split partial tiles and a fixed-order reduction keeps the result deterministic.
#include "kernels.cuh"

#include <cstdlib>
#include <type_traits>

namespace pbk {

namespace {

constexpr int BM = 128, BN = 128;       // CTA tile
// warps: 4 (m) x WN (n); WN = 2: 8 warps with 32 x 64 warp tiles, WN = 4: 16 warps with 32 x 32 warp tiles
constexpr int GEMM_STAGES = 3;

// ROW_BYTES: bytes of one operand row per pipeline stage (128: 16 doubles, 256: 32 doubles -- half as many CTA barriers
// per flop, 221 KB of shared memory for the three stages)
template<class Real, int ROW_BYTES> struct Geom {
    static constexpr int TKE = ROW_BYTES / sizeof(Real);          // k-extent of a stage in elements
    static constexpr int PAD = sizeof(Real) == 8 ? 32 : 16;       // bytes: stride = 20 doubles / 36 floats
    static constexpr int STRIDE = ROW_BYTES + PAD;                // bytes between rows in shared memory
    static constexpr int TILE = BM * STRIDE;                      // one operand, one stage
    static constexpr int STAGE = 2 * TILE;
    static constexpr int SMEM = GEMM_STAGES * STAGE;
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {  // zero-fills past src_bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double lds_as_double(uint32_t addr, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ double lds_as_double(uint32_t addr, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return static_cast<double>(v); }

/// part[z][Mp][Mp] = A[:, kz] * B[:, kz]^T  over the K-range of split z.  `ld` = row stride in real elements
/// (a multiple of 16 bytes); kchunk is a multiple of the stage extent.
template<class Real, bool IMAG, int WN, int ROW_BYTES>
__global__ void __launch_bounds__(128 * WN, 1) kubo_gemm_kernel(const Real* __restrict__ A, const Real* __restrict__ B, int M, int64_t K,
                                                                    int64_t ld, double* __restrict__ part, int Mp, int64_t kchunk) {
    using G = Geom<Real, ROW_BYTES>;
    constexpr int TKE = G::TKE;
    constexpr int EPC = 16 / sizeof(Real);   // elements per 16-byte chunk
    extern __shared__ __align__(128) unsigned char gemm_smem[];
    uint32_t const smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(gemm_smem));

    constexpr int GEMM_THREADS = 128 * WN;
    constexpr int NB = BN / WN / 8;          // 8-column blocks per warp: 8 (WN = 2) or 4 (WN = 4)
    constexpr int WCOLS = BN / WN;
    constexpr int CPRW = ROW_BYTES / 16;                  // 16-byte chunks per operand row and stage
    constexpr int LOADS = BM * CPRW / GEMM_THREADS;       // 16-byte chunks per thread, per operand, per stage
    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const wm = warp / WN, wn = warp % WN;
    int const m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int64_t const kbeg = static_cast<int64_t>(blockIdx.z) * kchunk;
    int64_t const kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
    int const nstage = static_cast<int>((kend - kbeg + TKE - 1) / TKE);

    // 8 x 8 blocks of this warp that intersect the matrix (warp-uniform)
    int mi_cnt = (M - (m0 + wm * 32) + 7) / 8; mi_cnt = mi_cnt < 0 ? 0 : (mi_cnt > 4 ? 4 : mi_cnt);
    int nj_cnt = (M - (n0 + wn * WCOLS) + 7) / 8; nj_cnt = nj_cnt < 0 ? 0 : (nj_cnt > NB ? NB : nj_cnt);

    // ---- producer side: each thread moves LOADS chunks of A and of B per stage ----
    constexpr int RSTEP = GEMM_THREADS / CPRW;
    int const lrow = tid / CPRW, lch = tid % CPRW;       // rows lrow + RSTEP i, 16-byte chunk lch of the operand row
    auto load_stage = [&](int st, int slot) {
        int64_t const k0 = kbeg + static_cast<int64_t>(st) * TKE + lch * EPC;
        int64_t const left = (kend - k0) * static_cast<int64_t>(sizeof(Real));
        uint32_t const kbytes = left <= 0 ? 0u : (left >= 16 ? 16u : static_cast<uint32_t>(left));
        uint32_t const dst0 = smem0 + slot * G::STAGE + lrow * G::STRIDE + lch * 16;
#pragma unroll
        for (int i = 0; i < LOADS; ++i) {
            int const r = lrow + RSTEP * i;
            bool const okA = (m0 + r < M) && kbytes > 0, okB = (n0 + r < M) && kbytes > 0;
            const Real* const srcA = okA ? A + static_cast<int64_t>(m0 + r) * ld + k0 : A;
            const Real* const srcB = okB ? B + static_cast<int64_t>(n0 + r) * ld + k0 : B;
            cp_async16(dst0 + i * RSTEP * G::STRIDE, srcA, okA ? kbytes : 0u);
            cp_async16(dst0 + G::TILE + i * RSTEP * G::STRIDE, srcB, okB ? kbytes : 0u);
        }
    };

    double c[4][NB][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

#pragma unroll
    for (int st = 0; st < GEMM_STAGES - 1; ++st) {
        if (st < nstage) load_stage(st, st);
        cp_async_commit();
    }

    // fragment addresses: thread holds A[row = lane / 4][k = lane % 4] of each 8 x 4 block (B alike)
    int const fr = lane >> 2, fk = lane & 3;
    int const fkA = IMAG ? (fk ^ 1) : fk;                     // A'' = [ai, -ar]: neighbour element, sign by parity
    double const sgnA = (IMAG && (fk & 1)) ? -1.0 : 1.0;
    uint32_t const offA = (wm * 32 + fr) * G::STRIDE + fkA * sizeof(Real);
    uint32_t const offB = G::TILE + (wn * WCOLS + fr) * G::STRIDE + fk * sizeof(Real);

    // main loop, instantiated twice: interior tiles run the branch-free version, tiles on the matrix edge skip the
    // 8 x 8 blocks that lie outside (both conditions are CTA-uniform)
    auto mainloop = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        int slot = 0;
        for (int st = 0; st < nstage; ++st) {
            cp_async_wait<GEMM_STAGES - 2>();   // stage st has landed (for this thread's copies)
            __syncthreads();                    // ... for everyone's; and everyone is done with the slot refilled below
            {
                int const nxt = st + GEMM_STAGES - 1;
                int nslot = slot + GEMM_STAGES - 1; if (nslot >= GEMM_STAGES) nslot -= GEMM_STAGES;
                if (nxt < nstage) load_stage(nxt, nslot);
                cp_async_commit();
            }
            uint32_t const base = smem0 + slot * G::STAGE;
#pragma unroll
            for (int kk = 0; kk < TKE; kk += 4) {
                double a[4], b[NB];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sgnA * lds_as_double(base + offA + i * 8 * G::STRIDE + kk * sizeof(Real), Real{});
#pragma unroll
                for (int j = 0; j < NB; ++j) b[j] = lds_as_double(base + offB + j * 8 * G::STRIDE + kk * sizeof(Real), Real{});
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!EDGE || i < mi_cnt) {
#pragma unroll
                        for (int j = 0; j < NB; ++j)
                            if (!EDGE || j < nj_cnt) dmma_m8n8k4(c[i][j][0], c[i][j][1], a[i], b[j]);
                    }
                }
            }
            if (++slot == GEMM_STAGES) slot = 0;
        }
    };
    if (m0 + BM <= M && n0 + BN <= M) mainloop(std::false_type{});
    else mainloop(std::true_type{});
    cp_async_wait<0>();

    double* out = part + static_cast<int64_t>(blockIdx.z) * Mp * Mp;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            int const row = m0 + wm * 32 + i * 8 + (lane >> 2);
            int const col = n0 + wn * WCOLS + j * 8 + (lane & 3) * 2;
            *reinterpret_cast<double2*>(out + static_cast<int64_t>(row) * Mp + col) = make_double2(c[i][j][0], c[i][j][1]);
        }
}

__global__ void kubo_reduce_kernel(const double* part, int ksplit, int Mp, int M, double* C, int comp) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * M) return;
    int const m = i / M, n = i % M;
    double s = 0.0;
    for (int z = 0; z < ksplit; ++z) s += part[(static_cast<int64_t>(z) * Mp + m) * Mp + n];
    C[static_cast<int64_t>(i) * 2 + comp] += s;
}

static int gemm_row_bytes() {
    static int const rb = [] { char const* v = std::getenv("PBK_KUBO_ROW"); return (v && std::atoi(v) == 128) ? 128 : 256; }();
    return rb;
}

template<class Real>
void gemm_plan(int M, int64_t N, bool cplx, int num_sms, int* Mp, int64_t* K, int* ksplit, int64_t* kchunk) {
    using G = Geom<Real, 256>;   // split sizes in units of the larger stage extent (valid for both)
    int const tiles = (M + BM - 1) / BM;
    *Mp = tiles * BM;
    *K = cplx ? 2 * N : N;
    // split-K: a few waves of CTAs (one CTA per SM) so that neither the light edge tiles nor the last wave leave
    // SMs idle for long; each split costs one 128 x 128 partial tile of traffic, negligible next to its k-range
    static int const waves = [] { char const* v = std::getenv("PBK_KUBO_WAVES"); int w = v ? std::atoi(v) : 8; return w < 1 ? 1 : w; }();
    int ks = (waves * num_sms + tiles * tiles - 1) / (tiles * tiles);
    int64_t const max_split = (*K + 8 * G::TKE - 1) / (8 * G::TKE);   // at least 8 stages per CTA
    if (ks > max_split) ks = static_cast<int>(max_split);
    if (ks > 512) ks = 512;
    if (ks < 1) ks = 1;
    int64_t kc = (*K + ks - 1) / ks;
    kc = (kc + G::TKE - 1) / G::TKE * G::TKE;
    *ksplit = static_cast<int>((*K + kc - 1) / kc);
    *kchunk = kc;
}

template<class Real, int WN, int ROW_BYTES>
cudaError_t gemm_launch(const Real* a, const Real* b, int M, int64_t K, int64_t ld, bool cplx, double* C, double* part, int Mp, int ksplit,
                        int64_t kchunk, cudaStream_t s) {
    using G = Geom<Real, ROW_BYTES>;
    static bool raised = false;
    auto const k_re = kubo_gemm_kernel<Real, false, WN, ROW_BYTES>;
    auto const k_im = kubo_gemm_kernel<Real, true, WN, ROW_BYTES>;
    if (!raised) {
        cudaError_t e = cudaFuncSetAttribute(k_re, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_im, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        if (e != cudaSuccess) return e;
        raised = true;
    }
    int const tiles = Mp / BM;
    dim3 const grid(tiles, tiles, ksplit);
    k_re<<<grid, 128 * WN, G::SMEM, s>>>(a, b, M, K, ld, part, Mp, kchunk);
    kubo_reduce_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(part, ksplit, Mp, M, C, 0);
    if (cplx) {
        k_im<<<grid, 128 * WN, G::SMEM, s>>>(a, b, M, K, ld, part, Mp, kchunk);
        kubo_reduce_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(part, ksplit, Mp, M, C, 1);
    }
    return cudaGetLastError();
}

template<class Real>
cudaError_t gemm_t(const void* A, const void* B, int M, int64_t N, int64_t pitch_bytes, bool cplx, double* C, double* workspace,
                   size_t workspace_bytes, int num_sms, cudaStream_t s, double* flops) {
    if (pitch_bytes % 16 != 0 || reinterpret_cast<uintptr_t>(A) % 16 != 0 || reinterpret_cast<uintptr_t>(B) % 16 != 0) return cudaErrorInvalidValue;
    int Mp = 0, ksplit = 0;
    int64_t K = 0, kchunk = 0;
    gemm_plan<Real>(M, N, cplx, num_sms, &Mp, &K, &ksplit, &kchunk);
    if (workspace_bytes < sizeof(double) * static_cast<size_t>(ksplit) * Mp * Mp) return cudaErrorInvalidValue;
    int64_t const ld = pitch_bytes / static_cast<int64_t>(sizeof(Real));
    static int const warps_n = [] { char const* v = std::getenv("PBK_KUBO_WN"); return (v && std::atoi(v) == 2) ? 2 : 4; }();
    auto const* a = static_cast<const Real*>(A);
    auto const* b = static_cast<const Real*>(B);
    cudaError_t err;
    if (gemm_row_bytes() == 128) {
        err = warps_n == 2 ? gemm_launch<Real, 2, 128>(a, b, M, K, ld, cplx, C, workspace, Mp, ksplit, kchunk, s)
                           : gemm_launch<Real, 4, 128>(a, b, M, K, ld, cplx, C, workspace, Mp, ksplit, kchunk, s);
    } else {
        err = warps_n == 2 ? gemm_launch<Real, 2, 256>(a, b, M, K, ld, cplx, C, workspace, Mp, ksplit, kchunk, s)
                           : gemm_launch<Real, 4, 256>(a, b, M, K, ld, cplx, C, workspace, Mp, ksplit, kchunk, s);
    }
    if (flops) *flops = 2.0 * M * M * static_cast<double>(K) * (cplx ? 2 : 1);
    return err;
}

/// K7: sum_{m,n} mu_mn * Gamma_mn(E) / (1 - E^2)^2 for one energy sample per block, with
/// Gamma = g + g^H, g_mn = (E - i n sqrt(1 - E^2)) exp(i n acos E) cos(m acos E)   (kpm/reconstruct.hpp:108-131)
__global__ void __launch_bounds__(256) kubo_gamma_sum_kernel(const double2* __restrict__ mu, int M, const double* __restrict__ samples, double2* out) {
    extern __shared__ double sh[];
    double2* a = reinterpret_cast<double2*>(sh);  // a_n = sqrt_n * exp_n  (column factor)
    double* t = sh + 2 * M;                        // t_m = cos(m acos E)   (row factor)
    double const e = samples[blockIdx.x];
    double const ac = acos(e);
    double const sq = sqrt(1.0 - e * e);
    for (int q = threadIdx.x; q < M; q += blockDim.x) {
        double s, c;
        sincos(ac * q, &s, &c);
        // (e - i q sq) * (c + i s)
        a[q] = make_double2(e * c + q * sq * s, e * s - q * sq * c);
        t[q] = c;
    }
    __syncthreads();
    double re = 0.0, im = 0.0;
    for (int64_t i = threadIdx.x; i < static_cast<int64_t>(M) * M; i += blockDim.x) {
        int const m = static_cast<int>(i / M), n = static_cast<int>(i % M);
        // gamma = a_n t_m + conj(a_m t_n)
        double const gr = a[n].x * t[m] + a[m].x * t[n];
        double const gi = a[n].y * t[m] - a[m].y * t[n];
        double2 const v = mu[i];
        re += v.x * gr - v.y * gi;
        im += v.x * gi + v.y * gr;
    }
    __shared__ double red[2 * 256];
    red[threadIdx.x] = re; red[256 + threadIdx.x] = im;
    for (int s = 128; s > 0; s >>= 1) {
        __syncthreads();
        if (threadIdx.x < s) { red[threadIdx.x] += red[threadIdx.x + s]; red[256 + threadIdx.x] += red[256 + threadIdx.x + s]; }
    }
    if (threadIdx.x == 0) {
        double const k = 1.0 / ((1.0 - e * e) * (1.0 - e * e));
        out[blockIdx.x] = make_double2(k * red[0], k * red[256]);
    }
}

} // anonymous namespace

cudaError_t launch_kubo_gamma_sum(const double* mu_c128, int M, const double* scaled_samples, int np, double* out_c128, cudaStream_t s) {
    size_t const smem = sizeof(double) * 3 * static_cast<size_t>(M);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t err = cudaFuncSetAttribute(kubo_gamma_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    kubo_gamma_sum_kernel<<<np, 256, smem, s>>>(reinterpret_cast<const double2*>(mu_c128), M, scaled_samples, reinterpret_cast<double2*>(out_c128));
    return cudaGetLastError();
}

size_t kubo_gemm_workspace_bytes(int dtype, int M, int64_t N, int num_sms) {
    int Mp = 0, ksplit = 0;
    int64_t K = 0, kchunk = 0;
    bool const cplx = dtype_complex(dtype);
    if (dtype == F32 || dtype == C64) gemm_plan<float>(M, N, cplx, num_sms, &Mp, &K, &ksplit, &kchunk);
    else gemm_plan<double>(M, N, cplx, num_sms, &Mp, &K, &ksplit, &kchunk);
    return sizeof(double) * static_cast<size_t>(ksplit) * Mp * Mp;
}

cudaError_t launch_kubo_gemm(int dtype, const void* A, const void* B, int M, int64_t N, int64_t pitch_bytes, double* C_c128,
                             double* workspace, size_t workspace_bytes, int num_sms, cudaStream_t s, double* flops) {
    switch (dtype) {
        case F32: return gemm_t<float>(A, B, M, N, pitch_bytes, false, C_c128, workspace, workspace_bytes, num_sms, s, flops);
        case C64: return gemm_t<float>(A, B, M, N, pitch_bytes, true, C_c128, workspace, workspace_bytes, num_sms, s, flops);
        case F64: return gemm_t<double>(A, B, M, N, pitch_bytes, false, C_c128, workspace, workspace_bytes, num_sms, s, flops);
        case C128: return gemm_t<double>(A, B, M, N, pitch_bytes, true, C_c128, workspace, workspace_bytes, num_sms, s, flops);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
