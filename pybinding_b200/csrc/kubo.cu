// kubo.cu -- K4: the Kubo-Bastin moment matrix  mu (M x M) += L (M x N) * R^H (N x M).
//
// Replaces MomentMultiplication::matrix_mul_add (cppcore/src/kpm/Moments.cpp:92-101,123-126), which
// the reference evaluates as a single-threaded Eigen GEMM on two host-resident M x N stacks.
// This is the one dense, tensor-core-shaped piece of the KPM path.  tcgen05/UMMA has no fp64 kind,
// so the fp64 tensor pipe is driven with mma.sync.m8n8k4.f64 (DMMA); f32/c64 stacks are widened to
// fp64 while staging into shared memory, so every scalar type accumulates in double.
//
// Both stacks are K-major (each moment row is contiguous over the N sites), i.e. the "TN" case where
// A and B fragments use the same access pattern.  Complex stacks are treated as real M x 2N matrices:
//   Re(mu) = A' * B'^T                      with A' = [.. ar_k, ai_k ..], B' = [.. br_k, bi_k ..]
//   Im(mu) = A'' * B'^T                     with A'' = [.. ai_k, -ar_k ..]   (pair swap + negate on load)
// Split-K over N with per-split partial tiles and a fixed-order reduction keeps the result deterministic.
#include "kernels.cuh"

namespace pbk {

namespace {

constexpr int TM = 64, TN = 64, TK = 32, PAD = 4;
constexpr int GEMM_THREADS = 128;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

/// part[z][Mp][Mp] = A[:, kz] * B[:, kz]^T  over the K-range of split z.  `ld` = row stride in real elements.
template<class Real, bool IMAG>
__global__ void __launch_bounds__(GEMM_THREADS) kubo_gemm_kernel(const Real* __restrict__ A, const Real* __restrict__ B, int M, int64_t K,
                                                                 int64_t ld, double* __restrict__ part, int Mp, int64_t kchunk) {
    __shared__ double As[TM][TK + PAD];
    __shared__ double Bs[TN][TK + PAD];
    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const wm = warp >> 1, wn = warp & 1;
    int const m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    int64_t const kbeg = static_cast<int64_t>(blockIdx.z) * kchunk;
    int64_t const kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;

    double c[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

    for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll 4
        for (int i = tid; i < TM * TK; i += GEMM_THREADS) {
            int const r = i / TK, cc = i % TK;
            int64_t const kk = k0 + cc;
            double va = 0.0, vb = 0.0;
            if (kk < kend) {
                if (m0 + r < M) {
                    if (IMAG) { double const t = static_cast<double>(A[static_cast<int64_t>(m0 + r) * ld + (kk ^ 1)]); va = (kk & 1) ? -t : t; }
                    else va = static_cast<double>(A[static_cast<int64_t>(m0 + r) * ld + kk]);
                }
                if (n0 + r < M) vb = static_cast<double>(B[static_cast<int64_t>(n0 + r) * ld + kk]);
            }
            As[r][cc] = va;
            Bs[r][cc] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[wm * 32 + i * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[wn * 32 + j * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    double* out = part + static_cast<int64_t>(blockIdx.z) * Mp * Mp;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int const row = m0 + wm * 32 + i * 8 + (lane >> 2);
            int const col = n0 + wn * 32 + j * 8 + (lane & 3) * 2;
            out[static_cast<int64_t>(row) * Mp + col] = c[i][j][0];
            out[static_cast<int64_t>(row) * Mp + col + 1] = c[i][j][1];
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

template<class Real>
cudaError_t gemm_t(const void* A, const void* B, int M, int64_t N, bool cplx, double* C, int num_sms, cudaStream_t s, double* flops) {
    int const tiles = (M + TM - 1) / TM;
    int const Mp = tiles * TM;
    int64_t const K = cplx ? 2 * N : N;
    int ksplit = (2 * num_sms + tiles * tiles - 1) / (tiles * tiles);
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 64) ksplit = 64;
    int64_t kchunk = (K + ksplit - 1) / ksplit;
    kchunk = (kchunk + TK - 1) / TK * TK;
    ksplit = static_cast<int>((K + kchunk - 1) / kchunk);

    double* part = nullptr;
    cudaError_t err = cudaMallocAsync(&part, sizeof(double) * ksplit * Mp * Mp, s);
    if (err != cudaSuccess) return err;
    dim3 const grid(tiles, tiles, ksplit);
    auto const* a = static_cast<const Real*>(A);
    auto const* b = static_cast<const Real*>(B);
    kubo_gemm_kernel<Real, false><<<grid, GEMM_THREADS, 0, s>>>(a, b, M, K, K, part, Mp, kchunk);
    kubo_reduce_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(part, ksplit, Mp, M, C, 0);
    if (cplx) {
        kubo_gemm_kernel<Real, true><<<grid, GEMM_THREADS, 0, s>>>(a, b, M, K, K, part, Mp, kchunk);
        kubo_reduce_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(part, ksplit, Mp, M, C, 1);
    }
    err = cudaGetLastError();
    cudaFreeAsync(part, s);
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

cudaError_t launch_kubo_gemm(int dtype, const void* A, const void* B, int M, int64_t N, double* C_c128, int num_sms,
                             cudaStream_t s, double* flops) {
    switch (dtype) {
        case F32: return gemm_t<float>(A, B, M, N, false, C_c128, num_sms, s, flops);
        case C64: return gemm_t<float>(A, B, M, N, true, C_c128, num_sms, s, flops);
        case F64: return gemm_t<double>(A, B, M, N, false, C_c128, num_sms, s, flops);
        case C128: return gemm_t<double>(A, B, M, N, true, C_c128, num_sms, s, flops);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
