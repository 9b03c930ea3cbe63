// mt19937.cu -- K5: the reference's stochastic starters, generated on the device.
//
// The reference draws every random starter from one default-seeded std::mt19937 per calculation
// (cppcore/src/kpm/Starter.cpp:48-83, numeric/random.hpp:25-50): vector j consumes draws
// [j*N*w, (j+1)*N*w) with w = 1 word (float) or 2 words (double) per site.  To use *identical*
// starting vectors the same bit stream is reproduced here:
//   * the 624-word twist runs in three dependent stages of <= 227 independent elements
//     (k < 227 reads only old words; 227 <= k < 454 and k >= 454 read words written one stage earlier),
//   * libstdc++'s generate_canonical: float  x = float(u) * 2^-32           (1 word)
//                                     double x = (u1 + u2 * 2^32) * 2^-64    (2 words), clamped below 1,
//   * real starters:    r = (x < 0.5) ? -1 : +1
//     complex starters: r = exp(i * k * x),  k = 2 * pi_float = 6.2831854820251465 (Starter.cpp:75).
#include "kernels.cuh"

namespace pbk {

namespace {

constexpr int MT_M = 397;
constexpr int MT_THREADS = 256;

__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__device__ __forceinline__ uint32_t twist_word(uint32_t cur, uint32_t nxt, uint32_t far_) {
    uint32_t const yy = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far_ ^ (yy >> 1) ^ ((yy & 1u) ? 0x9908b0dfu : 0u);
}

/// Regenerate the 624-word block in shared memory (all threads of the block call this)
__device__ void twist(uint32_t* mt) {
    int const t = threadIdx.x;
    // stage A: k in [0, 227) -- inputs are all old
    uint32_t w = 0;
    if (t < 227) w = twist_word(mt[t], mt[t + 1], mt[t + MT_M]);
    __syncthreads();
    if (t < 227) mt[t] = w;
    __syncthreads();
    // stage B: k in [227, 454) -- mt[k - 227] is new (stage A)
    int k = t + 227;
    if (t < 227) w = twist_word(mt[k], mt[k + 1], mt[k - 227]);
    __syncthreads();
    if (t < 227) mt[k] = w;
    __syncthreads();
    // stage C: k in [454, 624) -- mt[k - 227] is new (stage B); k = 623 wraps to the new mt[0]
    k = t + 454;
    if (k < MT_N) w = twist_word(mt[k], mt[(k + 1) % MT_N], mt[k - 227]);
    __syncthreads();
    if (k < MT_N) mt[k] = w;
    __syncthreads();
}

__global__ void mt_seed_kernel(uint32_t* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t x = 5489u;
        state[0] = x;
        for (uint32_t i = 1; i < MT_N; ++i) { x = 1812433253u * (x ^ (x >> 30)) + i; state[i] = x; }
        state[MT_N] = MT_N;  // position: the first draw triggers a twist
    }
}

__global__ void __launch_bounds__(MT_THREADS) mt_generate_kernel(uint32_t* state, uint32_t* out, int64_t count) {
    __shared__ uint32_t mt[MT_N];
    for (int i = threadIdx.x; i < MT_N; i += MT_THREADS) mt[i] = state[i];
    int pos = static_cast<int>(state[MT_N]);
    __syncthreads();
    int64_t produced = 0;
    while (produced < count) {
        if (pos == MT_N) { twist(mt); pos = 0; }
        int64_t const left = count - produced;
        int const avail = static_cast<int>(left < (MT_N - pos) ? left : (MT_N - pos));
        if (out) {
            for (int i = threadIdx.x; i < avail; i += MT_THREADS) out[produced + i] = temper(mt[pos + i]);
        }
        pos += avail;
        produced += avail;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MT_N; i += MT_THREADS) state[i] = mt[i];
    if (threadIdx.x == 0) state[MT_N] = static_cast<uint32_t>(pos);
}

__device__ __forceinline__ float canonical_f32(uint32_t u) {
    float x = __uint2float_rn(u) * 2.3283064365386963e-10f;  // * 2^-32, exact
    return x >= 1.0f ? 0.99999994f : x;
}
__device__ __forceinline__ double canonical_f64(uint32_t u1, uint32_t u2) {
    double const sum = fma(static_cast<double>(u2), 4294967296.0, static_cast<double>(u1));
    double x = sum * 5.421010862427522e-20;  // * 2^-64, exact
    return x >= 1.0 ? 0.99999999999999989 : x;
}

template<class T> __device__ __forceinline__ T make_starter(const uint32_t* raw, int64_t i);
template<> __device__ __forceinline__ float make_starter<float>(const uint32_t* raw, int64_t i) {
    return canonical_f32(raw[i]) < 0.5f ? -1.f : 1.f;
}
template<> __device__ __forceinline__ double make_starter<double>(const uint32_t* raw, int64_t i) {
    return canonical_f64(raw[2 * i], raw[2 * i + 1]) < 0.5 ? -1.0 : 1.0;
}
template<> __device__ __forceinline__ float2 make_starter<float2>(const uint32_t* raw, int64_t i) {
    float const kx = 6.2831854820251465f * canonical_f32(raw[i]);  // float multiply, like complex<float> * float
    double const a = static_cast<double>(kx);
    return make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
}
template<> __device__ __forceinline__ double2 make_starter<double2>(const uint32_t* raw, int64_t i) {
    double const kx = 6.2831854820251465 * canonical_f64(raw[2 * i], raw[2 * i + 1]);
    double s, c;
    sincos(kx, &s, &c);
    return make_double2(c, s);
}
template<class T> __device__ __forceinline__ T zero_starter();
template<> __device__ __forceinline__ float zero_starter<float>() { return 0.f; }
template<> __device__ __forceinline__ double zero_starter<double>() { return 0.0; }
template<> __device__ __forceinline__ float2 zero_starter<float2>() { return make_float2(0.f, 0.f); }
template<> __device__ __forceinline__ double2 zero_starter<double2>() { return make_double2(0.0, 0.0); }

template<class T, int W>
__global__ void random_transform_kernel(const uint32_t* raw, int64_t n, int R, int lanes_filled, const int32_t* perm, T* dst) {
    int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= n * R) return;
    int64_t const i = g / R;
    int const lane = static_cast<int>(g % R);
    int64_t const row = perm ? perm[i] : i;
    T v = zero_starter<T>();
    if (lane < lanes_filled) v = make_starter<T>(raw + static_cast<int64_t>(lane) * n * W, i);
    dst[row * R + lane] = v;
}

/// Lanczos start vector: uniform [0, 1) reals, zero imaginary part (compute/lanczos.hpp:105-107)
template<class T> __device__ __forceinline__ T make_uniform(const uint32_t* raw, int64_t i);
template<> __device__ __forceinline__ float make_uniform<float>(const uint32_t* raw, int64_t i) { return canonical_f32(raw[i]); }
template<> __device__ __forceinline__ double make_uniform<double>(const uint32_t* raw, int64_t i) { return canonical_f64(raw[2 * i], raw[2 * i + 1]); }
template<> __device__ __forceinline__ float2 make_uniform<float2>(const uint32_t* raw, int64_t i) { return make_float2(canonical_f32(raw[i]), 0.f); }
template<> __device__ __forceinline__ double2 make_uniform<double2>(const uint32_t* raw, int64_t i) { return make_double2(canonical_f64(raw[2 * i], raw[2 * i + 1]), 0.0); }

template<class T>
__global__ void uniform_transform_kernel(const uint32_t* raw, int64_t n, T* dst) {
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_uniform<T>(raw, i);
}

} // anonymous namespace

cudaError_t launch_uniform_transform(int dtype, const uint32_t* raw, int64_t n, void* dst, cudaStream_t s) {
    unsigned const grid = static_cast<unsigned>((n + 255) / 256);
    switch (dtype) {
        case F32: uniform_transform_kernel<float><<<grid, 256, 0, s>>>(raw, n, static_cast<float*>(dst)); break;
        case C64: uniform_transform_kernel<float2><<<grid, 256, 0, s>>>(raw, n, static_cast<float2*>(dst)); break;
        case F64: uniform_transform_kernel<double><<<grid, 256, 0, s>>>(raw, n, static_cast<double*>(dst)); break;
        case C128: uniform_transform_kernel<double2><<<grid, 256, 0, s>>>(raw, n, static_cast<double2*>(dst)); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_mt_seed(uint32_t* state_dev, cudaStream_t s) {
    mt_seed_kernel<<<1, 32, 0, s>>>(state_dev);
    return cudaGetLastError();
}

cudaError_t launch_mt_generate(uint32_t* state_dev, uint32_t* out, int64_t count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    mt_generate_kernel<<<1, MT_THREADS, 0, s>>>(state_dev, out, count);
    return cudaGetLastError();
}

cudaError_t launch_random_transform(int dtype, const uint32_t* raw, int64_t n, int R, int lanes_filled, const int32_t* perm_dev,
                                    void* dst, cudaStream_t s) {
    int64_t const total = n * R;
    unsigned const grid = static_cast<unsigned>((total + 255) / 256);
    switch (dtype) {
        case F32: random_transform_kernel<float, 1><<<grid, 256, 0, s>>>(raw, n, R, lanes_filled, perm_dev, static_cast<float*>(dst)); break;
        case C64: random_transform_kernel<float2, 1><<<grid, 256, 0, s>>>(raw, n, R, lanes_filled, perm_dev, static_cast<float2*>(dst)); break;
        case F64: random_transform_kernel<double, 2><<<grid, 256, 0, s>>>(raw, n, R, lanes_filled, perm_dev, static_cast<double*>(dst)); break;
        case C128: random_transform_kernel<double2, 2><<<grid, 256, 0, s>>>(raw, n, R, lanes_filled, perm_dev, static_cast<double2*>(dst)); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace pbk
