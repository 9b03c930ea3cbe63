// reconstruct.cu -- Chebyshev series evaluation on the device: reconstruct<SpectralDensity> and
// reconstruct<GreensFunction> of the reference (cppcore/include/kpm/reconstruct.hpp:16-70), the step *after* the
// moment recursion.  At BASELINE sizes (256 LDOS sites x 500 energies x 1400 moments) the host version costs
// seconds of transcendental evaluations -- longer than the recursion of a Green's function -- so the sums run here:
// one thread per (energy, column), moments broadcast through the read-only path, double precision throughout and
// the same term order as the reference (q ascending), with its single-precision constants applied by the caller.
#include "kernels.cuh"

namespace pbk {

namespace {

/// out[c * ne + i] = k / sqrt(1 - E_i^2) * sum_q Re(mu[q * n_stride + c * col_stride]) * cos(q * acos(E_i))
__global__ void __launch_bounds__(128) spectral_density_kernel(const double2* __restrict__ mu, int M, int64_t col_stride, int64_t n_stride,
                                                               const double* __restrict__ scaled_energy, int ne, double k, double* __restrict__ out) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    int const c = blockIdx.y;
    if (i >= ne) return;
    double const E = scaled_energy[i];
    double const ac = acos(E);
    const double2* const m = mu + static_cast<int64_t>(c) * col_stride;
    double sum = 0.0;
    for (int q = 0; q < M; ++q) sum += __ldg(&m[static_cast<int64_t>(q) * n_stride].x) * cos(q * ac);
    out[static_cast<int64_t>(c) * ne + i] = k / sqrt(1.0 - E * E) * sum;
}

/// out[c * ne + i] = (-2i / a) / sqrt(1 - E_i^2) * sum_q mu[c * M + q] * exp(-i q acos(E_i))
__global__ void __launch_bounds__(128) greens_kernel(const double2* __restrict__ mu, int M, const double* __restrict__ scaled_energy, int ne,
                                                     double inv_a, double2* __restrict__ out) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    int const c = blockIdx.y;
    if (i >= ne) return;
    double const E = scaled_energy[i];
    double const ac = acos(E);
    const double2* const m = mu + static_cast<int64_t>(c) * M;
    double re = 0.0, im = 0.0;
    for (int q = 0; q < M; ++q) {
        double s, co;
        sincos(q * ac, &s, &co);      // exp(-i q ac) = co - i s
        double2 const v = __ldg(&m[q]);
        re += v.x * co + v.y * s;
        im += v.y * co - v.x * s;
    }
    double const f = 2.0 * inv_a / sqrt(1.0 - E * E);   // (-2i f') * (re + i im) = f * (im - i re)
    out[static_cast<int64_t>(c) * ne + i] = make_double2(f * im, -f * re);
}

} // anonymous namespace

cudaError_t launch_spectral_density(const double* mu_c128, int M, int cols, int64_t col_stride, int64_t n_stride, const double* scaled_energy,
                                    int ne, double k, double* out, cudaStream_t s) {
    if (ne <= 0 || cols <= 0) return cudaSuccess;
    dim3 const grid((ne + 127) / 128, cols);
    spectral_density_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const double2*>(mu_c128), M, col_stride, n_stride, scaled_energy, ne, k, out);
    return cudaGetLastError();
}

cudaError_t launch_greens(const double* mu_c128, int M, int cols, const double* scaled_energy, int ne, double inv_a, double* out_c128,
                          cudaStream_t s) {
    if (ne <= 0 || cols <= 0) return cudaSuccess;
    dim3 const grid((ne + 127) / 128, cols);
    greens_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const double2*>(mu_c128), M, scaled_energy, ne, inv_a, reinterpret_cast<double2*>(out_c128));
    return cudaGetLastError();
}

} // namespace pbk
