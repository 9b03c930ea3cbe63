// build.cu -- device-side construction of the scaled (and relabelled) ELL Hamiltonian from the caller's CSR arrays.
//
// Replaces, from the reference (cppcore/): OptimizedHamiltonian::create_scaled / create_reordered
// (src/kpm/OptimizedHamiltonian.cpp:55-152) fused with csr_to_ell (include/numeric/ellmatrix.hpp:65-82), and the
// velocity operator V_ij = H_ij (pos_i - pos_j) of MomentMultiplication (src/kpm/Moments.cpp:132-156).
//
// The unscaled CSR is uploaded once per Hamiltonian; every device layout (full-system locality order, breadth-first order
// of a Green's function source, unscaled copy for the Lanczos bounds, velocity operators) is then one kernel launch: thread
// = new row, reads the CSR entries of its original row, applies the scalar map with the reference's association
// ((v - b) * f unreordered, v * f - b * f reordered: explicit round-to-nearest intrinsics, no FMA contraction, so the
// values are bit-identical to the host restatement in engine.cu), relabels the columns through the
// order map, sorts the row by new column and writes slot-major ELL with coalesced stores.  The host never touches an
// array of the size of the system for this.
#include "kernels.cuh"
#include "step_common.cuh"

namespace pbk {
namespace {

constexpr int BUILD_KMAX = 32;   // rows longer than this are built on the host (engine.cu: build_ell_host)

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

template<class T> struct Ops;   // scalar maps on T with a real factor, mirroring std::complex<R> op R
template<> struct Ops<float> {
    using R = float;
    static __device__ float scale(float v, R f) { return mul_rn(v, f); }                                    // v * f
    static __device__ float diag_reordered(float v, R f, R sb) { return sub_rn(mul_rn(v, f), mul_rn(sb, f)); }  // v * f - sb * f
    static __device__ float diag_plain(float v, R f, R sb) { return mul_rn(sub_rn(v, sb), f); }                 // (v - sb) * f
    static __device__ float new_diag_reordered(R f, R sb) { return -mul_rn(sb, f); }                             // T{-sb * f}
    static __device__ float new_diag_plain(R f, R sb) { return mul_rn(sub_rn(0.f, sb), f); }                    // (T{0} - T{sb}) * f
};
template<> struct Ops<double> {
    using R = double;
    static __device__ double scale(double v, R f) { return mul_rn(v, f); }
    static __device__ double diag_reordered(double v, R f, R sb) { return sub_rn(mul_rn(v, f), mul_rn(sb, f)); }
    static __device__ double diag_plain(double v, R f, R sb) { return mul_rn(sub_rn(v, sb), f); }
    static __device__ double new_diag_reordered(R f, R sb) { return -mul_rn(sb, f); }
    static __device__ double new_diag_plain(R f, R sb) { return mul_rn(sub_rn(0.0, sb), f); }
};
template<> struct Ops<float2> {
    using R = float;
    static __device__ float2 scale(float2 v, R f) { return make_float2(mul_rn(v.x, f), mul_rn(v.y, f)); }
    static __device__ float2 diag_reordered(float2 v, R f, R sb) { return make_float2(sub_rn(mul_rn(v.x, f), mul_rn(sb, f)), mul_rn(v.y, f)); }
    static __device__ float2 diag_plain(float2 v, R f, R sb) { return make_float2(mul_rn(sub_rn(v.x, sb), f), mul_rn(v.y, f)); }
    static __device__ float2 new_diag_reordered(R f, R sb) { return make_float2(-mul_rn(sb, f), 0.f); }
    static __device__ float2 new_diag_plain(R f, R sb) { return make_float2(mul_rn(sub_rn(0.f, sb), f), mul_rn(0.f, f)); }
};
template<> struct Ops<double2> {
    using R = double;
    static __device__ double2 scale(double2 v, R f) { return make_double2(mul_rn(v.x, f), mul_rn(v.y, f)); }
    static __device__ double2 diag_reordered(double2 v, R f, R sb) { return make_double2(sub_rn(mul_rn(v.x, f), mul_rn(sb, f)), mul_rn(v.y, f)); }
    static __device__ double2 diag_plain(double2 v, R f, R sb) { return make_double2(mul_rn(sub_rn(v.x, sb), f), mul_rn(v.y, f)); }
    static __device__ double2 new_diag_reordered(R f, R sb) { return make_double2(-mul_rn(sb, f), 0.0); }
    static __device__ double2 new_diag_plain(R f, R sb) { return make_double2(mul_rn(sub_rn(0.0, sb), f), mul_rn(0.0, f)); }
};

/// widest row of the layout: the CSR row length, plus one where the b offset has to create a diagonal entry
__global__ void row_width_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n, int insert_diag,
                                 int* __restrict__ out) {
    int local = 0;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < n; row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int const b = indptr[row], e = indptr[row + 1];
        int cnt = e - b;
        if (insert_diag) {
            bool has = false;
            for (int p = b; p < e; ++p) has |= (indices[p] == row);
            if (!has) ++cnt;
        }
        local = max(local, cnt);
    }
    for (int o = 16; o > 0; o >>= 1) local = max(local, __shfl_xor_sync(0xffffffffu, local, o));
    if ((threadIdx.x & 31) == 0 && local > 0) atomicMax(out, local);
}

/// perm[queue[i]] = i   (original site -> row of the layout)
__global__ void invert_order_kernel(const int32_t* __restrict__ queue, int64_t n, int32_t* __restrict__ perm) {
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) perm[queue[i]] = static_cast<int32_t>(i);
}

#pragma nv_diag_suppress 549   // "c is used before its value is set": the insertion sort only reads entries it has written
template<class T>
__global__ void __launch_bounds__(128) csr_to_ell_kernel(BuildArgs a, const T* __restrict__ data, T* __restrict__ val) {
    using R = typename Ops<T>::R;
    int64_t const new_row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (new_row >= a.pitch) return;
    int32_t* __restrict__ col = a.col;
    if (new_row >= a.n) {   // tail rows of the pitch: zeros
        for (int s = 0; s < a.k; ++s) { val[s * a.pitch + new_row] = zero_(T{}); col[s * a.pitch + new_row] = 0; }
        return;
    }
    bool const reordered = a.queue != nullptr;
    int64_t const row = reordered ? a.queue[new_row] : new_row;
    R const f = static_cast<R>(a.f), sb = static_cast<R>(a.sb);
    bool const offset = a.mode == BUILD_SCALED && sb != R{0};
    int32_t c[BUILD_KMAX]; T v[BUILD_KMAX];   // filled from the front by `insert`; entry j - 1 is read only when j > 0
    int cnt = 0;
    bool diag_done = !offset;
    auto insert = [&](int32_t cc, T vv) {   // insertion sort by new column (rows are short)
        int j = cnt++;
        while (j > 0 && c[j - 1] > cc) { c[j] = c[j - 1]; v[j] = v[j - 1]; --j; }
        c[j] = cc; v[j] = vv;
    };
    for (int p = a.indptr[row]; p < a.indptr[row + 1]; ++p) {
        int32_t const cc = a.indices[p];
        T vv = data[p];
        if (a.mode == BUILD_SCALED) {
            if (offset && cc == row) { vv = reordered ? Ops<T>::diag_reordered(vv, f, sb) : Ops<T>::diag_plain(vv, f, sb); diag_done = true; }
            else vv = Ops<T>::scale(vv, f);
        } else if (a.mode == BUILD_VELOCITY) {
            vv = Ops<T>::scale(vv, static_cast<R>(sub_rn(a.positions[row], a.positions[cc])));   // H_ij * T(pos_i - pos_j), float difference
        }
        insert(reordered ? a.perm[cc] : cc, vv);
    }
    if (!diag_done) insert(static_cast<int32_t>(new_row), reordered ? Ops<T>::new_diag_reordered(f, sb) : Ops<T>::new_diag_plain(f, sb));
    for (int s = 0; s < a.k; ++s) {
        bool const in = s < cnt;
        val[s * a.pitch + new_row] = in ? v[s] : zero_(T{});
        col[s * a.pitch + new_row] = in ? c[s] : static_cast<int32_t>(new_row);
    }
}

} // anonymous namespace

int build_max_width() { return BUILD_KMAX; }

cudaError_t launch_row_width(const int32_t* indptr, const int32_t* indices, int64_t n, bool insert_diag, int* out_dev, cudaStream_t s) {
    cudaError_t err = cudaMemsetAsync(out_dev, 0, sizeof(int), s);
    if (err != cudaSuccess) return err;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    row_width_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(indptr, indices, n, insert_diag ? 1 : 0, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_invert_order(const int32_t* queue, int64_t n, int32_t* perm, cudaStream_t s) {
    invert_order_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(queue, n, perm);
    return cudaGetLastError();
}

cudaError_t launch_csr_to_ell(int dtype, BuildArgs const& a, const void* data, void* val, cudaStream_t s) {
    if (a.k > BUILD_KMAX) return cudaErrorInvalidValue;
    int const grid = static_cast<int>((a.pitch + 127) / 128);
    switch (dtype) {
        case F32: csr_to_ell_kernel<float><<<grid, 128, 0, s>>>(a, static_cast<const float*>(data), static_cast<float*>(val)); break;
        case C64: csr_to_ell_kernel<float2><<<grid, 128, 0, s>>>(a, static_cast<const float2*>(data), static_cast<float2*>(val)); break;
        case F64: csr_to_ell_kernel<double><<<grid, 128, 0, s>>>(a, static_cast<const double*>(data), static_cast<double*>(val)); break;
        case C128: csr_to_ell_kernel<double2><<<grid, 128, 0, s>>>(a, static_cast<const double2*>(data), static_cast<double2*>(val)); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace pbk
