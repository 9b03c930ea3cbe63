// kpm_oracle.cpp -- TEST INFRASTRUCTURE ONLY (the parity checker and the timed CPU baseline).
//
// A CPU restatement of pybinding's kernel-polynomial-method engine, written from the reference's
// algorithm description file by file (citations are relative to /root/reference/cppcore).  Nothing
// in the product (pybinding_b200/, libpbkpm.so) may include, link or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// Parity status: PINNED at curve level against the reference's own golden vectors
// (tests/baseline_data/kpm/*.pbz, copied as arrays into tests/golden/) and against the exact-integer
// known answers of cppcore/tests/test_kpm.cpp:33-171 (see tests/test_oracle_golden.py).  Raw-moment
// parity is pinned only through this restatement: the reference stores no raw-moment goldens.
//
// The reference itself cannot be compiled here (its CMake downloads Eigen/libsimdpp/mapbox/fmt),
// so every Eigen expression is restated as the plain loop it evaluates to.
//
// Two precision modes:
//   native : moments are accumulated and reconstructed in the scalar type of the Hamiltonian, like
//            the reference (f32 sums for f32 models).  Summation *order* inside SIMD reductions is
//            not reproduced (sequential sums are used instead).
//   hp     : vectors stay in the Hamiltonian's scalar type, but dot products, moment bookkeeping
//            and reconstruction run in double.  This is the arithmetic the GPU engine implements
//            and the mode the GPU parity tests compare against.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <xmmintrin.h>

namespace orc {

using cd = std::complex<double>;
using cf = std::complex<float>;

// numeric/constant.hpp:6-20 -- the reference's constants are single precision
constexpr float pi_f = 3.14159265358979323846f;
constexpr float kb_f = 8.6173303e-5f;

template<class T> struct traits;
template<> struct traits<float>  { using real_t = float;  using hp_t = double; static constexpr bool cplx = false; };
template<> struct traits<double> { using real_t = double; using hp_t = double; static constexpr bool cplx = false; };
template<> struct traits<cf>     { using real_t = float;  using hp_t = cd;     static constexpr bool cplx = true; };
template<> struct traits<cd>     { using real_t = double; using hp_t = cd;     static constexpr bool cplx = true; };
template<class T> using real_of = typename traits<T>::real_t;
template<class T> using hp_of = typename traits<T>::hp_t;

// compute/detail.hpp:15-42 -- raw complex multiply without inf/nan fix-ups
inline float mul(float a, float b) { return a * b; }
inline double mul(double a, double b) { return a * b; }
template<class R> inline std::complex<R> mul(std::complex<R> a, std::complex<R> b) {
    return {a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real()};
}
inline float conj_(float a) { return a; }
inline double conj_(double a) { return a; }
template<class R> inline std::complex<R> conj_(std::complex<R> a) { return {a.real(), -a.imag()}; }
inline float real_(float a) { return a; }
inline double real_(double a) { return a; }
template<class R> inline R real_(std::complex<R> a) { return a.real(); }

template<class A, class T> inline A cast_to(T v) { return static_cast<A>(v); }
template<> inline float cast_to<float, cd>(cd v) { return static_cast<float>(v.real()); }
template<> inline double cast_to<double, cd>(cd v) { return v.real(); }
template<> inline cf cast_to<cf, cd>(cd v) { return cf(static_cast<float>(v.real()), static_cast<float>(v.imag())); }
template<> inline cd cast_to<cd, cf>(cf v) { return cd(v.real(), v.imag()); }
template<> inline cd cast_to<cd, float>(float v) { return cd(v, 0); }
template<> inline cd cast_to<cd, double>(double v) { return cd(v, 0); }
template<> inline cf cast_to<cf, float>(float v) { return cf(v, 0); }
template<> inline double cast_to<double, cf>(cf v) { return v.real(); }
template<> inline float cast_to<float, cf>(cf v) { return v.real(); }

inline cd to_cd(float v) { return cd(v, 0); }
inline cd to_cd(double v) { return cd(v, 0); }
inline cd to_cd(cf v) { return cd(v.real(), v.imag()); }
inline cd to_cd(cd v) { return v; }

/// |a|^2 accumulated in type A (A is T or the double-precision counterpart of T)
template<class A, class T> inline A square_as(T a) {
    auto const x = cast_to<A>(a);
    return cast_to<A>(real_(mul(conj_(x), x)));
}
/// conj(a) * b accumulated in type A
template<class A, class T> inline A cdot_as(T a, T b) { return mul(conj_(cast_to<A>(a)), cast_to<A>(b)); }

struct ftz_guard {  // support/simd.hpp:140-143
    ftz_guard() { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); }
    ~ftz_guard() { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_OFF); }
};

// ------------------------------------------------------------------------------------------------
// Sparse containers (numeric/sparse.hpp, numeric/ellmatrix.hpp)
// ------------------------------------------------------------------------------------------------
template<class T> struct Csr {
    int rows = 0;
    std::vector<int> indptr{0};
    std::vector<int> indices;
    std::vector<T> data;
    bool empty() const { return rows == 0; }
    int nnz() const { return static_cast<int>(indices.size()); }
    int max_nnz_per_row() const {
        int m = 0;
        for (int r = 0; r < rows; ++r) m = std::max(m, indptr[r + 1] - indptr[r]);
        return m;
    }
};

/// y = A * x with Eigen's row-major sparse * dense evaluation order (sequential over the row)
template<class T> std::vector<T> spmv_plain(Csr<T> const& a, std::vector<T> const& x) {
    std::vector<T> y(a.rows);
    for (int row = 0; row < a.rows; ++row) {
        T tmp{0};
        for (int n = a.indptr[row]; n < a.indptr[row + 1]; ++n) tmp += a.data[n] * x[a.indices[n]];
        y[row] = tmp;
    }
    return y;
}

template<class T> struct Ell {  // slot-major: element (row, n) lives at n * pitch + row
    int rows = 0, k = 0, pitch = 0;
    std::vector<T> data;
    std::vector<int> idx;
    T const& d(int row, int n) const { return data[static_cast<size_t>(n) * pitch + row]; }
    int i(int row, int n) const { return idx[static_cast<size_t>(n) * pitch + row]; }
};

/// numeric/ellmatrix.hpp:65-82 -- pad with value 0 and the previous row's column index
template<class T> Ell<T> csr_to_ell(Csr<T> const& csr) {
    Ell<T> ell;
    ell.rows = csr.rows;
    ell.k = csr.max_nnz_per_row();
    ell.pitch = (csr.rows + 7) / 8 * 8;
    ell.data.assign(static_cast<size_t>(ell.pitch) * ell.k, T{0});
    ell.idx.assign(static_cast<size_t>(ell.pitch) * ell.k, 0);
    for (int row = 0; row < csr.rows; ++row) {
        int n = 0;
        for (int p = csr.indptr[row]; p < csr.indptr[row + 1]; ++p, ++n) {
            ell.data[static_cast<size_t>(n) * ell.pitch + row] = csr.data[p];
            ell.idx[static_cast<size_t>(n) * ell.pitch + row] = csr.indices[p];
        }
        for (; n < ell.k; ++n) {
            ell.idx[static_cast<size_t>(n) * ell.pitch + row] = (row > 0) ? ell.i(row - 1, n) : 0;
        }
    }
    return ell;
}

// ------------------------------------------------------------------------------------------------
// Scale (kpm/Bounds.hpp:11-33)
// ------------------------------------------------------------------------------------------------
struct ScaleD {
    double a = 0, b = 0;
    ScaleD() = default;
    ScaleD(double min_energy, double max_energy) {
        constexpr auto tolerance = 0.01f;
        a = 0.5f * (max_energy - min_energy) * (1 + tolerance);
        b = 0.5f * (max_energy + min_energy);
        if (std::abs(b / a) < 0.01f * tolerance) { b = 0; }
    }
};
template<class R> struct Scale {
    R a, b;
    explicit Scale(ScaleD s) : a(static_cast<R>(s.a)), b(static_cast<R>(s.b)) {}
};

// ------------------------------------------------------------------------------------------------
// Kernels (kpm/Kernel.hpp:9-57, src/kpm/Kernel.cpp:6-49)
// ------------------------------------------------------------------------------------------------
inline int round_num_moments(int n) {
    if (n < 2) { return 2; }
    while ((n - 2) % 4 != 0) { ++n; }
    return n;
}

struct Kernel {
    int kind = 0;  // 0 jackson, 1 lorentz, 2 dirichlet
    double lambda = 4.0;

    std::vector<double> damping(int num_moments) const {
        std::vector<double> g(num_moments);
        auto const N = static_cast<double>(num_moments);
        if (kind == 0) {
            auto const Np = N + 1;
            constexpr auto pi = double{pi_f};
            for (int i = 0; i < num_moments; ++i) {
                auto const n = static_cast<double>(i);
                g[i] = ((Np - n) * std::cos(pi * n / Np) + std::sin(pi * n / Np) / std::tan(pi / Np)) / Np;
            }
        } else if (kind == 1) {
            for (int i = 0; i < num_moments; ++i) {
                auto const n = static_cast<double>(i);
                g[i] = std::sinh(lambda * (1 - n / N)) / std::sinh(lambda);
            }
        } else {
            std::fill(g.begin(), g.end(), 1.0);
        }
        return g;
    }

    int required_num_moments(double scaled_broadening) const {
        auto const num = (kind == 1) ? lambda : static_cast<double>(pi_f);
        return round_num_moments(static_cast<int>(num / scaled_broadening) + 1);
    }
};

// ------------------------------------------------------------------------------------------------
// Slice map + optimized Hamiltonian (kpm/OptimizedHamiltonian.hpp:54-96, src/...cpp:5-152)
// ------------------------------------------------------------------------------------------------
struct SliceMap {
    std::vector<int> data;
    int src_offset = 0, dest_offset = 0;

    int last_index() const { return static_cast<int>(data.size()) - 1; }
    int index(int n, int num_moments) const {
        auto const mid = (num_moments - 1 + dest_offset - src_offset) / 2;
        auto const max = std::min(last_index(), mid + src_offset);
        if (n < mid) { return std::min(max, n + src_offset); }
        return std::min(max, num_moments - 1 - n + dest_offset);
    }
    int optimal_size(int n, int num_moments) const { return data[index(n, num_moments)]; }
    bool uses_full_system(int num_moments) const { return static_cast<int>(data.size()) < num_moments / 2; }
};

struct Indices {
    std::vector<int> src, dest;
    bool is_diagonal() const { return src == dest; }
    bool operator==(Indices const& o) const { return src == o.src && dest == o.dest; }
};

template<class T> struct OptimizedHamiltonian {
    Csr<T> csr;
    Ell<T> ell;
    bool use_ell = true;
    bool is_reordered = true;
    bool valid = false;
    Indices original_idx, idx;
    SliceMap map;
    std::vector<int> reorder_map;
    double seconds = 0;

    int size() const { return csr.rows; }

    template<class V> void reorder(std::vector<V>& v) const {
        if (reorder_map.empty()) { return; }
        std::vector<V> out(v.size());
        for (size_t i = 0; i < v.size(); ++i) out[reorder_map[i]] = v[i];
        v.swap(out);
    }

    void reorder(Csr<T>& m) const {  // OptimizedHamiltonian.hpp:140-155 (rows end up column-sorted)
        if (reorder_map.empty() || m.empty()) { return; }
        int const n = m.rows;
        std::vector<int> inv(n);
        for (int i = 0; i < n; ++i) inv[reorder_map[i]] = i;
        Csr<T> out;
        out.rows = n;
        out.indptr.assign(n + 1, 0);
        out.indices.reserve(m.indices.size());
        out.data.reserve(m.data.size());
        std::vector<std::pair<int, T>> row_buf;
        for (int new_row = 0; new_row < n; ++new_row) {
            int const row = inv[new_row];
            row_buf.clear();
            for (int p = m.indptr[row]; p < m.indptr[row + 1]; ++p) {
                row_buf.emplace_back(reorder_map[m.indices[p]], m.data[p]);
            }
            std::sort(row_buf.begin(), row_buf.end(), [](auto const& l, auto const& r) { return l.first < r.first; });
            for (auto const& e : row_buf) { out.indices.push_back(e.first); out.data.push_back(e.second); }
            out.indptr[new_row + 1] = static_cast<int>(out.indices.size());
        }
        m = std::move(out);
    }

    void optimize_for(Csr<T> const& h, Indices const& target, ScaleD s) {
        if (valid && original_idx == target) { return; }
        auto const t0 = std::chrono::steady_clock::now();
        if (is_reordered) { create_reordered(h, target, s); } else { create_scaled(h, target, s); }
        if (use_ell) { ell = csr_to_ell(csr); }
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        original_idx = target;
        valid = true;
    }

    /// src/kpm/OptimizedHamiltonian.cpp:55-73 : H2 = (H - I*b) * (2/a), union sparsity pattern
    void create_scaled(Csr<T> const& h, Indices const& target, ScaleD s) {
        using R = real_of<T>;
        auto const scale = Scale<R>(s);
        R const f = 2 / scale.a;
        csr = Csr<T>();
        csr.rows = h.rows;
        csr.indptr.assign(h.rows + 1, 0);
        for (int row = 0; row < h.rows; ++row) {
            bool diagonal_done = (scale.b == 0);
            for (int p = h.indptr[row]; p < h.indptr[row + 1]; ++p) {
                int const col = h.indices[p];
                if (!diagonal_done && col > row) {
                    csr.indices.push_back(row);
                    csr.data.push_back((T{0} - T{scale.b}) * f);
                    diagonal_done = true;
                }
                T v = h.data[p];
                if (!diagonal_done && col == row) { v = v - T{scale.b}; diagonal_done = true; }
                csr.indices.push_back(col);
                csr.data.push_back(v * f);
            }
            if (!diagonal_done) {
                csr.indices.push_back(row);
                csr.data.push_back((T{0} - T{scale.b}) * f);
            }
            csr.indptr[row + 1] = static_cast<int>(csr.indices.size());
        }
        idx = target;
        reorder_map.clear();
        map = SliceMap{{h.rows}, 0, 0};
    }

    /// src/kpm/OptimizedHamiltonian.cpp:75-152 : BFS relabel from src[0], scale, record slice borders
    void create_reordered(Csr<T> const& h, Indices const& target, ScaleD s) {
        using R = real_of<T>;
        auto const scale = Scale<R>(s);
        int const system_size = h.rows;
        auto const inverted_a = R{2 / scale.a};

        std::vector<int> index_queue;
        index_queue.reserve(system_size);
        index_queue.push_back(target.src[0]);
        reorder_map.assign(system_size, -1);
        reorder_map[target.src[0]] = 0;
        std::vector<int> borders{1};

        csr = Csr<T>();
        csr.rows = system_size;
        csr.indptr.assign(system_size + 1, 0);
        std::vector<std::pair<int, T>> row_buf;
        for (int h2_row = 0; h2_row < system_size; ++h2_row) {
            if (h2_row >= static_cast<int>(index_queue.size())) {
                throw std::runtime_error("oracle: the Hamiltonian graph is not connected (the reference "
                                         "reads past its index queue here)");
            }
            bool diagonal_inserted = false;
            int const row = index_queue[h2_row];
            row_buf.clear();
            for (int p = h.indptr[row]; p < h.indptr[row + 1]; ++p) {
                int const col = h.indices[p];
                if (reorder_map[col] < 0) {
                    reorder_map[col] = static_cast<int>(index_queue.size());
                    index_queue.push_back(col);
                }
                T v = h.data[p] * inverted_a;
                if (row == col) { v -= scale.b * inverted_a; diagonal_inserted = true; }
                row_buf.emplace_back(reorder_map[col], v);
            }
            if (scale.b != 0 && !diagonal_inserted) { row_buf.emplace_back(h2_row, T{-scale.b * inverted_a}); }
            std::sort(row_buf.begin(), row_buf.end(), [](auto const& l, auto const& r) { return l.first < r.first; });
            for (auto const& e : row_buf) { csr.indices.push_back(e.first); csr.data.push_back(e.second); }
            csr.indptr[h2_row + 1] = static_cast<int>(csr.indices.size());

            if (h2_row == borders.back() - 1) { borders.push_back(static_cast<int>(index_queue.size())); }
        }
        borders.pop_back();

        idx.src.clear(); idx.dest.clear();
        for (int i : target.src) idx.src.push_back(reorder_map[i]);
        for (int i : target.dest) idx.dest.push_back(reorder_map[i]);

        auto find_offset = [&](std::vector<int> const& v) {  // SliceMap ctor, :5-18
            int const max_index = *std::max_element(v.begin(), v.end());
            auto const it = std::find_if(borders.begin(), borders.end(), [&](int b) { return b > max_index; });
            return static_cast<int>(it - borders.begin());
        };
        map.src_offset = find_offset(idx.src);
        map.dest_offset = find_offset(idx.dest);
        map.data = std::move(borders);
    }

    // Stats helpers (src/kpm/OptimizedHamiltonian.cpp:154-207)
    size_t nnz_upto(int rows) const {
        return use_ell ? static_cast<size_t>(rows) * ell.k : static_cast<size_t>(csr.indptr[rows]);
    }
    size_t num_nonzeros(int num_moments, bool optimal_size) const {
        size_t result = 0;
        if (!optimal_size) {
            result = static_cast<size_t>(num_moments) * nnz_upto(size());
        } else {
            for (int n = 0; n < num_moments; ++n) result += nnz_upto(map.optimal_size(n, num_moments));
        }
        if (idx.is_diagonal()) { result /= 2; }
        return result;
    }
};

// ------------------------------------------------------------------------------------------------
// KPM SpMV kernels (compute/kernel_polynomial.hpp:19-326).  B = number of interleaved vectors
// (row-major N x B block, like the reference's MatrixX batches); B == 1 is the plain vector case.
// ------------------------------------------------------------------------------------------------
template<class T, int B>
void kpm_spmv(int start, int end, Csr<T> const& m, T const* x, T* y) {
    for (int row = start; row < end; ++row) {
        T r[B];
        for (int j = 0; j < B; ++j) r[j] = T{0};
        for (int n = m.indptr[row]; n < m.indptr[row + 1]; ++n) {
            T const a = m.data[n];
            T const* xr = x + static_cast<size_t>(m.indices[n]) * B;
            for (int j = 0; j < B; ++j) r[j] += mul(a, xr[j]);
        }
        T* yr = y + static_cast<size_t>(row) * B;
        for (int j = 0; j < B; ++j) yr[j] = r[j] - yr[j];
    }
}

template<class T, int B>
void kpm_spmv(int start, int end, Ell<T> const& m, T const* x, T* y, int skip_last_n = 0) {
    for (int n = 0; n < m.k - skip_last_n; ++n) {
        T const* data = m.data.data() + static_cast<size_t>(n) * m.pitch;
        int const* idx = m.idx.data() + static_cast<size_t>(n) * m.pitch;
        for (int row = start; row < end; ++row) {
            T const a = data[row];
            T const* xr = x + static_cast<size_t>(idx[row]) * B;
            T* yr = y + static_cast<size_t>(row) * B;
            if (n == 0) {
                for (int j = 0; j < B; ++j) yr[j] = mul(a, xr[j]) - yr[j];
            } else {
                for (int j = 0; j < B; ++j) yr[j] = mul(a, xr[j]) + yr[j];
            }
        }
    }
}

template<class T, class A, int B>
void kpm_spmv_diagonal(int start, int end, Csr<T> const& m, T const* x, T* y, A* m2, A* m3) {
    kpm_spmv<T, B>(start, end, m, x, y);
    for (int row = start; row < end; ++row) {
        T const* xr = x + static_cast<size_t>(row) * B;
        T const* yr = y + static_cast<size_t>(row) * B;
        for (int j = 0; j < B; ++j) {
            m2[j] += square_as<A>(xr[j]);
            m3[j] += cdot_as<A>(yr[j], xr[j]);
        }
    }
}

/// ELL + diagonal: the last slot is fused with the two sums (kernel_polynomial.hpp:221-323)
template<class T, class A, int B>
void kpm_spmv_diagonal(int start, int end, Ell<T> const& m, T const* x, T* y, A* m2, A* m3) {
    kpm_spmv<T, B>(start, end, m, x, y, 1);
    int const n = m.k - 1;
    T const* data = m.data.data() + static_cast<size_t>(n) * m.pitch;
    int const* idx = m.idx.data() + static_cast<size_t>(n) * m.pitch;
    for (int row = start; row < end; ++row) {
        T const a = data[row];
        T const* xb = x + static_cast<size_t>(idx[row]) * B;
        T const* xr = x + static_cast<size_t>(row) * B;
        T* yr = y + static_cast<size_t>(row) * B;
        for (int j = 0; j < B; ++j) {
            T const c = (n == 0) ? -yr[j] : yr[j];
            T const r2 = mul(a, xb[j]) + c;
            m2[j] += square_as<A>(xr[j]);
            m3[j] += cdot_as<A>(r2, xr[j]);
            yr[j] = r2;
        }
    }
}

// r1 = 0.5 * h2 * r0  (kpm/Starter.hpp:55-118)
template<class T, int B> std::vector<T> make_r1(Csr<T> const& h2, std::vector<T> const& r0) {
    std::vector<T> r1(r0.size());
    for (int row = 0; row < h2.rows; ++row) {
        T tmp[B];
        for (int j = 0; j < B; ++j) tmp[j] = T{0};
        for (int n = h2.indptr[row]; n < h2.indptr[row + 1]; ++n) {
            for (int j = 0; j < B; ++j) tmp[j] += mul(h2.data[n], r0[static_cast<size_t>(h2.indices[n]) * B + j]);
        }
        for (int j = 0; j < B; ++j) r1[static_cast<size_t>(row) * B + j] = tmp[j] * T{0.5};
    }
    return r1;
}
template<class T, int B> std::vector<T> make_r1(Ell<T> const& h2, std::vector<T> const& r0) {
    std::vector<T> r1(r0.size(), T{0});
    for (int n = 0; n < h2.k; ++n) {
        for (int row = 0; row < h2.rows; ++row) {
            T const a = h2.d(row, n);
            size_t const c = static_cast<size_t>(h2.i(row, n)) * B;
            for (int j = 0; j < B; ++j) r1[static_cast<size_t>(row) * B + j] += mul(a, r0[c + j]) * T{0.5};
        }
    }
    return r1;
}

// ------------------------------------------------------------------------------------------------
// Collectors (src/kpm/default/collectors.cpp:6-106)
// ------------------------------------------------------------------------------------------------
/// Timing aid of bench.py's cpu_baseline (not part of the reference): wall time a job spends between recursion steps
/// n1 and n2, so that a bounded sample reports the asymptotic cost per moment without the per-vector fixed cost
/// (starter generation, allocation and first touch of the vector blocks, r1 = H r0 / 2).
struct StepProbe {
    int n1 = 0, n2 = 0;
    std::mutex mutex;
    double max_elapsed = 0;   // slowest job: all jobs of one wave run concurrently
    int reports = 0;
    void report(double seconds) { std::lock_guard<std::mutex> lk(mutex); max_elapsed = std::max(max_elapsed, seconds); ++reports; }
};

template<class T, class A, int B> struct DiagonalCollector {  // B == 1: DiagonalCollector, else Batch
    static constexpr bool diagonal = true;
    StepProbe* probe = nullptr;
    std::chrono::steady_clock::time_point probe_t0;
    int num_moments;
    std::vector<A> moments;  // num_moments x B, row-major
    A m0[B], m1[B];
    explicit DiagonalCollector(int n) : num_moments(n), moments(static_cast<size_t>(n) * B) {}
    int size() const { return num_moments; }
    void initial(std::vector<T> const& r0, std::vector<T> const& r1) {
        size_t const rows = r0.size() / B;
        for (int j = 0; j < B; ++j) {
            A s0{0}, s1{0};
            for (size_t i = 0; i < rows; ++i) {
                s0 += square_as<A>(r0[i * B + j]);
                s1 += cdot_as<A>(r1[i * B + j], r0[i * B + j]);
            }
            moments[0 * B + j] = m0[j] = s0 * A{0.5};
            moments[1 * B + j] = m1[j] = s1;
        }
    }
    void operator()(int n, A const* m2, A const* m3) {
        if (probe) {
            if (n == probe->n1) probe_t0 = std::chrono::steady_clock::now();
            else if (n == probe->n2) probe->report(std::chrono::duration<double>(std::chrono::steady_clock::now() - probe_t0).count());
        }
        for (int j = 0; j < B; ++j) {
            moments[static_cast<size_t>(2 * (n - 1)) * B + j] = A{2} * (m2[j] - m0[j]);
            moments[static_cast<size_t>(2 * (n - 1) + 1) * B + j] = A{2} * m3[j] - m1[j];
        }
    }
};

template<class T> struct OffDiagonalCollector {
    static constexpr bool diagonal = false;
    virtual ~OffDiagonalCollector() = default;
    virtual int size() const = 0;
    virtual void initial(std::vector<T> const& r0, std::vector<T> const& r1) = 0;
    virtual void operator()(int n, std::vector<T> const& r1) = 0;
};

template<class T, class A> struct GenericCollector : OffDiagonalCollector<T> {
    std::vector<A> moments;
    std::vector<T> beta;
    Csr<T> op;
    GenericCollector(int n, OptimizedHamiltonian<T> const& oh, std::vector<T> beta_, Csr<T> op_)
        : moments(n), beta(std::move(beta_)), op(std::move(op_)) {
        oh.reorder(beta);
        oh.reorder(op);
    }
    int size() const override { return static_cast<int>(moments.size()); }
    A expval(std::vector<T> const& r) const {
        A s{0};
        if (!op.empty()) {
            auto const v = spmv_plain(op, r);
            for (size_t i = 0; i < v.size(); ++i) s += cdot_as<A>(beta[i], v[i]);
        } else {
            for (size_t i = 0; i < r.size(); ++i) s += cdot_as<A>(beta[i], r[i]);
        }
        return s;
    }
    void initial(std::vector<T> const& r0, std::vector<T> const& r1) override {
        moments[0] = expval(r0) * A{0.5};
        moments[1] = expval(r1);
    }
    void operator()(int n, std::vector<T> const& r1) override { moments[n] = expval(r1); }
};

template<class T, class A> struct MultiUnitCollector : OffDiagonalCollector<T> {
    Indices const& idx;
    int num_moments;
    std::vector<std::vector<A>> moments;
    MultiUnitCollector(int n, Indices const& idx)
        : idx(idx), num_moments(n), moments(idx.dest.size(), std::vector<A>(n)) {}
    int size() const override { return num_moments; }
    void initial(std::vector<T> const& r0, std::vector<T> const& r1) override {
        for (size_t i = 0; i < idx.dest.size(); ++i) {
            moments[i][0] = cast_to<A>(r0[idx.dest[i]] * real_of<T>{0.5});
            moments[i][1] = cast_to<A>(r1[idx.dest[i]]);
        }
    }
    void operator()(int n, std::vector<T> const& r1) override {
        for (size_t i = 0; i < idx.dest.size(); ++i) moments[i][n] = cast_to<A>(r1[idx.dest[i]]);
    }
};

template<class T> struct DenseMatrixCollector : OffDiagonalCollector<T> {
    Csr<T> op;
    int num_moments, n_rows;
    std::vector<T> moments;  // num_moments x N, row-major
    DenseMatrixCollector(int n, OptimizedHamiltonian<T> const& oh, Csr<T> op_)
        : op(std::move(op_)), num_moments(n), n_rows(oh.size()), moments(static_cast<size_t>(n) * oh.size()) {
        oh.reorder(op);
    }
    int size() const override { return num_moments; }
    void store(int n, std::vector<T> const& r, bool half) {
        T* out = moments.data() + static_cast<size_t>(n) * n_rows;
        if (!op.empty()) {
            auto const v = spmv_plain(op, r);
            for (int i = 0; i < n_rows; ++i) out[i] = half ? v[i] * real_of<T>{0.5} : v[i];
        } else {
            for (int i = 0; i < n_rows; ++i) out[i] = half ? r[i] * real_of<T>{0.5} : r[i];
        }
    }
    void initial(std::vector<T> const& r0, std::vector<T> const& r1) override { store(0, r0, true); store(1, r1, false); }
    void operator()(int n, std::vector<T> const& r1) override { store(n, r1, false); }
};

// ------------------------------------------------------------------------------------------------
// Recursion drivers (kpm/calc_moments.hpp:36-146)
// ------------------------------------------------------------------------------------------------
template<class T, class A, int B, class Matrix>
void diagonal_basic(DiagonalCollector<T, A, B>& collect, std::vector<T> r0, std::vector<T> r1,
                    Matrix const& h2, SliceMap const& map, bool opt_size) {
    int const num_moments = collect.size();
    for (int n = 2; n <= num_moments / 2; ++n) {
        A m2[B], m3[B];
        for (int j = 0; j < B; ++j) { m2[j] = A{0}; m3[j] = A{0}; }
        int const size = opt_size ? map.optimal_size(n, num_moments) : h2.rows;
        kpm_spmv_diagonal<T, A, B>(0, size, h2, r1.data(), r0.data(), m2, m3);
        collect(n, m2, m3);
        r1.swap(r0);
    }
}

template<class T, class A, int B, class Matrix>
void diagonal_interleaved(DiagonalCollector<T, A, B>& collect, std::vector<T> r0, std::vector<T> r1,
                          Matrix const& h2, SliceMap const& map, bool opt_size) {
    int const num_moments = collect.size();
    for (int n = 2; n <= num_moments / 2; n += 2) {
        A m2[B], m3[B], m4[B], m5[B];
        for (int j = 0; j < B; ++j) { m2[j] = m3[j] = m4[j] = m5[j] = A{0}; }
        int const max1 = opt_size ? map.index(n, num_moments) : map.last_index();
        int const max2 = opt_size ? map.index(n + 1, num_moments) : map.last_index();
        for (int k = 0, start0 = 0, start1 = 0; k <= max1; ++k) {
            int const end0 = map.data[k];
            int const end1 = (k == max1) ? map.data[max2] : start0;
            kpm_spmv_diagonal<T, A, B>(start0, end0, h2, r1.data(), r0.data(), m2, m3);
            kpm_spmv_diagonal<T, A, B>(start1, end1, h2, r0.data(), r1.data(), m4, m5);
            start1 = end1;
            start0 = end0;
        }
        collect(n, m2, m3);
        collect(n + 1, m4, m5);
    }
}

template<class T, class Matrix>
void offdiagonal_basic(OffDiagonalCollector<T>& collect, std::vector<T> r0, std::vector<T> r1,
                       Matrix const& h2, SliceMap const& map, bool opt_size) {
    int const num_moments = collect.size();
    for (int n = 2; n < num_moments; ++n) {
        int const size = opt_size ? map.optimal_size(n, num_moments) : h2.rows;
        kpm_spmv<T, 1>(0, size, h2, r1.data(), r0.data());
        r1.swap(r0);
        collect(n, r1);
    }
}

template<class T, class Matrix>
void offdiagonal_interleaved(OffDiagonalCollector<T>& collect, std::vector<T> r0, std::vector<T> r1,
                             Matrix const& h2, SliceMap const& map, bool opt_size) {
    int const num_moments = collect.size();
    for (int n = 2; n < num_moments; n += 2) {
        int const max1 = opt_size ? map.index(n, num_moments) : map.last_index();
        int const max2 = opt_size ? map.index(n + 1, num_moments) : map.last_index();
        for (int k = 0, start0 = 0, start1 = 0; k <= max1; ++k) {
            int const end0 = map.data[k];
            int const end1 = (k == max1) ? map.data[max2] : start0;
            kpm_spmv<T, 1>(start0, end0, h2, r1.data(), r0.data());
            kpm_spmv<T, 1>(start1, end1, h2, r0.data(), r1.data());
            start1 = start0;
            start0 = end0;
        }
        collect(n, r0);
        collect(n + 1, r1);
    }
}

// ------------------------------------------------------------------------------------------------
// Starters (src/kpm/Starter.cpp:9-99, numeric/random.hpp:25-50)
// ------------------------------------------------------------------------------------------------
template<class T> struct Starter {
    std::function<std::vector<T>()> make;
    // Optional split of `make` for starters that consume a shared random stream: `draw` takes the next vector's worth
    // of uniform reals and must run under the mutex (the stream order defines which vector gets which numbers, exactly as
    // in the reference where all of `make` runs under Starter::mutex, Starter.hpp:18-21); `finish` turns them into the
    // starter vector and may run concurrently.  Same values as `make`, only less serialised -- an oracle-side speed-up
    // for checks at benchmark size (tens of millions of sites x 64 vectors).
    std::function<std::vector<real_of<T>>()> draw;
    std::function<std::vector<T>(std::vector<real_of<T>> const&)> finish;
    int vector_size = 0;
    int count = 0;
    std::unique_ptr<std::mutex> mutex = std::make_unique<std::mutex>();
};

template<class R> std::vector<R> make_random_real(int size, std::mt19937& generator) {
    std::uniform_real_distribution<R> distribution;
    std::vector<R> v(size);
    for (auto& value : v) value = distribution(generator);
    return v;
}

template<class T> struct RandomStarterFn {
    OptimizedHamiltonian<T> const* oh;
    Csr<T> op;  // applied in the original ordering, before the reorder
    std::shared_ptr<std::mt19937> generator = std::make_shared<std::mt19937>();

    std::vector<T> operator()() { return finish(draw()); }
    std::vector<real_of<T>> draw() { return make_random_real<real_of<T>>(oh->size(), *generator); }
    std::vector<T> finish(std::vector<real_of<T>> const& u) const {
        using R = real_of<T>;
        std::vector<T> r0(oh->size());
        if constexpr (!traits<T>::cplx) {
            for (size_t i = 0; i < r0.size(); ++i) r0[i] = (u[i] < 0.5f) ? R{-1.f} : R{1.f};
        } else {
            auto const k = std::complex<R>{2 * pi_f * cf(0, 1)};
            for (size_t i = 0; i < r0.size(); ++i) r0[i] = std::exp(k * u[i]);
        }
        if (!op.empty()) { r0 = spmv_plain(op, r0); }
        oh->reorder(r0);
        return r0;
    }
};

/// make_r0 for split starters: the draws happen under the caller's lock, the rest outside (see Starter::draw)
template<class T, int B> std::vector<std::vector<real_of<T>>> draw_r0(Starter<T>& starter) {
    starter.count += B;
    std::vector<std::vector<real_of<T>>> u(B);
    for (int j = 0; j < B; ++j) u[j] = starter.draw();
    return u;
}
template<class T, int B> std::vector<T> finish_r0(Starter<T> const& starter, std::vector<std::vector<real_of<T>>>& u) {
    if (B == 1) return starter.finish(u[0]);
    std::vector<T> r0(static_cast<size_t>(starter.vector_size) * B);
    for (int j = 0; j < B; ++j) {
        auto const col = starter.finish(u[j]);
        std::vector<real_of<T>>().swap(u[j]);
        for (int i = 0; i < starter.vector_size; ++i) r0[static_cast<size_t>(i) * B + j] = col[i];
    }
    return r0;
}

template<class T, int B> std::vector<T> make_r0(Starter<T>& starter) {
    if (B == 1) { ++starter.count; return starter.make(); }
    starter.count += B;
    std::vector<T> r0(static_cast<size_t>(starter.vector_size) * B);
    for (int j = 0; j < B; ++j) {
        auto const col = starter.make();
        for (int i = 0; i < starter.vector_size; ++i) r0[static_cast<size_t>(i) * B + j] = col[i];
    }
    return r0;
}

// ------------------------------------------------------------------------------------------------
// Lanczos bounds (compute/lanczos.hpp:26-154, compute/eigen3/lanczos.hpp:7-90)
// ------------------------------------------------------------------------------------------------
template<class R> void make_givens(R p, R q, R& c, R& s) {  // Eigen::JacobiRotation::makeGivens, real case
    if (q == R{0}) { c = p < R{0} ? R{-1} : R{1}; s = R{0}; }
    else if (p == R{0}) { c = R{0}; s = q < R{0} ? R{1} : R{-1}; }
    else if (std::abs(p) > std::abs(q)) {
        R const t = q / p;
        R u = std::sqrt(R{1} + t * t);
        if (p < R{0}) { u = -u; }
        c = R{1} / u; s = -t * c;
    } else {
        R const t = p / q;
        R u = std::sqrt(R{1} + t * t);
        if (q < R{0}) { u = -u; }
        s = -R{1} / u; c = -t * s;
    }
}

template<class R> void tridiagonal_qr_step(R* diag, R* subdiag, int start, int end) {
    auto td = (diag[end - 1] - diag[end]) * R(0.5);
    auto e = subdiag[end - 1];
    auto mu = diag[end];
    if (td == 0) {
        mu -= std::abs(e);
    } else {
        auto e2 = subdiag[end - 1] * subdiag[end - 1];
        auto h = std::hypot(td, e);
        if (e2 == 0) { mu -= (e / (td + (td > 0 ? 1 : -1))) * (e / h); }
        else { mu -= e2 / (td + (td > 0 ? h : -h)); }
    }
    auto x = diag[start] - mu;
    auto z = subdiag[start];
    for (auto k = start; k < end; ++k) {
        R c, s;
        make_givens(x, z, c, s);
        auto sdk = s * diag[k] + c * subdiag[k];
        auto dkp1 = s * subdiag[k] + c * diag[k + 1];
        diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
        diag[k + 1] = s * sdk + c * dkp1;
        subdiag[k] = c * sdk - s * dkp1;
        if (k > start) { subdiag[k - 1] = c * subdiag[k - 1] - s * z; }
        x = subdiag[k];
        if (k < end - 1) {
            z = -s * subdiag[k + 1];
            subdiag[k + 1] = c * subdiag[k + 1];
        }
    }
}

template<class R> std::vector<R> tridiagonal_eigenvalues(std::vector<R> const& alpha, std::vector<R> const& beta) {
    std::vector<R> eigenvalues = alpha;
    std::vector<R> temp = beta;
    int start = 0;
    int end = static_cast<int>(eigenvalues.size()) - 1;
    int iter = 0;
    constexpr int max_iterations = 30;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            auto a = std::abs(temp[i]);
            auto b = std::abs(eigenvalues[i]) + std::abs(eigenvalues[i + 1]);
            if (a < b * std::numeric_limits<R>::epsilon()) { temp[i] = 0; }
        }
        while (end > 0 && temp[end - 1] == 0) { end--; }
        if (end <= 0) { break; }
        if (++iter > max_iterations * static_cast<int>(eigenvalues.size())) { throw std::runtime_error{"Tridiagonal QR error"}; }
        start = end - 1;
        while (start > 0 && temp[start - 1] != 0) { start--; }
        tridiagonal_qr_step(eigenvalues.data(), temp.data(), start, end);
    }
    return eigenvalues;
}

struct LanczosBounds { double min, max; int loops; };

template<class T> LanczosBounds minmax_eigenvalues(Csr<T> const& matrix, double precision_percent) {
    using R = real_of<T>;
    ftz_guard guard;
    int const size = matrix.rows;
    std::vector<T> v0(size, T{0}), v1(size);
    {
        std::mt19937 generator;
        auto const u = make_random_real<R>(size, generator);
        R norm2 = 0;
        for (int i = 0; i < size; ++i) { v1[i] = T{u[i]}; norm2 += u[i] * u[i]; }
        R const norm = std::sqrt(norm2);
        for (auto& v : v1) v /= norm;
    }
    std::vector<R> alpha, beta;
    auto previous_min = std::numeric_limits<R>::max();
    auto previous_max = std::numeric_limits<R>::lowest();
    auto const precision = static_cast<R>(precision_percent / 100);

    for (int i = 0; i < 1000; ++i) {
        auto const b_prev = !beta.empty() ? beta.back() : R{0};
        R a{0};
        for (int row = 0; row < size; ++row) {  // lanczos_spmv
            T tmp{0};
            for (int n = matrix.indptr[row]; n < matrix.indptr[row + 1]; ++n) tmp += mul(matrix.data[n], v1[matrix.indices[n]]);
            v0[row] = tmp - b_prev * v0[row];
            a += real_(mul(conj_(tmp), v1[row]));
        }
        R norm2{0};
        for (int k = 0; k < size; ++k) {  // lanczos_axpy
            auto const l = v0[k] - a * v1[k];
            norm2 += real_(mul(conj_(l), l));
            v0[k] = l;
        }
        R const b = std::sqrt(norm2);
        R const inv_b = 1 / b;
        for (auto& v : v0) v *= inv_b;
        v0.swap(v1);
        alpha.push_back(a);
        beta.push_back(b);

        auto const ev = tridiagonal_eigenvalues(alpha, beta);
        auto const min = *std::min_element(ev.begin(), ev.end());
        auto const max = *std::max_element(ev.begin(), ev.end());
        auto const is_converged_min = std::abs((previous_min - min) / min) < precision;
        auto const is_converged_max = std::abs((previous_max - max) / max) < precision;
        if (is_converged_min && is_converged_max) { return {static_cast<double>(min), static_cast<double>(max), i}; }
        previous_min = min;
        previous_max = max;
    }
    throw std::runtime_error{"Lanczos algorithm did not converge for the min/max eigenvalues."};
}

// ------------------------------------------------------------------------------------------------
// Reconstruction (kpm/reconstruct.hpp:16-143); R = float for f32/c64 models in native mode
// ------------------------------------------------------------------------------------------------
template<class R, class A>
void spectral_density(std::vector<A> const& moments, int num_moments, int cols, double const* energy, int ne,
                      ScaleD s, double* out /* column-major ne x cols */) {
    auto const scale = Scale<R>(s);
    R const k = R{2 / pi_f} / scale.a;
    for (int c = 0; c < cols; ++c) {
        for (int i = 0; i < ne; ++i) {
            R const E = (static_cast<R>(energy[i]) - scale.b) / scale.a;
            R const acos_e = std::acos(E);
            R sum{0};
            for (int n = 0; n < num_moments; ++n) {
                sum += static_cast<R>(real_(moments[static_cast<size_t>(n) * cols + c])) * std::cos(static_cast<R>(n) * acos_e);
            }
            out[static_cast<size_t>(c) * ne + i] = static_cast<double>(k / std::sqrt(1 - E * E) * sum);
        }
    }
}

template<class R, class A>
void greens_function(std::vector<A> const& moments, double const* energy, int ne, ScaleD s, cd* out) {
    using C = std::complex<R>;
    auto const scale = Scale<R>(s);
    C const i1(0, 1);
    C const k = -R{2} * i1 / scale.a;
    for (int i = 0; i < ne; ++i) {
        R const E = (static_cast<R>(energy[i]) - scale.b) / scale.a;
        R const acos_e = std::acos(E);
        C sum{0};
        for (size_t n = 0; n < moments.size(); ++n) {
            auto const m = to_cd(moments[n]);
            sum += C(static_cast<R>(m.real()), static_cast<R>(m.imag())) * std::exp(-i1 * (static_cast<R>(n) * acos_e));
        }
        auto const g = k / std::sqrt(1 - E * E) * sum;
        out[i] = cd(g.real(), g.imag());
    }
}

/// moments: M x M row-major, already damped.  Result is complex (the facade takes .real()).
template<class R>
void kubo_bastin(std::vector<cd> const& moments, int M, double const* chemical_pot, int nmu,
                 std::vector<double> const& energy_samples, double temperature, ScaleD s, bool scalar_is_complex, cd* out) {
    using C = std::complex<R>;
    auto const scale = Scale<R>(s);
    auto const inv_kbt_sc = static_cast<R>(scale.a / (kb_f * temperature));
    int const np = static_cast<int>(energy_samples.size());
    std::vector<R> en(np);
    for (int i = 0; i < np; ++i) en[i] = (static_cast<R>(energy_samples[i]) - scale.b) / scale.a;

    std::vector<C> mu(static_cast<size_t>(M) * M);
    for (size_t i = 0; i < mu.size(); ++i) mu[i] = C(static_cast<R>(moments[i].real()), static_cast<R>(moments[i].imag()));

    std::vector<C> sum_nm(np);
    std::vector<C> a_n(M);
    std::vector<R> t_m(M);
    C const i1(0, 1);
    for (int p = 0; p < np; ++p) {
        R const e = en[p];
        R const ac = std::acos(e);
        R const sq = std::sqrt(R{1} - e * e);
        for (int n = 0; n < M; ++n) {
            R const rn = static_cast<R>(n);
            a_n[n] = (e - i1 * rn * sq) * std::exp(i1 * ac * rn);  // sqrt_n * exp_n, depends on the column
            t_m[n] = std::cos(ac * rn);                            // depends on the row
        }
        // g_p(m, n) = a_n[n] * t_m[m];  gamma = g_p + g_p^H;  sum over moments(m, n) * gamma(m, n)
        C total{0};
        for (int m = 0; m < M; ++m) {
            for (int n = 0; n < M; ++n) {
                C const gamma = a_n[n] * t_m[m] + std::conj(a_n[m] * t_m[n]);
                total += mu[static_cast<size_t>(m) * M + n] * gamma;
            }
        }
        R const k = R{1} / ((R{1} - e * e) * (R{1} - e * e));
        sum_nm[p] = k * total;
    }

    R const en_max = *std::max_element(en.begin(), en.end());
    R const en_min = *std::min_element(en.begin(), en.end());
    R const coeff = (en_max - en_min) / static_cast<R>(2 * np);
    C const prefix = C(R{4}) / (scale.a * scale.a);
    (void)scalar_is_complex;
    for (int j = 0; j < nmu; ++j) {
        R const mi = (static_cast<R>(chemical_pot[j]) - scale.b) / scale.a;
        C total{0}, first{0}, last{0};
        for (int p = 0; p < np; ++p) {
            R const fd = R{1} / (R{1} + std::exp((en[p] - mi) * inv_kbt_sc));
            C const f = fd * sum_nm[p];
            total += f;
            if (p == 0) { first = f; }
            if (p == np - 1) { last = f; }
        }
        C const r = prefix * (coeff * (R{2} * total - first - last));
        out[j] = cd(r.real(), r.imag());
    }
}

// ------------------------------------------------------------------------------------------------
// Thread pool (detail/thread.hpp:167-201) -- jobs are taken in submission order
// ------------------------------------------------------------------------------------------------
inline void run_jobs(std::vector<std::function<void()>>& jobs, int num_threads) {
    std::atomic<size_t> next{0};
    std::exception_ptr error;
    std::mutex error_mutex;
    auto worker = [&] {
        for (;;) {
            size_t const i = next.fetch_add(1);
            if (i >= jobs.size()) { return; }
            try { jobs[i](); } catch (...) { std::lock_guard<std::mutex> lk(error_mutex); error = std::current_exception(); }
        }
    };
    int const nt = std::max(1, std::min<int>(num_threads, static_cast<int>(jobs.size())));
    std::vector<std::thread> threads;
    for (int t = 1; t < nt; ++t) threads.emplace_back(worker);
    worker();
    for (auto& t : threads) t.join();
    if (error) { std::rethrow_exception(error); }
}

// ------------------------------------------------------------------------------------------------
// Core (src/kpm/Core.cpp:35-156) + DefaultCompute (src/kpm/default/Compute.cpp:14-131)
// ------------------------------------------------------------------------------------------------
struct Config {
    float min_energy = 0, max_energy = 0;
    Kernel kernel;
    bool use_ell = true;
    bool optimal_size = true, interleaved = true;
    float lanczos_precision = 0.002f;
    int num_threads = 1;
    bool hp = false;
};

struct CoreBase {
    virtual ~CoreBase() = default;
    Config config;
    double bounds_min = 0, bounds_max = 0;
    int lanczos_loops = 0;
    bool have_bounds = false;
    double moments_seconds = 0;
    int last_num_moments = 0;

    virtual int size() const = 0;
    virtual void compute_bounds() = 0;
    ScaleD scaling_factors() { compute_bounds(); return {bounds_min, bounds_max}; }

    virtual void optimize_for(Indices const& idx) = 0;
    virtual void oh_info(int* out /*nslices, src_offset, dest_offset, nnz*/) = 0;
    virtual std::vector<int> oh_slices() = 0;
    virtual Indices oh_idx() = 0;
    virtual std::vector<int> oh_reorder_map() = 0;
    virtual int oh_slice_index(int n, int num_moments) = 0;
    virtual void oh_matrix(int* indptr, int* indices, cd* data) = 0;

    virtual void random_vectors(int count, cd* out) = 0;
    virtual void dos_moments(int num_moments, int num_random, cd* out) = 0;
    virtual void ldos_moments(int num_moments, std::vector<int> const& idx, cd* out /*M x nidx row-major*/) = 0;
    virtual void greens_moments(int num_moments, int row, std::vector<int> const& cols, cd* out /*ncols x M*/) = 0;
    virtual void kubo_moments(int num_moments, float const* left, float const* right, int num_random, cd* out) = 0;
    virtual void moments(int num_moments, cd const* alpha, cd const* beta, int op_rows, int const* op_indptr,
                         int const* op_indices, cd const* op_data, cd* out) = 0;
    virtual void calc_dos(double const* e, int ne, double broadening, int num_random, double* out) = 0;
    virtual void calc_ldos(double const* e, int ne, double broadening, std::vector<int> const& idx, double* out) = 0;
    virtual void calc_greens(int row, std::vector<int> const& cols, double const* e, int ne, double broadening, cd* out) = 0;
    virtual void calc_conductivity(float const* left, float const* right, double const* mu, int nmu, double broadening,
                                   double temperature, int num_random, int num_points, cd* out) = 0;
    virtual double time_dos_steps(int num_moments, int num_random, int num_threads, bool cheap_starter) = 0;
    StepProbe* step_probe = nullptr;   // set for the duration of a timed run (bench.py cpu_baseline)
};

template<class T, bool HP> struct Core : CoreBase {
    using R = real_of<T>;
    using A = std::conditional_t<HP, hp_of<T>, T>;
    using RR = std::conditional_t<HP, double, R>;  // reconstruction precision

    Csr<T> h;
    OptimizedHamiltonian<T> oh;

    int size() const override { return h.rows; }

    void compute_bounds() override {
        if (have_bounds) { return; }
        if (config.min_energy == config.max_energy) {
            auto const lb = minmax_eigenvalues(h, config.lanczos_precision);
            bounds_min = lb.min; bounds_max = lb.max; lanczos_loops = lb.loops;
        } else {
            bounds_min = config.min_energy; bounds_max = config.max_energy;
        }
        have_bounds = true;
    }

    void optimize_for(Indices const& idx) override {
        oh.use_ell = config.use_ell;
        oh.is_reordered = config.optimal_size || config.interleaved;
        oh.optimize_for(h, idx, scaling_factors());
    }
    void oh_info(int* out) override {
        out[0] = static_cast<int>(oh.map.data.size()); out[1] = oh.map.src_offset; out[2] = oh.map.dest_offset;
        out[3] = oh.csr.nnz();
    }
    std::vector<int> oh_slices() override { return oh.map.data; }
    Indices oh_idx() override { return oh.idx; }
    std::vector<int> oh_reorder_map() override { return oh.reorder_map; }
    int oh_slice_index(int n, int num_moments) override { return oh.map.index(n, num_moments); }
    void oh_matrix(int* indptr, int* indices, cd* data) override {
        std::copy(oh.csr.indptr.begin(), oh.csr.indptr.end(), indptr);
        std::copy(oh.csr.indices.begin(), oh.csr.indices.end(), indices);
        for (size_t i = 0; i < oh.csr.data.size(); ++i) data[i] = to_cd(oh.csr.data[i]);
    }

    // ---- DefaultCompute::SelectAlgorithm::with (Compute.cpp:23-44) ----
    template<int B, class Collector>
    int with(Collector& collect, Starter<T>& starter, bool opt_size) {
        ftz_guard guard;
        int idx = 0;
        std::vector<T> r0;
        if (starter.draw && starter.finish) {
            starter.mutex->lock();
            idx = starter.count;
            auto u = draw_r0<T, B>(starter);
            starter.mutex->unlock();
            r0 = finish_r0<T, B>(starter, u);
        } else {
            starter.mutex->lock();
            idx = starter.count;
            r0 = make_r0<T, B>(starter);
            starter.mutex->unlock();
        }
        std::vector<T> r1 = oh.use_ell ? make_r1<T, B>(oh.ell, r0) : make_r1<T, B>(oh.csr, r0);
        collect.initial(r0, r1);
        run<B>(collect, std::move(r0), std::move(r1), opt_size);
        return idx;
    }
    template<int B, class Collector>
    void run(Collector& collect, std::vector<T> r0, std::vector<T> r1, bool opt_size) {
        if constexpr (Collector::diagonal) {
            if (config.interleaved) {
                if (oh.use_ell) diagonal_interleaved<T, A, B>(collect, std::move(r0), std::move(r1), oh.ell, oh.map, opt_size);
                else diagonal_interleaved<T, A, B>(collect, std::move(r0), std::move(r1), oh.csr, oh.map, opt_size);
            } else {
                if (oh.use_ell) diagonal_basic<T, A, B>(collect, std::move(r0), std::move(r1), oh.ell, oh.map, opt_size);
                else diagonal_basic<T, A, B>(collect, std::move(r0), std::move(r1), oh.csr, oh.map, opt_size);
            }
        } else {
            if (config.interleaved) {
                if (oh.use_ell) offdiagonal_interleaved<T>(collect, std::move(r0), std::move(r1), oh.ell, oh.map, opt_size);
                else offdiagonal_interleaved<T>(collect, std::move(r0), std::move(r1), oh.csr, oh.map, opt_size);
            } else {
                if (oh.use_ell) offdiagonal_basic<T>(collect, std::move(r0), std::move(r1), oh.ell, oh.map, opt_size);
                else offdiagonal_basic<T>(collect, std::move(r0), std::move(r1), oh.csr, oh.map, opt_size);
            }
        }
    }

    struct timer_scope {
        double& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        explicit timer_scope(double& a) : acc(a) {}
        ~timer_scope() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    };

    static constexpr int simd_batch = 32 / sizeof(T);  // support/simd.hpp:42-44 (AVX register width)

    /// SelectAlgorithm::operator()(BatchDiagonalMoments*) -- Compute.cpp:52-88.
    /// `accumulate`: BatchAccumulator (mean over vectors) else BatchConcatenator (M x nvec, row-major)
    std::vector<A> batch_diagonal(int num_moments, int num_vectors, Starter<T>& starter, bool opt_size, bool accumulate) {
        timer_scope ts(moments_seconds);
        constexpr int B = simd_batch;
        int const num_threads = std::max(1, config.num_threads);
        int num_batches = num_vectors / B;
        int num_singles = num_vectors % B;
        if (num_singles > num_threads * B / 2) { num_batches += 1; num_singles = 0; }

        std::vector<A> result(accumulate ? num_moments : static_cast<size_t>(num_moments) * num_vectors, A{0});
        std::mutex result_mutex;
        auto add = [&](std::vector<A> const& m, int cols_in, int idx) {  // Moments.cpp:7-81
            std::lock_guard<std::mutex> lk(result_mutex);
            int const remaining = num_vectors - idx;
            int const cols = remaining > cols_in ? cols_in : remaining;
            for (int n = 0; n < num_moments; ++n) {
                for (int j = 0; j < cols; ++j) {
                    if (accumulate) result[n] += m[static_cast<size_t>(n) * cols_in + j];
                    else result[static_cast<size_t>(n) * num_vectors + idx + j] = m[static_cast<size_t>(n) * cols_in + j];
                }
            }
        };
        std::vector<std::function<void()>> jobs;
        for (int i = 0; i < num_batches; ++i) {
            jobs.emplace_back([&] {
                DiagonalCollector<T, A, B> collect(num_moments);
                collect.probe = step_probe;
                auto const idx = with<B>(collect, starter, opt_size);
                add(collect.moments, B, idx);
            });
        }
        for (int i = 0; i < num_singles; ++i) {
            jobs.emplace_back([&] {
                DiagonalCollector<T, A, 1> collect(num_moments);
                collect.probe = step_probe;
                auto const idx = with<1>(collect, starter, opt_size);
                add(collect.moments, 1, idx);
            });
        }
        run_jobs(jobs, num_threads);
        if (accumulate && num_vectors != 1) {
            for (auto& v : result) v /= static_cast<real_of<A>>(num_vectors);
        }
        return result;
    }

    Starter<T> random_starter(Csr<T> op = {}) {
        Starter<T> s;
        s.vector_size = oh.size();
        auto fn = std::make_shared<RandomStarterFn<T>>(RandomStarterFn<T>{&oh, std::move(op)});
        s.make = [fn] { return (*fn)(); };
        s.draw = [fn] { return fn->draw(); };
        s.finish = [fn](std::vector<real_of<T>> const& u) { return fn->finish(u); };
        return s;
    }
    Starter<T> unit_starter() {
        Starter<T> s;
        s.vector_size = oh.size();
        auto i = std::make_shared<size_t>(0);
        auto sources = oh.idx.src;
        int const n = oh.size();
        s.make = [i, sources, n] {
            std::vector<T> r0(n, T{0});
            if (*i < sources.size()) { r0[sources[*i]] = T{1}; ++*i; }
            return r0;
        };
        return s;
    }
    Starter<T> constant_starter(std::vector<T> alpha) {
        Starter<T> s;
        s.vector_size = oh.size();
        auto const* poh = &oh;
        s.make = [alpha, poh] { auto r0 = alpha; poh->reorder(r0); return r0; };
        return s;
    }

    void random_vectors(int count, cd* out) override {
        OptimizedHamiltonian<T> plain;  // no reorder map
        plain.csr.rows = h.rows;
        RandomStarterFn<T> fn{&plain, {}};
        for (int j = 0; j < count; ++j) {
            auto const v = fn();
            for (int i = 0; i < h.rows; ++i) out[static_cast<size_t>(j) * h.rows + i] = to_cd(v[i]);
        }
    }

    // ---- raw (undamped) moments of each quantity ----
    std::vector<A> dos_moments_impl(int num_moments, int num_random) {
        optimize_for({{0}, {0}});
        auto starter = random_starter();
        return batch_diagonal(num_moments, num_random, starter, /*opt_size*/false, /*accumulate*/true);
    }
    void dos_moments(int num_moments, int num_random, cd* out) override {
        auto const m = dos_moments_impl(num_moments, num_random);
        for (int n = 0; n < num_moments; ++n) out[n] = to_cd(m[n]);
    }

    std::vector<A> ldos_moments_impl(int num_moments, std::vector<int> const& idx) {
        optimize_for({idx, idx});
        auto starter = unit_starter();
        return batch_diagonal(num_moments, static_cast<int>(idx.size()), starter, config.optimal_size, /*accumulate*/false);
    }
    void ldos_moments(int num_moments, std::vector<int> const& idx, cd* out) override {
        auto const m = ldos_moments_impl(num_moments, idx);
        for (size_t i = 0; i < m.size(); ++i) out[i] = to_cd(m[i]);
    }

    std::vector<std::vector<A>> greens_moments_impl(int num_moments, int row, std::vector<int> const& cols) {
        optimize_for({{row}, cols});
        auto starter = unit_starter();
        timer_scope ts(moments_seconds);
        if (oh.idx.is_diagonal()) {
            DiagonalCollector<T, A, 1> collect(num_moments);
            with<1>(collect, starter, config.optimal_size);
            return {collect.moments};
        }
        MultiUnitCollector<T, A> collect(num_moments, oh.idx);
        with<1>(static_cast<OffDiagonalCollector<T>&>(collect), starter, config.optimal_size);
        return collect.moments;
    }
    void greens_moments(int num_moments, int row, std::vector<int> const& cols, cd* out) override {
        auto const m = greens_moments_impl(num_moments, row, cols);
        for (size_t i = 0; i < m.size(); ++i) for (int n = 0; n < num_moments; ++n) out[i * num_moments + n] = to_cd(m[i][n]);
    }

    Csr<T> velocity(float const* alpha) const {  // Moments.cpp:132-156 (unscaled H, float positions)
        Csr<T> v = h;
        for (int row = 0; row < v.rows; ++row) {
            for (int n = v.indptr[row]; n < v.indptr[row + 1]; ++n) v.data[n] *= static_cast<T>(alpha[row] - alpha[v.indices[n]]);
        }
        return v;
    }

    /// Core::conductivity's moment part (Core.cpp:119-146): mu = (1/R) sum_j L_j * R_j^H
    std::vector<cd> kubo_moments_impl(int num_moments, float const* left, float const* right, int num_random) {
        optimize_for({{0}, {0}});
        auto starter_l = random_starter(velocity(left));
        auto starter_r = random_starter();
        int const M = num_moments;
        int const N = oh.size();
        std::vector<A> total(static_cast<size_t>(M) * M, A{0});
        for (int j = 0; j < num_random; ++j) {
            timer_scope ts(moments_seconds);
            DenseMatrixCollector<T> ml(M, oh, {});
            with<1>(static_cast<OffDiagonalCollector<T>&>(ml), starter_l, false);
            DenseMatrixCollector<T> mr(M, oh, velocity(right));
            with<1>(static_cast<OffDiagonalCollector<T>&>(mr), starter_r, false);
            std::vector<std::function<void()>> jobs;  // rows of the M x M product are independent
            for (int m = 0; m < M; ++m) {
                jobs.emplace_back([&, m] {
                    T const* a = ml.moments.data() + static_cast<size_t>(m) * N;
                    for (int n = 0; n < M; ++n) {
                        T const* b = mr.moments.data() + static_cast<size_t>(n) * N;
                        A s{0};
                        for (int i = 0; i < N; ++i) s += mul(cast_to<A>(a[i]), conj_(cast_to<A>(b[i])));
                        total[static_cast<size_t>(m) * M + n] += s;
                    }
                });
            }
            run_jobs(jobs, std::max(1, config.num_threads));
        }
        std::vector<cd> out(total.size());
        for (size_t i = 0; i < total.size(); ++i) {
            A const v = total[i] / static_cast<real_of<A>>(num_random);
            out[i] = to_cd(v);
        }
        return out;
    }
    void kubo_moments(int num_moments, float const* left, float const* right, int num_random, cd* out) override {
        auto const m = kubo_moments_impl(num_moments, left, right, num_random);
        std::copy(m.begin(), m.end(), out);
    }

    // ---- Core::moments (Core.cpp:35-56): damped, truncated to the requested length ----
    void moments(int num_moments, cd const* alpha_, cd const* beta_, int op_rows, int const* op_indptr,
                 int const* op_indices, cd const* op_data, cd* out) override {
        optimize_for({{0}, {0}});
        int const N = h.rows;
        int const M = round_num_moments(num_moments);
        std::vector<T> alpha(N);
        for (int i = 0; i < N; ++i) alpha[i] = cast_to<T>(alpha_[i]);
        auto starter = constant_starter(alpha);
        std::vector<A> m;
        {
            timer_scope ts(moments_seconds);
            if (!beta_ && op_rows == 0) {
                DiagonalCollector<T, A, 1> collect(M);
                with<1>(collect, starter, false);
                m = collect.moments;
            } else {
                std::vector<T> beta(N);
                for (int i = 0; i < N; ++i) beta[i] = cast_to<T>(beta_ ? beta_[i] : alpha_[i]);
                Csr<T> op;
                if (op_rows != 0) {
                    op.rows = op_rows;
                    op.indptr.assign(op_indptr, op_indptr + op_rows + 1);
                    op.indices.assign(op_indices, op_indices + op_indptr[op_rows]);
                    op.data.resize(op.indices.size());
                    for (size_t i = 0; i < op.data.size(); ++i) op.data[i] = cast_to<T>(op_data[i]);
                }
                GenericCollector<T, A> collect(M, oh, std::move(beta), std::move(op));
                with<1>(static_cast<OffDiagonalCollector<T>&>(collect), starter, false);
                m = collect.moments;
            }
        }
        auto const g = config.kernel.damping(M);
        for (int n = 0; n < num_moments; ++n) out[n] = to_cd(m[n] * static_cast<real_of<A>>(g[n]));
    }

    template<class V> void apply_damping(V& m, int M, int cols) const {
        auto const g = config.kernel.damping(M);
        for (int n = 0; n < M; ++n) for (int c = 0; c < cols; ++c) m[static_cast<size_t>(n) * cols + c] *= static_cast<real_of<A>>(g[n]);
    }

    int num_moments_for(double broadening) {
        auto const scale = scaling_factors();
        last_num_moments = config.kernel.required_num_moments(broadening / scale.a);
        return last_num_moments;
    }

    void calc_dos(double const* e, int ne, double broadening, int num_random, double* out) override {
        int const M = num_moments_for(broadening);
        auto m = dos_moments_impl(M, num_random);
        apply_damping(m, M, 1);
        spectral_density<RR, A>(m, M, 1, e, ne, scaling_factors(), out);
    }
    void calc_ldos(double const* e, int ne, double broadening, std::vector<int> const& idx, double* out) override {
        int const M = num_moments_for(broadening);
        auto m = ldos_moments_impl(M, idx);
        apply_damping(m, M, static_cast<int>(idx.size()));
        spectral_density<RR, A>(m, M, static_cast<int>(idx.size()), e, ne, scaling_factors(), out);
    }
    void calc_greens(int row, std::vector<int> const& cols, double const* e, int ne, double broadening, cd* out) override {
        int const M = num_moments_for(broadening);
        auto mv = greens_moments_impl(M, row, cols);
        for (size_t i = 0; i < mv.size(); ++i) {
            apply_damping(mv[i], M, 1);
            greens_function<RR, A>(mv[i], e, ne, scaling_factors(), out + i * ne);
        }
    }
    void calc_conductivity(float const* left, float const* right, double const* mu, int nmu, double broadening,
                           double temperature, int num_random, int num_points, cd* out) override {
        int const M = num_moments_for(broadening);
        auto m = kubo_moments_impl(M, left, right, num_random);
        auto const g = config.kernel.damping(M);
        for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) {
            auto const gg = static_cast<real_of<A>>(g[i]) * static_cast<real_of<A>>(g[j]);  // Kernel.hpp:49-56
            m[static_cast<size_t>(i) * M + j] *= static_cast<double>(gg);
        }
        std::vector<double> samples(num_points);  // Bounds.hpp:54 ArrayXd::LinSpaced(size, min, max)
        compute_bounds();
        for (int i = 0; i < num_points; ++i) {
            samples[i] = (num_points == 1) ? bounds_max
                       : (i == num_points - 1) ? bounds_max
                       : bounds_min + i * ((bounds_max - bounds_min) / (num_points - 1));
        }
        kubo_bastin<RR>(m, M, mu, nmu, samples, temperature, scaling_factors(), traits<T>::cplx, out);
    }

    /// CPU baseline: time the reference-shaped DOS moment computation (threads over SIMD batches)
    double time_dos_steps(int num_moments, int num_random, int num_threads, bool cheap_starter) override {
        optimize_for({{0}, {0}});
        auto const saved = config.num_threads;
        config.num_threads = num_threads;
        moments_seconds = 0;
        auto starter = random_starter();
        if (cheap_starter) {
            // Timing aid for bounded CPU samples: unit-modulus +-1 entries from a xorshift generator instead of
            // the reference's mutex-serialised MT19937 (+ complex exp) starter, whose fixed cost per vector would
            // dominate a run with few moments.  The recursion cost does not depend on the values.
            auto state = std::make_shared<uint64_t>(0x9E3779B97F4A7C15ull);
            int const size = oh.size();
            starter.draw = nullptr; starter.finish = nullptr;
            starter.make = [state, size]() {
                std::vector<T> r0(size);
                uint64_t x = *state;
                for (auto& v : r0) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (x & 1u) ? T(1) : T(-1); }
                *state = x;
                return r0;
            };
        }
        (void)batch_diagonal(num_moments, num_random, starter, false, true);
        config.num_threads = saved;
        return moments_seconds;
    }
};

template<class T> CoreBase* make_core(bool hp) { return hp ? static_cast<CoreBase*>(new Core<T, true>()) : new Core<T, false>(); }

template<class T, bool HP> void load_h(Core<T, HP>* c, int n, int const* indptr, int const* indices, void const* data) {
    c->h.rows = n;
    c->h.indptr.assign(indptr, indptr + n + 1);
    c->h.indices.assign(indices, indices + indptr[n]);
    auto const* d = static_cast<T const*>(data);
    c->h.data.assign(d, d + indptr[n]);
}

thread_local std::string last_error;

} // namespace orc

// ------------------------------------------------------------------------------------------------
// C interface for ctypes (oracle/oracle.py)
// ------------------------------------------------------------------------------------------------
using namespace orc;

#define ORC_TRY try {
#define ORC_CATCH } catch (std::exception const& e) { last_error = e.what(); return 1; } return 0;

extern "C" {

const char* orc_error() { return last_error.c_str(); }

/// dtype: 0 f32, 1 c64, 2 f64, 3 c128
void* orc_create(int dtype, int n, int const* indptr, int const* indices, void const* data,
                 float emin, float emax, int kernel_kind, double lambda, int use_ell, int optimal_size,
                 int interleaved, float lanczos_precision, int num_threads, int hp) {
    try {
        if (emin > emax) { throw std::invalid_argument("KPM: Invalid energy range specified (min > max)."); }
        CoreBase* core = nullptr;
        bool const is_hp = hp != 0;
        switch (dtype) {
            case 0: core = make_core<float>(is_hp); break;
            case 1: core = make_core<cf>(is_hp); break;
            case 2: core = make_core<double>(is_hp); break;
            case 3: core = make_core<cd>(is_hp); break;
            default: throw std::invalid_argument("bad dtype");
        }
#define ORC_LOAD(T) if (is_hp) load_h(static_cast<Core<T, true>*>(core), n, indptr, indices, data); \
                    else load_h(static_cast<Core<T, false>*>(core), n, indptr, indices, data);
        switch (dtype) { case 0: ORC_LOAD(float) break; case 1: ORC_LOAD(cf) break; case 2: ORC_LOAD(double) break; default: ORC_LOAD(cd) break; }
        core->config.min_energy = emin; core->config.max_energy = emax;
        core->config.kernel.kind = kernel_kind; core->config.kernel.lambda = lambda;
        if (kernel_kind == 1 && lambda <= 0) { delete core; throw std::invalid_argument("Lorentz kernel: lambda must be positive."); }
        core->config.use_ell = use_ell != 0;
        core->config.optimal_size = optimal_size != 0; core->config.interleaved = interleaved != 0;
        core->config.lanczos_precision = lanczos_precision;
        core->config.num_threads = num_threads > 0 ? num_threads : static_cast<int>(std::thread::hardware_concurrency());
        core->config.hp = is_hp;
        return core;
    } catch (std::exception const& e) { last_error = e.what(); return nullptr; }
}
void orc_destroy(void* p) { delete static_cast<CoreBase*>(p); }

int orc_bounds(void* p, double* out /*min, max, a, b, loops*/) { ORC_TRY
    auto* c = static_cast<CoreBase*>(p);
    auto const s = c->scaling_factors();
    out[0] = c->bounds_min; out[1] = c->bounds_max; out[2] = s.a; out[3] = s.b; out[4] = c->lanczos_loops;
ORC_CATCH }

int orc_required_num_moments(void* p, double broadening) {
    try { auto* c = static_cast<CoreBase*>(p); auto const s = c->scaling_factors(); return c->config.kernel.required_num_moments(broadening / s.a); }
    catch (std::exception const& e) { last_error = e.what(); return -1; }
}
int orc_kernel_required_num_moments(int kind, double lambda, double scaled_broadening) { return Kernel{kind, lambda}.required_num_moments(scaled_broadening); }
void orc_kernel_damping(int kind, double lambda, int n, double* out) { auto const g = Kernel{kind, lambda}.damping(n); std::copy(g.begin(), g.end(), out); }
void orc_scale(double emin, double emax, double* out) { ScaleD s(emin, emax); out[0] = s.a; out[1] = s.b; }

int orc_optimize_for(void* p, int const* src, int nsrc, int const* dest, int ndest, int* info) { ORC_TRY
    auto* c = static_cast<CoreBase*>(p);
    c->optimize_for({std::vector<int>(src, src + nsrc), std::vector<int>(dest, dest + ndest)});
    c->oh_info(info);
ORC_CATCH }
int orc_oh_get(void* p, int* slices, int* src, int* dest, int* reorder_map) { ORC_TRY
    auto* c = static_cast<CoreBase*>(p);
    auto const s = c->oh_slices(); std::copy(s.begin(), s.end(), slices);
    auto const idx = c->oh_idx(); std::copy(idx.src.begin(), idx.src.end(), src); std::copy(idx.dest.begin(), idx.dest.end(), dest);
    auto const m = c->oh_reorder_map(); std::copy(m.begin(), m.end(), reorder_map);
ORC_CATCH }
int orc_oh_matrix(void* p, int* indptr, int* indices, cd* data) { ORC_TRY static_cast<CoreBase*>(p)->oh_matrix(indptr, indices, data); ORC_CATCH }
int orc_oh_slice_index(void* p, int n, int num_moments) { return static_cast<CoreBase*>(p)->oh_slice_index(n, num_moments); }

int orc_random_vectors(void* p, int count, cd* out) { ORC_TRY static_cast<CoreBase*>(p)->random_vectors(count, out); ORC_CATCH }
int orc_dos_moments(void* p, int M, int num_random, cd* out) { ORC_TRY static_cast<CoreBase*>(p)->dos_moments(M, num_random, out); ORC_CATCH }
int orc_ldos_moments(void* p, int M, int const* idx, int nidx, cd* out) { ORC_TRY
    static_cast<CoreBase*>(p)->ldos_moments(M, std::vector<int>(idx, idx + nidx), out); ORC_CATCH }
int orc_greens_moments(void* p, int M, int row, int const* cols, int ncols, cd* out) { ORC_TRY
    static_cast<CoreBase*>(p)->greens_moments(M, row, std::vector<int>(cols, cols + ncols), out); ORC_CATCH }
int orc_kubo_moments(void* p, int M, float const* left, float const* right, int num_random, cd* out) { ORC_TRY
    static_cast<CoreBase*>(p)->kubo_moments(M, left, right, num_random, out); ORC_CATCH }
int orc_moments(void* p, int num_moments, cd const* alpha, cd const* beta, int op_rows, int const* op_indptr,
                int const* op_indices, cd const* op_data, cd* out) { ORC_TRY
    static_cast<CoreBase*>(p)->moments(num_moments, alpha, beta, op_rows, op_indptr, op_indices, op_data, out); ORC_CATCH }

int orc_calc_dos(void* p, double const* e, int ne, double broadening, int num_random, double* out) { ORC_TRY
    static_cast<CoreBase*>(p)->calc_dos(e, ne, broadening, num_random, out); ORC_CATCH }
int orc_calc_ldos(void* p, double const* e, int ne, double broadening, int const* idx, int nidx, double* out) { ORC_TRY
    static_cast<CoreBase*>(p)->calc_ldos(e, ne, broadening, std::vector<int>(idx, idx + nidx), out); ORC_CATCH }
int orc_calc_greens(void* p, int row, int const* cols, int ncols, double const* e, int ne, double broadening, cd* out) { ORC_TRY
    static_cast<CoreBase*>(p)->calc_greens(row, std::vector<int>(cols, cols + ncols), e, ne, broadening, out); ORC_CATCH }
int orc_calc_conductivity(void* p, float const* left, float const* right, double const* mu, int nmu, double broadening,
                          double temperature, int num_random, int num_points, cd* out) { ORC_TRY
    static_cast<CoreBase*>(p)->calc_conductivity(left, right, mu, nmu, broadening, temperature, num_random, num_points, out); ORC_CATCH }

int orc_last_num_moments(void* p) { return static_cast<CoreBase*>(p)->last_num_moments; }
double orc_moments_seconds(void* p) { return static_cast<CoreBase*>(p)->moments_seconds; }
int orc_time_dos(void* p, int M, int num_random, int num_threads, int cheap_starter, double* seconds) { ORC_TRY
    *seconds = static_cast<CoreBase*>(p)->time_dos_steps(M, num_random, num_threads, cheap_starter != 0); ORC_CATCH }
/// Same run with a probe around recursion steps [n1, n2) (collector calls n1 .. n2 of calc_moments' diagonal loop, each
/// step = 2 moments): out[0] = seconds of the whole moments phase (what the reference's moments_timer covers),
/// out[1] = seconds the slowest job spent between the two steps, out[2] = jobs that reported.
int orc_time_dos_probe(void* p, int M, int num_random, int num_threads, int cheap_starter, int n1, int n2, double* out) { ORC_TRY
    auto* core = static_cast<CoreBase*>(p);
    StepProbe probe;
    probe.n1 = n1; probe.n2 = n2;
    core->step_probe = &probe;
    try { out[0] = core->time_dos_steps(M, num_random, num_threads, cheap_starter != 0); }
    catch (...) { core->step_probe = nullptr; throw; }
    core->step_probe = nullptr;
    out[1] = probe.max_elapsed; out[2] = probe.reports; ORC_CATCH }
int orc_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

} // extern "C"
