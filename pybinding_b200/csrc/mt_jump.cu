// mt_jump.cu -- segment-parallel generation of the reference's MT19937 stream (jump-ahead).
//
// The reference draws all stochastic starters of one calculation from a single default-seeded
// std::mt19937 (cppcore/src/kpm/Starter.cpp:48-83): vector j consumes draws [j*N*w, (j+1)*N*w).  A
// single generator is inherently sequential (one 624-word block after the other), which would leave
// one SM producing 2.4e9 words for the 38 M-site / 64-vector benchmark while 147 SMs wait, and would
// make a rank that owns vectors [j0, j1) first skip j0*N*w draws.  MT19937 is a linear recurrence
// over GF(2), so its state after J steps is a fixed polynomial in the one-step map applied to the
// current state (Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008):
//
//      U[t + J] = XOR_{i : g_i = 1} U[t + i],      g(x) = x^J mod phi(x),
//
// where U is the raw (untempered) word stream and phi the characteristic polynomial (degree 19937).
// Host side: phi by Berlekamp-Massey on one output bit, g by square-and-multiply (cached per J).
// Device side: one CTA per jump expands 19937 + 624 raw words from the source state into shared
// memory and forms the 624 words of the target state as XOR-convolutions with g.  The word range
// of a batch is cut into S segments; the S start states are produced by log2(S) rounds of jumps
// (state k -> state k + 2^r, polynomial x^(L 2^r)), after which S CTAs generate their segments
// concurrently.  The output is bit-identical to the sequential generator (tests compare both).
//
// State convention: a "window" W_t = (U[t], ..., U[t+623]).  The low 31 bits of a window's first word
// are not part of the generator's state and are not reproduced by a jump, so the state for output
// position P is the window W_{P-1} with the read position set to 1.
#include "kernels.cuh"

#include <array>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace pbk {

namespace {

constexpr int DEG = 19937;
constexpr int PW = 313;                       // 64-bit words holding bits 0..19937 (+ slack)
constexpr int MT_M = 397;
constexpr int JUMP_THREADS = 640;
constexpr int RAW_CHUNKS = 33;                // 33 * 624 = 20592 >= 19937 + 623 raw words per jump
constexpr int RAW_WORDS = RAW_CHUNKS * MT_N;
constexpr int GWORDS = MT_N;                  // jump polynomial as 624 32-bit words (bits >= 19937 are zero)

using Bits = std::vector<uint64_t>;

inline uint32_t twist_word_h(uint32_t cur, uint32_t nxt, uint32_t far_) {
    uint32_t const yy = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far_ ^ (yy >> 1) ^ ((yy & 1u) ? 0x9908b0dfu : 0u);
}

/// raw stream U[t0 .. t0 + count) continuing the window `w` (= U[t0 - 624 .. t0))
void host_raw_stream(const uint32_t* w, uint32_t* out, int64_t count) {
    std::vector<uint32_t> buf(w, w + MT_N);
    buf.resize(MT_N + static_cast<size_t>(count));
    for (int64_t t = 0; t < count; ++t) buf[MT_N + t] = twist_word_h(buf[t], buf[t + 1], buf[t + MT_M]);
    std::memcpy(out, buf.data() + MT_N, sizeof(uint32_t) * static_cast<size_t>(count));
}

void seed_window(uint32_t* w) {  // std::mt19937 default seed: the array before the first twist = W_{-624}
    uint32_t x = 5489u;
    w[0] = x;
    for (uint32_t i = 1; i < MT_N; ++i) { x = 1812433253u * (x ^ (x >> 30)) + i; w[i] = x; }
}

inline bool get_bit(Bits const& b, int64_t i) { return (b[static_cast<size_t>(i >> 6)] >> (i & 63)) & 1u; }
inline void flip_bit(Bits& b, int64_t i) { b[static_cast<size_t>(i >> 6)] ^= uint64_t{1} << (i & 63); }

/// dst ^= src << shift   (src has `nw` words; dst must hold nw + shift/64 + 1 words)
void xor_shifted(Bits& dst, Bits const& src, size_t nw, int64_t shift) {
    size_t const ws = static_cast<size_t>(shift >> 6);
    int const bs = static_cast<int>(shift & 63);
    if (bs == 0) {
        for (size_t i = 0; i < nw; ++i) dst[i + ws] ^= src[i];
    } else {
        uint64_t carry = 0;
        for (size_t i = 0; i < nw; ++i) {
            dst[i + ws] ^= (src[i] << bs) | carry;
            carry = src[i] >> (64 - bs);
        }
        dst[nw + ws] ^= carry;
    }
}

/// characteristic polynomial of MT19937 (bit i = coefficient of x^i, degree 19937) by Berlekamp-Massey
/// on the sequence of one raw output bit
Bits compute_phi() {
    int const N = 2 * DEG + 64;
    std::vector<uint32_t> seedw(MT_N), raw(N);
    seed_window(seedw.data());
    host_raw_stream(seedw.data(), raw.data(), N);

    size_t const W = static_cast<size_t>(N / 64 + 2);
    Bits C(W, 0), B(W, 0), T(W, 0), S(W, 0);  // S: reversed history, bit i = s[n - i]
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < N; ++n) {
        // shift the history left by one and insert s[n] at bit 0
        uint64_t carry = (raw[n] >> 7) & 1u;  // any fixed bit of the raw word works; bit 7 here
        size_t const used = static_cast<size_t>(n / 64 + 1);
        for (size_t i = 0; i < used && i < W; ++i) { uint64_t const nc = S[i] >> 63; S[i] = (S[i] << 1) | carry; carry = nc; }
        // discrepancy = parity(C & S) over bits 0..L
        uint64_t acc = 0;
        size_t const lw = static_cast<size_t>(L / 64 + 1);
        for (size_t i = 0; i < lw; ++i) acc ^= C[i] & S[i];
        bool const d = __builtin_parityll(acc);
        if (!d) { ++m; continue; }
        size_t const bw = static_cast<size_t>((n + 1) / 64 + 1);
        if (2 * L <= n) {
            T = C;
            xor_shifted(C, B, std::min(bw, W - static_cast<size_t>(m / 64) - 1), m);
            L = n + 1 - L;
            B = T;
            m = 1;
        } else {
            xor_shifted(C, B, std::min(bw, W - static_cast<size_t>(m / 64) - 1), m);
            ++m;
        }
    }
    // connection polynomial C (s[n] = sum_{i>=1} C_i s[n-i]) -> characteristic polynomial phi(x) = x^L C(1/x)
    Bits phi(PW, 0);
    if (L != DEG) return Bits();  // cannot happen for MT19937; signalled to the caller
    for (int i = 0; i <= L; ++i) if (get_bit(C, i)) flip_bit(phi, L - i);
    return phi;
}

struct PolyContext {
    Bits phi;                          // degree 19937
    std::vector<Bits> phi_sh;          // phi << s for s = 0..63 (PW + 1 words)
    PolyContext() {
        phi = compute_phi();
        if (phi.empty()) return;
        phi_sh.assign(64, Bits(PW + 1, 0));
        for (int s = 0; s < 64; ++s) xor_shifted(phi_sh[s], phi, PW, s);
    }
    /// a (2 * PW words, degree < 2 * DEG) mod phi, in place; result in bits 0..DEG-1
    void reduce(Bits& a) const {
        for (int64_t i = 2 * DEG - 2; i >= DEG; --i) {
            if (!get_bit(a, i)) continue;
            int64_t const sh = i - DEG;
            size_t const ws = static_cast<size_t>(sh >> 6);
            Bits const& p = phi_sh[static_cast<size_t>(sh & 63)];
            for (size_t k = 0; k < static_cast<size_t>(PW + 1); ++k) a[k + ws] ^= p[k];
        }
    }
    static uint64_t spread(uint32_t v) {  // bit i -> bit 2i
        uint64_t x = v;
        x = (x | (x << 16)) & 0x0000ffff0000ffffull;
        x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
        x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
        x = (x | (x << 2)) & 0x3333333333333333ull;
        x = (x | (x << 1)) & 0x5555555555555555ull;
        return x;
    }
    Bits square(Bits const& a) const {  // a^2 mod phi (squaring over GF(2) only spreads the bits)
        Bits r(2 * PW + 2, 0);
        for (size_t i = 0; i < static_cast<size_t>(PW); ++i) {
            r[2 * i] = spread(static_cast<uint32_t>(a[i]));
            r[2 * i + 1] = spread(static_cast<uint32_t>(a[i] >> 32));
        }
        reduce(r);
        r.resize(PW);  // reduce() cleared every bit >= 19937
        return r;
    }
    Bits times_x(Bits const& a) const {
        Bits r(PW, 0);
        uint64_t carry = 0;
        for (size_t i = 0; i < static_cast<size_t>(PW); ++i) { r[i] = (a[i] << 1) | carry; carry = a[i] >> 63; }
        if (get_bit(r, DEG)) for (size_t i = 0; i < static_cast<size_t>(PW); ++i) r[i] ^= phi[i];
        return r;
    }
    Bits power_of_x(uint64_t J) const {  // x^J mod phi
        Bits r(PW, 0);
        r[0] = 1;
        for (int b = 63; b >= 0; --b) {
            bool const any_above = (b < 63) && (J >> (b + 1)) != 0;
            if (any_above) r = square(r);
            if ((J >> b) & 1u) r = times_x(r);
        }
        return r;
    }
};

PolyContext const& poly_context() {
    static PolyContext ctx;
    return ctx;
}

struct PolyCache {
    std::mutex mutex;
    std::map<uint64_t, std::shared_ptr<std::vector<uint32_t>>> polys;  // J -> 624 32-bit words
};
PolyCache& poly_cache() { static PolyCache c; return c; }

std::vector<uint32_t> to_words32(Bits const& g) {
    std::vector<uint32_t> w(GWORDS, 0);
    for (int i = 0; i < GWORDS; ++i) {
        uint64_t const v = g[static_cast<size_t>(i >> 1)];
        w[i] = static_cast<uint32_t>((i & 1) ? (v >> 32) : v);
    }
    return w;
}

/// jump polynomial x^J mod phi as 624 32-bit words (cached; x^(J/2) in the cache gives x^J by one squaring)
std::shared_ptr<std::vector<uint32_t>> jump_poly(uint64_t J) {
    auto& cache = poly_cache();
    {
        std::lock_guard<std::mutex> lock(cache.mutex);
        auto const it = cache.polys.find(J);
        if (it != cache.polys.end()) return it->second;
    }
    auto const& ctx = poly_context();
    if (ctx.phi.empty()) return nullptr;
    std::shared_ptr<std::vector<uint32_t>> half;
    if (J % 2 == 0 && J > 0) {
        std::lock_guard<std::mutex> lock(cache.mutex);
        auto const it = cache.polys.find(J / 2);
        if (it != cache.polys.end()) half = it->second;
    }
    Bits g;
    if (half) {  // x^J = (x^(J/2))^2
        Bits h(PW, 0);
        for (int i = 0; i < GWORDS; ++i) h[static_cast<size_t>(i >> 1)] |= static_cast<uint64_t>((*half)[i]) << ((i & 1) ? 32 : 0);
        g = ctx.square(h);
    } else {
        g = ctx.power_of_x(J);
    }
    auto result = std::make_shared<std::vector<uint32_t>>(to_words32(g));
    std::lock_guard<std::mutex> lock(cache.mutex);
    cache.polys[J] = result;
    return result;
}

// ------------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t temper_d(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__device__ __forceinline__ uint32_t twist_word_d(uint32_t cur, uint32_t nxt, uint32_t far_) {
    uint32_t const yy = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far_ ^ (yy >> 1) ^ ((yy & 1u) ? 0x9908b0dfu : 0u);
}

constexpr int STATE_STRIDE = MT_N + 8;  // words per stored state: 624 window words + read position

/// dst state [first_dst + b] = jump(src state [b]) with polynomial g, for b < count (one CTA per jump)
__global__ void __launch_bounds__(JUMP_THREADS) mt_jump_kernel(uint32_t* states, const uint32_t* __restrict__ g, int first_dst, int count) {
    extern __shared__ uint32_t sh[];
    uint32_t* U = sh;                 // RAW_WORDS raw words, U[0..623] = the source window
    uint32_t* gs = sh + RAW_WORDS;    // GWORDS
    int const b = blockIdx.x;
    if (b >= count) return;
    int const t = threadIdx.x;
    const uint32_t* src = states + static_cast<int64_t>(b) * STATE_STRIDE;
    for (int i = t; i < MT_N; i += JUMP_THREADS) { U[i] = src[i]; gs[i] = g[i]; }
    __syncthreads();
    // expand the raw stream chunk by chunk (three dependent stages per 624-word chunk)
    for (int c = 1; c < RAW_CHUNKS; ++c) {
        uint32_t* o = U + (c - 1) * MT_N;
        uint32_t* nw = U + c * MT_N;
        if (t < 227) nw[t] = twist_word_d(o[t], o[t + 1], o[t + MT_M]);
        __syncthreads();
        if (t >= 227 && t < 454) nw[t] = twist_word_d(o[t], o[t + 1], nw[t - 227]);
        __syncthreads();
        if (t >= 454 && t < MT_N) nw[t] = twist_word_d(o[t], o[t + 1], nw[t - 227]);  // o[624] == nw[0]
        __syncthreads();
    }
    if (t < MT_N) {
        uint32_t acc0 = 0, acc1 = 0;
        for (int wi = 0; wi < GWORDS; ++wi) {
            uint32_t gw = gs[wi];
            const uint32_t* base = U + wi * 32 + t;
            while (gw) {
                int const bit = __ffs(static_cast<int>(gw)) - 1;
                gw &= gw - 1;
                acc0 ^= base[bit];
                if (gw) {
                    int const bit2 = __ffs(static_cast<int>(gw)) - 1;
                    gw &= gw - 1;
                    acc1 ^= base[bit2];
                }
            }
        }
        uint32_t* dst = states + static_cast<int64_t>(first_dst + b) * STATE_STRIDE;
        dst[t] = acc0 ^ acc1;
        if (t == 0) dst[MT_N] = 1u;  // read position: word 0 of a jumped window only carries the top state bit
    }
}

__global__ void mt_seed_state_kernel(uint32_t* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t x = 5489u;
        state[0] = x;
        for (uint32_t i = 1; i < MT_N; ++i) { x = 1812433253u * (x ^ (x >> 30)) + i; state[i] = x; }
        state[MT_N] = MT_N;
    }
}

/// CTA b writes the tempered outputs of segment b: out[b * L .. min((b + 1) * L, total))
__global__ void __launch_bounds__(256) mt_segments_kernel(const uint32_t* __restrict__ states, uint32_t* __restrict__ out, int64_t L, int64_t total) {
    __shared__ uint32_t mt[MT_N + 1];
    int const t = threadIdx.x;
    const uint32_t* st = states + static_cast<int64_t>(blockIdx.x) * STATE_STRIDE;
    for (int i = t; i < MT_N; i += 256) mt[i] = st[i];
    int pos = static_cast<int>(st[MT_N]);
    __syncthreads();
    int64_t const begin = static_cast<int64_t>(blockIdx.x) * L;
    int64_t const count = (begin + L <= total) ? L : (total - begin);
    uint32_t* dst = out + begin;
    int64_t produced = 0;
    while (produced < count) {
        if (pos == MT_N) {  // in-place twist, three dependent stages
            uint32_t w = 0;
            if (t < 227) w = twist_word_d(mt[t], mt[t + 1], mt[t + MT_M]);
            __syncthreads();
            if (t < 227) mt[t] = w;
            __syncthreads();
            int k = t + 227;
            if (t < 227) w = twist_word_d(mt[k], mt[k + 1], mt[k - 227]);
            __syncthreads();
            if (t < 227) mt[k] = w;
            __syncthreads();
            k = t + 454;
            if (k < MT_N) w = twist_word_d(mt[k], mt[(k + 1) % MT_N], mt[k - 227]);
            __syncthreads();
            if (k < MT_N) mt[k] = w;
            __syncthreads();
            pos = 0;
        }
        int64_t const left = count - produced;
        int const avail = static_cast<int>(left < (MT_N - pos) ? left : (MT_N - pos));
        for (int i = t; i < avail; i += 256) dst[produced + i] = temper_d(mt[pos + i]);
        pos += avail;
        produced += avail;
    }
}

} // anonymous namespace

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------
int64_t mt_stream_scratch_words(int max_segments) { return static_cast<int64_t>(max_segments) * STATE_STRIDE + 2 * GWORDS; }

/// Host-only reference of the jump (used by the CPU tests): window W_{position-1} of the default-seeded stream
bool mt_jump_window_host(uint64_t position, uint32_t* window) {
    auto const g = jump_poly(position + 623);
    if (!g) return false;
    std::vector<uint32_t> seedw(MT_N), raw(MT_N + RAW_WORDS);
    seed_window(seedw.data());
    std::memcpy(raw.data(), seedw.data(), sizeof(uint32_t) * MT_N);
    host_raw_stream(seedw.data(), raw.data() + MT_N, RAW_WORDS);
    for (int k = 0; k < MT_N; ++k) {
        uint32_t acc = 0;
        for (int i = 0; i < DEG; ++i) if (((*g)[static_cast<size_t>(i >> 5)] >> (i & 31)) & 1u) acc ^= raw[static_cast<size_t>(i + k)];
        window[k] = acc;
    }
    return true;
}

/// Write the tempered outputs [position, position + count) of the default-seeded std::mt19937 stream to `out`.
/// `states_dev`: scratch of mt_stream_scratch_words(max_segments) words.  Segment-parallel via jump-ahead.
cudaError_t launch_mt_stream(uint32_t* states_dev, int max_segments, uint64_t position, int64_t count, uint32_t* out, cudaStream_t s,
                             int* launches) {
    if (count <= 0) return cudaSuccess;
    constexpr int64_t MIN_SEGMENT = 64 * MT_N;
    int S = static_cast<int>(std::min<int64_t>(max_segments, (count + MIN_SEGMENT - 1) / MIN_SEGMENT));
    if (S < 1) S = 1;
    int64_t const L = (count + S - 1) / S;
    S = static_cast<int>((count + L - 1) / L);
    uint32_t* gpoly_dev = states_dev + static_cast<int64_t>(max_segments) * STATE_STRIDE;
    size_t const smem = sizeof(uint32_t) * (RAW_WORDS + GWORDS);
    static std::once_flag attr_once;
    cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); });
    if (attr_err != cudaSuccess) return attr_err;
    int nl = 0;

    mt_seed_state_kernel<<<1, 32, 0, s>>>(states_dev);
    ++nl;
    auto jump = [&](uint64_t J, int first_dst, int cnt) -> cudaError_t {
        auto const g = jump_poly(J);
        if (!g) return cudaErrorUnknown;
        // the polynomial buffer is reused by the next jump: stream order keeps the copies and kernels apart
        cudaError_t err = cudaMemcpyAsync(gpoly_dev, g->data(), sizeof(uint32_t) * GWORDS, cudaMemcpyHostToDevice, s);
        if (err != cudaSuccess) return err;
        mt_jump_kernel<<<cnt, JUMP_THREADS, smem, s>>>(states_dev, gpoly_dev, first_dst, cnt);
        ++nl;
        // pageable-memory copies return once staged, but keep the host vector alive and ordered anyway
        return cudaStreamSynchronize(s);
    };
    cudaError_t err;
    if (!(position == 0 && S == 1)) {  // (a single segment at the start of the stream runs straight from the seed)
        // state 0: window W_{position-1} = jump of the seed window W_{-624} by position + 623 (in place)
        err = jump(position + 623, 0, 1);
        if (err != cudaSuccess) return err;
        for (int have = 1; have < S; have *= 2) {
            int const cnt = std::min(have, S - have);
            err = jump(static_cast<uint64_t>(L) * static_cast<uint64_t>(have), have, cnt);
            if (err != cudaSuccess) return err;
        }
    }
    mt_segments_kernel<<<S, 256, 0, s>>>(states_dev, out, L, count);
    ++nl;
    if (launches) *launches += nl;
    return cudaGetLastError();
}

} // namespace pbk
