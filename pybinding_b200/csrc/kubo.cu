// kubo.cu -- K4: the Kubo-Bastin moment matrix  mu (M x M) += L (M x N) * R^H (N x M).
//
// Replaces MomentMultiplication::matrix_mul_add (cppcore/src/kpm/Moments.cpp:92-101,123-126), which
// the reference evaluates as a single-threaded Eigen GEMM on two host-resident M x N stacks.
// This is the one dense, tensor-core-shaped piece of the KPM path.  tcgen05/UMMA has no fp64 kind, so the
// fp64 tensor pipe is driven with mma.sync f64 -- on sm_100a every f64 shape compiles to DMMA.8x8x4, measured
// peak 37.1 TFLOP/s (tools/probe/dmma_probe.cu, profiles/r01_fp64_mma_probe.jsonl).  f32/c64 stacks are widened
// to fp64 when the fragments are read from shared memory, so every scalar type accumulates in double.
//
// Stack layout ("k-blocked", written directly by the step kernel's blocked destinations, step_common.cuh): the inner
// dimension (sites x lanes, real and imaginary parts interleaved for complex types) is cut into blocks of 256 bytes; block
// kb holds those 256 bytes of all M moment rows one after the other, each followed by a pad that makes the fragment reads
// bank-conflict free (row stride 288 bytes for doubles, 272 for floats):
//     element (m, k) at  kb * (M * RS) + m * RS + (k mod TKE) * sizeof(Real),   kb = k / TKE,  TKE = 256 / sizeof(Real)
// One pipeline stage of a 128-row operand tile is therefore ONE contiguous 36 KB range of global memory and arrives by one
// bulk copy (TMA engine, mbarrier complete_tx) already in its padded shared-memory form.  The row-major M x N stacks of
// the first version needed 256 separate 256-byte reads per stage, 12 MB apart: 4096 cp.async instructions per stage, poor
// DRAM page and TLB locality, and edge tiles that cost as much as full ones (cuBLAS on that layout: 22.4 TFLOP/s at
// M = 514, profiles/r02_cublas_dgemm.log).
//
// Complex stacks are treated as real M x 2N matrices:
//   Re(mu) = A' * B'^T                      with A' = [.. ar_k, ai_k ..], B' = [.. br_k, bi_k ..]
//   Im(mu) = A'' * B'^T                     with A'' = [.. ai_k, -ar_k ..]   (pair swap + negate on fragment load)
//
// Kernel: 128 x 128 tile per CTA; 8 warps (4 x 2, warp tile 32 x 64 = 32 DMMAs per k4-step out of 12 fragment loads);
// lane 0 of warp 0 issues the bulk copies of a stage.  Three stages; full / empty mbarriers per stage, so the warps
// never meet at a CTA barrier and may drift a stage apart (tools/probe/gemm_loop_probe.cu: this main loop sustains
// 36.9 TFLOP/s from shared memory).  M is rarely a multiple of the tile (the
// reference's num_moments is 4k+2): edge tiles copy only their valid rows, warps whose 32 x 64 region lies outside the
// matrix do not take part at all, and 8 x 8 blocks outside are skipped, so the cost follows ceil(M/8)^2 blocks.
// Split-K over the blocks with per-split partial tiles and a fixed-order reduction keeps the result deterministic.
#include "bulk_common.cuh"

#include <cstdlib>
#include <type_traits>

namespace pbk {

namespace {

constexpr int BM = 128, BN = 128;       // CTA tile
constexpr int GEMM_STAGES = 3;
constexpr int CONSUMER_WARPS = 8;       // 4 (m) x 2 (n), warp tile 32 x 64
constexpr int GEMM_THREADS = 32 * CONSUMER_WARPS;
constexpr int BLOCK_BYTES = 256;        // payload of one stack row inside a k-block
constexpr int EXT_MAX = 6;              // a remainder of up to this many rows / columns (M mod 128) is folded into the last full tiles

template<class Real> struct Geom {
    static constexpr int TKE = BLOCK_BYTES / sizeof(Real);        // k-extent of a block in elements
    static constexpr int PAD = sizeof(Real) == 8 ? 32 : 16;       // bytes: row stride = 36 doubles / 68 floats
    static constexpr int RS = BLOCK_BYTES + PAD;                  // bytes between rows, in global and in shared memory
    static constexpr int TILE = (BM + EXT_MAX) * RS;              // one operand, one stage (room for the folded remainder rows)
    static constexpr int STAGE = 2 * TILE;
    static constexpr int SMEM = GEMM_STAGES * STAGE + 16 * GEMM_STAGES;
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds_as_double(uint32_t addr, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ double lds_as_double(uint32_t addr, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return static_cast<double>(v); }

/// part[z][Mp][Mp] = A[:, blocks of split z] * B[:, same blocks]^T.  A, B: k-blocked stacks (see above), `bs` = bytes
/// between consecutive blocks (M * RS), `bpc` = blocks per split.  fold > 0: M = 128 t + fold with fold <= EXT_MAX and the
/// grid has t x t tiles; the CTAs of the last tile row (column) also compute the `fold` remaining rows (columns).
template<class Real, bool IMAG>
__global__ void __launch_bounds__(GEMM_THREADS, 1) kubo_gemm_kernel(const unsigned char* __restrict__ A, const unsigned char* __restrict__ B, int M,
                                                                    int64_t nblk, int64_t bs, double* __restrict__ part, int Mp, int64_t bpc, int fold, int chunk_rows) {
    using G = Geom<Real>;
    constexpr int ES = sizeof(Real);
    extern __shared__ __align__(128) unsigned char gemm_smem[];
    uint32_t const smem0 = smem_u32(gemm_smem);
    uint32_t const full0 = smem0 + GEMM_STAGES * G::STAGE, empty0 = full0 + 8 * GEMM_STAGES;

    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int64_t const kb0 = static_cast<int64_t>(blockIdx.z) * bpc;
    int64_t const kb1 = (kb0 + bpc < nblk) ? kb0 + bpc : nblk;
    int const nstage = static_cast<int>(kb1 - kb0);
    int const ext_a = (fold > 0 && blockIdx.y + 1 == gridDim.y) ? fold : 0;                 // folded remainder rows / columns of this CTA
    int const ext_b = (fold > 0 && blockIdx.x + 1 == gridDim.x) ? fold : 0;
    int const rows_a = (M - m0 < BM) ? M - m0 : BM, rows_b = (M - n0 < BN) ? M - n0 : BN;   // valid rows of the two 128-row operand tiles
    int const active_m = (rows_a + 31) / 32, active_n = (rows_b + 63) / 64;                // consumer warps with work: wm < active_m, wn < active_n

    if (tid == 0) {
        for (int st = 0; st < GEMM_STAGES; ++st) { mbar_init(full0 + 8 * st, 1u); mbar_init(empty0 + 8 * st, static_cast<uint32_t>(32 * active_m * active_n)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // ---- producer: lane 0 of warp 0 (always an active warp), two bulk copies per stage -- the valid rows of the A tile and
    //      of the B tile, each with the folded remainder rows, which follow the tile's rows in the block.  It refills, at the
    //      top of stage st, the slot that stage st - 1 occupied: the only wait in the CTA that involves all warps, and only
    //      warp 0 takes it (a dedicated producer warp would cap the kernel at 168 registers: 9 warps put 3 on one scheduler).
    uint32_t const bytes_a = static_cast<uint32_t>(rows_a + ext_a) * G::RS, bytes_b = static_cast<uint32_t>(rows_b + ext_b) * G::RS;
    uint32_t const chunk = static_cast<uint32_t>(chunk_rows) * G::RS;
    const unsigned char* const pa0 = A + kb0 * bs + static_cast<int64_t>(m0) * G::RS;
    const unsigned char* const pb0 = B + kb0 * bs + static_cast<int64_t>(n0) * G::RS;
    auto produce = [&](int st) {       // stage st -> slot st % GEMM_STAGES
        int const slot = st % GEMM_STAGES, round = st / GEMM_STAGES;
        if (round > 0) mbar_wait(empty0 + 8 * slot, static_cast<uint32_t>(round - 1) & 1u);
        uint32_t const dst = smem0 + slot * G::STAGE, bar = full0 + 8 * slot;
        mbar_expect_tx(bar, bytes_a + bytes_b);
        // each operand tile in pieces of `chunk` bytes: several bulk copies in flight fetch a tile faster than one large
        // copy does (matters for the light CTAs of edge tiles, which do little arithmetic per byte)
        const unsigned char* const pa = pa0 + st * bs;
        const unsigned char* const pb = pb0 + st * bs;
        for (uint32_t o = 0; o < bytes_a; o += chunk) bulk_g2s(dst + o, pa + o, bytes_a - o < chunk ? bytes_a - o : chunk, bar);
        for (uint32_t o = 0; o < bytes_b; o += chunk) bulk_g2s(dst + G::TILE + o, pb + o, bytes_b - o < chunk ? bytes_b - o : chunk, bar);
    };
    if (tid == 0) {
        for (int st = 0; st < GEMM_STAGES - 1 && st < nstage; ++st) produce(st);
    }

    int const wm = warp >> 1, wn = warp & 1;
    if (wm >= active_m || wn >= active_n) return;      // this warp's 32 x 64 region lies outside the matrix
    constexpr int NB = 8;                               // 8-column blocks per warp
    // 8 x 8 blocks of this warp that intersect the matrix (warp-uniform)
    int mi_cnt = (M - (m0 + wm * 32) + 7) / 8; mi_cnt = mi_cnt > 4 ? 4 : mi_cnt;
    int nj_cnt = (M - (n0 + wn * 64) + 7) / 8; nj_cnt = nj_cnt > NB ? NB : nj_cnt;

    double c[4][NB][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
    // folded remainder: the extra 8-row block (rows 128.. of the A tile) against this warp's share of the columns -- the
    // four warps of a column half take two 8-column blocks each -- and the extra 8-column block against two of the warp's
    // four row blocks; the corner block goes to warp 0
    double ea[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, eb[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, ec[2] = {0.0, 0.0};

    // fragment addresses: thread holds A[row = lane / 4][k = lane % 4] of each 8 x 4 block (B alike)
    int const fr = lane >> 2, fk = lane & 3;
    int const fkA = IMAG ? (fk ^ 1) : fk;                     // A'' = [ai, -ar]: neighbour element, sign by parity
    double const sgnA = (IMAG && (fk & 1)) ? -1.0 : 1.0;
    uint32_t const offA = (wm * 32 + fr) * G::RS + fkA * ES;
    uint32_t const offB = G::TILE + (wn * 64 + fr) * G::RS + fk * ES;
    // lanes whose row of the extra block lies past the remainder re-read its last valid row (their results are dropped)
    uint32_t const offAe = (BM + (fr < ext_a ? fr : (ext_a > 0 ? ext_a - 1 : 0))) * G::RS + fkA * ES;
    uint32_t const offBe = G::TILE + (BN + (fr < ext_b ? fr : (ext_b > 0 ? ext_b - 1 : 0))) * G::RS + fk * ES;
    uint32_t const offBx = G::TILE + (wn * 64 + wm * 16 + fr) * G::RS + fk * ES;     // column blocks 2 wm, 2 wm + 1 of this half
    uint32_t const offAx = (wm * 32 + wn * 16 + fr) * G::RS + fkA * ES;               // row blocks 2 wn, 2 wn + 1 of this warp

    // main loop: interior warps run the branch-free version, warps on the matrix edge skip the 8 x 8 blocks that lie
    // outside; EA / EB add the folded remainder rows / columns (interior tiles only)
    auto mainloop = [&](auto edge_tag, auto ea_tag, auto eb_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value, EA = decltype(ea_tag)::value, EB = decltype(eb_tag)::value;
        int slot = 0; uint32_t round = 0;
        for (int st = 0; st < nstage; ++st) {
            if (tid == 0 && st + GEMM_STAGES - 1 < nstage) produce(st + GEMM_STAGES - 1);
            __syncwarp();
            mbar_wait(full0 + 8 * slot, round & 1u);
            uint32_t const base = smem0 + slot * G::STAGE;
#pragma unroll
            for (int kk = 0; kk < G::TKE; kk += 4) {
                double a[4], b[NB];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sgnA * lds_as_double(base + offA + i * 8 * G::RS + kk * ES, Real{});
#pragma unroll
                for (int j = 0; j < NB; ++j) b[j] = lds_as_double(base + offB + j * 8 * G::RS + kk * ES, Real{});
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!EDGE || i < mi_cnt) {
#pragma unroll
                        for (int j = 0; j < NB; ++j)
                            if (!EDGE || j < nj_cnt) dmma_m8n8k4(c[i][j][0], c[i][j][1], a[i], b[j]);
                    }
                }
                double ae = 0.0, be = 0.0;
                if constexpr (EA) {
                    ae = sgnA * lds_as_double(base + offAe + kk * ES, Real{});
#pragma unroll
                    for (int t = 0; t < 2; ++t) dmma_m8n8k4(ea[t][0], ea[t][1], ae, lds_as_double(base + offBx + t * 8 * G::RS + kk * ES, Real{}));
                }
                if constexpr (EB) {
                    be = lds_as_double(base + offBe + kk * ES, Real{});
#pragma unroll
                    for (int t = 0; t < 2; ++t) dmma_m8n8k4(eb[t][0], eb[t][1], sgnA * lds_as_double(base + offAx + t * 8 * G::RS + kk * ES, Real{}), be);
                }
                if constexpr (EA && EB) { if (warp == 0) dmma_m8n8k4(ec[0], ec[1], ae, be); }
            }
            // hand the stage back: every lane orders its own reads before the refill (proxy fence) and arrives itself, so
            // the reader -> producer edge does not pass through another lane (compute-sanitizer's racecheck follows it)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(empty0 + 8 * slot);
            if (++slot == GEMM_STAGES) { slot = 0; ++round; }
        }
    };
    using Y = std::true_type; using N = std::false_type;
    if (mi_cnt != 4 || nj_cnt != NB) mainloop(Y{}, N{}, N{});
    else if (ext_a && ext_b) mainloop(N{}, Y{}, Y{});
    else if (ext_a) mainloop(N{}, Y{}, N{});
    else if (ext_b) mainloop(N{}, N{}, Y{});
    else mainloop(N{}, N{}, N{});

    double* out = part + static_cast<int64_t>(blockIdx.z) * Mp * Mp;
    int const orow = lane >> 2, ocol = (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            int const row = m0 + wm * 32 + i * 8 + orow;
            int const col = n0 + wn * 64 + j * 8 + ocol;
            *reinterpret_cast<double2*>(out + static_cast<int64_t>(row) * Mp + col) = make_double2(c[i][j][0], c[i][j][1]);
        }
    if (ext_a) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
            *reinterpret_cast<double2*>(out + static_cast<int64_t>(m0 + BM + orow) * Mp + n0 + wn * 64 + (wm * 2 + t) * 8 + ocol) = make_double2(ea[t][0], ea[t][1]);
    }
    if (ext_b) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
            *reinterpret_cast<double2*>(out + static_cast<int64_t>(m0 + wm * 32 + (wn * 2 + t) * 8 + orow) * Mp + n0 + BN + ocol) = make_double2(eb[t][0], eb[t][1]);
    }
    if (ext_a && ext_b && warp == 0)
        *reinterpret_cast<double2*>(out + static_cast<int64_t>(m0 + BM + orow) * Mp + n0 + BN + ocol) = make_double2(ec[0], ec[1]);
}

__global__ void kubo_reduce_kernel(const double* part, int ksplit, int Mp, int M, double* C, int comp) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * M) return;
    int const m = i / M, n = i % M;
    double s = 0.0;
    for (int z = 0; z < ksplit; ++z) s += part[(static_cast<int64_t>(z) * Mp + m) * Mp + n];
    C[static_cast<int64_t>(i) * 2 + comp] += s;
}

/// split-K plan: `tiles` x `tiles` x `ksplit` CTAs of `bpc` blocks each; `fold` = remainder rows folded into the last tiles
void gemm_plan(int M, int64_t nblk, int num_sms, int* Mp, int* tiles, int* fold, int* ksplit, int64_t* bpc) {
    int const up = (M + BM - 1) / BM;
    *Mp = up * BM;
    *fold = (M >= BM && M % BM >= 1 && M % BM <= EXT_MAX) ? M % BM : 0;
    *tiles = *fold ? M / BM : up;
    // a few waves of CTAs (one per SM).  The CTAs of full tiles dominate the run time, so their number is made a
    // multiple of the SM count; each split costs one 128 x 128 partial tile of traffic, negligible next to its k-range
    static int const waves = [] { char const* v = std::getenv("PBK_KUBO_WAVES"); int w = v ? std::atoi(v) : 16; return w < 1 ? 1 : w; }();
    int const full = (M / BM) * (M / BM);
    int const heavy = full > 0 ? full : (*tiles) * (*tiles);
    int64_t ks = static_cast<int64_t>(waves) * num_sms / heavy;
    int64_t const max_split = (nblk + 7) / 8;                          // at least 8 stages per CTA
    if (ks > max_split) ks = max_split;
    if (ks > 1024) ks = 1024;
    if (ks < 1) ks = 1;
    int64_t const per = (nblk + ks - 1) / ks;
    *ksplit = static_cast<int>((nblk + per - 1) / per);
    *bpc = per;
}

template<class Real>
cudaError_t gemm_t(const void* A, const void* B, int M, int64_t nblk, bool cplx, double* C, double* workspace, size_t workspace_bytes,
                   int num_sms, cudaStream_t s) {
    using G = Geom<Real>;
    if (reinterpret_cast<uintptr_t>(A) % 16 != 0 || reinterpret_cast<uintptr_t>(B) % 16 != 0 || nblk < 1) return cudaErrorInvalidValue;
    int Mp = 0, tiles = 0, fold = 0, ksplit = 0;
    int64_t bpc = 0;
    gemm_plan(M, nblk, num_sms, &Mp, &tiles, &fold, &ksplit, &bpc);
    if (workspace_bytes < sizeof(double) * static_cast<size_t>(ksplit) * Mp * Mp) return cudaErrorInvalidValue;
    static bool raised = false;
    auto const k_re = kubo_gemm_kernel<Real, false>;
    auto const k_im = kubo_gemm_kernel<Real, true>;
    if (!raised) {
        cudaError_t e = cudaFuncSetAttribute(k_re, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_im, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
        if (e != cudaSuccess) return e;
        raised = true;
    }
    dim3 const grid(tiles, tiles, ksplit);
    int chunk_rows = 256;  // rows per bulk copy: one copy per operand tile (experiment knob PBK_KUBO_CHUNK; pieces of 32 rows measured 5 % slower)
    { char const* v = std::getenv("PBK_KUBO_CHUNK"); if (v && std::atoi(v) > 0) chunk_rows = std::atoi(v); }
    int64_t const bs = static_cast<int64_t>(M) * G::RS;
    auto const* a = static_cast<const unsigned char*>(A);
    auto const* b = static_cast<const unsigned char*>(B);
    k_re<<<grid, GEMM_THREADS, G::SMEM, s>>>(a, b, M, nblk, bs, workspace, Mp, bpc, fold, chunk_rows);
    kubo_reduce_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(workspace, ksplit, Mp, M, C, 0);
    if (cplx) {
        k_im<<<grid, GEMM_THREADS, G::SMEM, s>>>(a, b, M, nblk, bs, workspace, Mp, bpc, fold, chunk_rows);
        kubo_reduce_kernel<<<(M * M + 255) / 256, 256, 0, s>>>(workspace, ksplit, Mp, M, C, 1);
    }
    return cudaGetLastError();
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

KuboStackLayout kubo_stack_layout(int dtype, int M, size_t vector_bytes) {
    bool const single = dtype == F32 || dtype == C64;
    KuboStackLayout l{};
    l.row_stride = single ? Geom<float>::RS : Geom<double>::RS;
    l.blocks = static_cast<int64_t>((vector_bytes + BLOCK_BYTES - 1) / BLOCK_BYTES);
    l.block_stride = static_cast<int64_t>(M) * l.row_stride;
    l.bytes = static_cast<size_t>(l.blocks) * static_cast<size_t>(l.block_stride);
    return l;
}

size_t kubo_gemm_workspace_bytes(int M, int64_t blocks, int num_sms) {
    int Mp = 0, tiles = 0, fold = 0, ksplit = 0;
    int64_t bpc = 0;
    gemm_plan(M, blocks, num_sms, &Mp, &tiles, &fold, &ksplit, &bpc);
    return sizeof(double) * static_cast<size_t>(ksplit) * Mp * Mp;
}

cudaError_t launch_kubo_gemm(int dtype, const void* A, const void* B, int M, int64_t N, KuboStackLayout const& layout, double* C_c128,
                             double* workspace, size_t workspace_bytes, int num_sms, cudaStream_t s, double* flops) {
    bool const cplx = dtype_complex(dtype);
    if (flops) *flops = 2.0 * M * M * static_cast<double>(N) * (cplx ? 4 : 1);
    switch (dtype) {
        case F32: return gemm_t<float>(A, B, M, layout.blocks, false, C_c128, workspace, workspace_bytes, num_sms, s);
        case C64: return gemm_t<float>(A, B, M, layout.blocks, true, C_c128, workspace, workspace_bytes, num_sms, s);
        case F64: return gemm_t<double>(A, B, M, layout.blocks, false, C_c128, workspace, workspace_bytes, num_sms, s);
        case C128: return gemm_t<double>(A, B, M, layout.blocks, true, C_c128, workspace, workspace_bytes, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
