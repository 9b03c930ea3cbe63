// kernels_res.cu -- K1, resident-tile variant: the fused Chebyshev step for lattices whose locality clusters are mostly
// surface (three-dimensional lattices: a breadth-first ball of 256 sites of a cubic lattice references 0.9 x as many
// rows outside as inside).
//
//     y[row, :] = sum_s val[row][s] * x[col[row][s], :]  -  y[row, :]
//     m2[r] += |x[row, r]|^2 ,  m3[r] += conj(y_new[row, r]) * x[row, r]          (f64 accumulators)
//
// Why a second variant.  `cheb_step_bulk` gathers x[col] with global loads that are served by L1 / L2.  ncu on the
// 256^3 cubic lattice (profiles/r02_ncu_cubic_r64.csv) shows that kernel bound by the L1 data pipe, not by HBM:
// l1tex__data_pipe_lsu_wavefronts at 89 %, one wavefront per 32-byte sector of every gathered row (916 M sectors, 891 M
// wavefronts), DRAM at 58 %.  Seven gathers per row through a pipe that moves 32 bytes per cycle cannot keep up with HBM.
// Shared memory moves 128 bytes per cycle through the same pipe, and the bulk-copy engine fills it without using the
// pipe at all.  So here a CTA makes the x rows of one tile *resident in shared memory* -- the tile's own rows by one
// bulk copy, the rows of its halo (the sorted list of outside rows its matrix elements reference, precomputed per tile)
// by one small bulk copy each -- and every gather becomes a conflict-free ld.shared.v4.  The matrix carries 16-bit
// *local codes* (position inside [own rows | halo rows]) instead of 32-bit global columns, eight to a 16-byte load.
// y and the matrix records stream through a small ring of bulk-copy stages exactly like in `cheb_step_bulk`.
//
// Tiles are row ranges of the locality ordering with own + halo rows <= the shared-memory capacity; the metadata
// builder (below) halves the tiles that do not fit.  Rows are narrow on purpose (64 bytes: 16 float lanes per pass) so
// that two or three CTAs per SM overlap one tile's load with another's arithmetic.
//
// Replaces, from the reference (cppcore/): compute::kpm_spmv_diagonal (include/compute/kernel_polynomial.hpp:288-323)
// with the batching of DefaultCompute (src/kpm/default/Compute.cpp:52-88) and the Diagonal collectors.
#include "bulk_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace pbk {

namespace {

#ifndef PBK_RES_HPT
#define PBK_RES_HPT 5      // halo row numbers a thread holds in registers one tile ahead (6: 12 bytes of spills, 0.749 of the roofline on the cubic lattice; 5: 0.766; 4: 0.751)
#endif
constexpr uint32_t RES_DESC_WIN = 64;                  // tile descriptors per window
constexpr uint32_t RES_DESC_CAP = 2 * RES_DESC_WIN;    // two windows in shared memory (2 KB)

struct ResDev {  // kernel parameters
    const ResTile* tiles; int ntiles;
    const int32_t* halo_rows;
    const unsigned char* codes;   // uint16 [rows][kc]
    const unsigned char* vals;    // T [rows][kvb / sizeof T]
    const void* x; void* y;
    int R, cpr, rpb, k, stages;
    uint32_t row_bytes, cb, kvb;  // bytes per row of a vector block / of the code records / of the value records
    uint32_t xs_bytes, stage_bytes;   // xs_bytes: ONE resident-tile buffer; there are `buffers` of them
    int buffers;                  // 2: the next tile is loaded while the current one is being processed
    int l2_prefetch;              // the producer pulls the streamed operands of its next tile into L2
    double* partials; unsigned* counter; double* mom; double* m01; int M; int n; int fin;
};

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}
/// the barrier receives one arrival (counted in its initial count) once all cp.async of this thread so far have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 t;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
    return t;
}

// K > 0: ELL width known at compile time; K == 0: any width (slots walked four at a time, same FMA order)
template<class T, int V, int K, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) cheb_step_res(ResDev a) {
    using CH = Chunk<T, V>;
    static_assert(sizeof(CH) == 16, "the resident kernel moves 16-byte chunks");
    constexpr int C = ST<T>::C;
    constexpr int NACC = V * C;
    extern __shared__ __align__(128) unsigned char dyn_smem[];

    const unsigned char* __restrict__ xg_base = static_cast<const unsigned char*>(a.x);
    CH* __restrict__ y = static_cast<CH*>(a.y);

    uint32_t const tid = threadIdx.x;
    uint32_t const cpr = a.cpr, rpb = a.rpb;
    uint32_t const tx = tid % cpr, ty = tid / cpr;
    bool const active = ty < rpb;
    uint32_t const S = a.stages;
    uint32_t const row_bytes = a.row_bytes;
    uint32_t const smem0 = smem_u32(dyn_smem);
    uint32_t const NB = static_cast<uint32_t>(a.buffers);
    uint32_t const xs0 = smem0;                          // resident tiles: own rows, then halo rows; NB buffers
    uint32_t const ring0 = smem0 + NB * a.xs_bytes;
    uint32_t const ring_end = ring0 + S * a.stage_bytes;
    uint32_t const full0 = ring_end;                     // full[S], empty[S], xbar[NB]
    uint32_t const empty_off = 8u * S;
    uint32_t const xbar0 = full0 + 16u * S;
    uint32_t const desc0 = xbar0 + 16u;                  // this CTA's tile descriptors (RES_DESC_CAP x 16 bytes)
    uint32_t const ybytes_full = rpb * row_bytes;        // stage layout: y | codes | values
    uint32_t const coff = ybytes_full, voff = ybytes_full + rpb * a.cb;

    if (tid == 0) {
        for (uint32_t st = 0; st < S; ++st) { mbar_init(full0 + 8u * st, 1u); mbar_init(full0 + empty_off + 8u * st, TPB / 32); }
        for (uint32_t b = 0; b < NB; ++b) mbar_init(xbar0 + 8u * b, TPB + 1u);   // every thread's halo chunks + thread 0's bulk copy of the own rows
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // The descriptors of the tiles this CTA works on (tile blockIdx.x + i gridDim.x is its i-th) are kept in shared memory:
    // producer and consumers read them with ld.shared instead of waiting for a global load at every tile switch (8 % of the
    // stall samples in profiles/r02_ncu_cubic_res_r16_v2.csv).  Two windows of RES_DESC_WIN descriptors alternate: when the
    // consumers enter window k, window k + 1 is loaded over window k - 1, so every look-up (the producer and the tile
    // prefetch run a few tiles ahead) is a branch-free ld.shared at index i mod 2 RES_DESC_WIN.
    auto load_window = [&](uint32_t first) {   // descriptors first .. first + RES_DESC_WIN - 1
        for (uint32_t i = tid; i < RES_DESC_WIN; i += TPB) {
            int64_t const t = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(first + i) * gridDim.x;
            if (t < a.ntiles) {
                int4 const d = __ldg(reinterpret_cast<const int4*>(a.tiles) + t);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(desc0 + 16u * ((first + i) & (RES_DESC_CAP - 1u))), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w) : "memory");
            }
        }
    };
    load_window(0u);
    load_window(RES_DESC_WIN);
    __syncthreads();
    auto tile_desc = [&](uint32_t i) {   // descriptor of this CTA's i-th tile (i within the two resident windows)
        uint4 const d = lds_u4(desc0 + 16u * (i & (RES_DESC_CAP - 1u)));
        return ResTile{static_cast<int32_t>(d.x), static_cast<int32_t>(d.y), static_cast<int32_t>(d.z), static_cast<int32_t>(d.w)};
    };
    uint32_t const my_tiles = static_cast<int>(blockIdx.x) < a.ntiles ? (static_cast<uint32_t>(a.ntiles) - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;

    // ---- producer (thread 0) of the y / record ring: walks (tile, block-iteration) in the consumers' order, S - 1 ahead ----
    // When it starts on a tile it also asks the bulk-copy engine to pull the streamed operands of its NEXT tile (y, codes,
    // values, own x rows: ~70 KB) into L2, so that the ring -- only S - 1 stages deep, shared memory belongs to the x rows --
    // is refilled at L2 latency rather than DRAM latency.
    uint32_t pi = 0;                                     // index of the tile being produced among this CTA's tiles
    int32_t prow = 0, pleft = 0;
    auto prefetch_streams = [&](uint32_t i) {
        if (i >= my_tiles || !a.l2_prefetch) return;
        ResTile const t = tile_desc(i);
        size_t const r0 = static_cast<size_t>(t.row0);
        uint32_t const nr = static_cast<uint32_t>(t.nrows);
        bulk_prefetch_l2(static_cast<const unsigned char*>(a.y) + r0 * row_bytes, nr * row_bytes);
        bulk_prefetch_l2(xg_base + r0 * row_bytes, nr * row_bytes);
        bulk_prefetch_l2(a.codes + r0 * a.cb, nr * a.cb);
        bulk_prefetch_l2(a.vals + r0 * a.kvb, nr * a.kvb);
    };
    if (tid == 0 && my_tiles > 0) { ResTile const t0 = tile_desc(0); prow = t0.row0; pleft = t0.nrows; prefetch_streams(1); }
    uint32_t psb = ring0, pfb = full0, pround = 0;
    auto produce = [&]() {
        if (pround > 0) mbar_wait(pfb + empty_off, (pround - 1u) & 1u);
        uint32_t const rows = static_cast<uint32_t>(pleft) < rpb ? static_cast<uint32_t>(pleft) : rpb;
        mbar_expect_tx(pfb, rows * (row_bytes + a.cb + a.kvb));
        bulk_g2s(psb, static_cast<const unsigned char*>(a.y) + static_cast<size_t>(prow) * row_bytes, rows * row_bytes, pfb);
        bulk_g2s(psb + coff, a.codes + static_cast<size_t>(prow) * a.cb, rows * a.cb, pfb);
        bulk_g2s(psb + voff, a.vals + static_cast<size_t>(prow) * a.kvb, rows * a.kvb, pfb);
        prow += static_cast<int32_t>(rows); pleft -= static_cast<int32_t>(rows);
        if (pleft == 0) {
            ++pi;
            if (pi < my_tiles) { ResTile const tn = tile_desc(pi); prow = tn.row0; pleft = tn.nrows; prefetch_streams(pi + 1u); }
        }
        psb += a.stage_bytes; pfb += 8u;
        if (psb == ring_end) { psb = ring0; pfb = full0; ++pround; }
    };
    if (tid == 0) {
        for (uint32_t i = 0; i + 1 < S && pi < my_tiles; ++i) produce();
    }

    double acc[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = 0.0;

    uint32_t sb = ring0, fb = full0, cph = 0;
    uint32_t const my_vec = tid * 16u;
    uint32_t const my_code = coff + ty * a.cb;
    uint32_t const my_val = voff + ty * a.kvb;
    uint32_t const kk = K > 0 ? static_cast<uint32_t>(K) : static_cast<uint32_t>(a.k);

    // The x rows of a tile -> resident buffer b: thread 0 copies the tile's own rows (one bulk copy); the halo rows are
    // scattered, so they move as 16-byte cp.async chunks -- thread (tx, ty) takes chunk tx of halo rows ty, ty + rpb, ... --
    // which a whole warp issues in ONE instruction (a bulk copy per 64-byte row went through the uniform datapath one lane
    // at a time: ~15 % of the kernel's stall samples sat in that loop, profiles/r02_ncu_cubic_res_r16_v1.csv).  Each thread
    // then attaches its copies to the buffer's barrier (cp.async.mbarrier.arrive.noinc), thread 0 adds the expectation of
    // the bulk copy: the phase completes exactly when everything has landed, no CTA barrier between issue and use.
    // The tile descriptor and the thread's first halo row numbers are fetched one tile ahead (`prefetch_tile`), so that
    // nothing waits for a global load when the copies are issued.
    constexpr uint32_t HPT = PBK_RES_HPT;                            // halo rows per thread held in registers (more: loaded late)
    ResTile nxt{0, 0, 0, 0};
    int32_t hidx[HPT];
    auto prefetch_tile = [&](uint32_t i) {
        nxt = tile_desc(i);
#pragma unroll
        for (uint32_t q = 0; q < HPT; ++q) {
            uint32_t const j = ty + q * rpb;
            hidx[q] = (active && j < static_cast<uint32_t>(nxt.nh)) ? __ldg(a.halo_rows + nxt.halo_off + j) : 0;
        }
    };
    auto issue_tile = [&](uint32_t b) {
        uint32_t const nrows = static_cast<uint32_t>(nxt.nrows), nh = static_cast<uint32_t>(nxt.nh);
        uint32_t const xs = xs0 + b * a.xs_bytes, xbar = xbar0 + 8u * b;
        if (tid == 0) {
            mbar_expect_tx(xbar, nrows * row_bytes);
            bulk_g2s(xs, xg_base + static_cast<size_t>(nxt.row0) * row_bytes, nrows * row_bytes, xbar);
        }
        if (active) {
            uint32_t const dst0 = xs + nrows * row_bytes + tx * 16u;
            const unsigned char* const src0 = xg_base + tx * 16u;
#pragma unroll
            for (uint32_t q = 0; q < HPT; ++q) {
                uint32_t const j = ty + q * rpb;
                if (j < nh) cp_async_16(dst0 + j * row_bytes, src0 + static_cast<size_t>(hidx[q]) * row_bytes);
            }
            for (uint32_t j = ty + HPT * rpb; j < nh; j += rpb) {
                int32_t const r = __ldg(a.halo_rows + nxt.halo_off + j);
                cp_async_16(dst0 + j * row_bytes, src0 + static_cast<size_t>(r) * row_bytes);
            }
        }
        cp_async_arrive_noinc(xbar);
    };
    if (my_tiles > 0u) { prefetch_tile(0u); issue_tile(0u); }
    if (NB == 2u && my_tiles > 1u) { prefetch_tile(1u); issue_tile(1u); }

    for (uint32_t it = 0; it < my_tiles; ++it) {
        // entering a new window: the one after it replaces the one just left (nobody reads that any more; the stores are
        // published by the CTA barrier at the end of this tile, long before the first look-up into the new window)
        if (it != 0u && (it & (RES_DESC_WIN - 1u)) == 0u) load_window(it + RES_DESC_WIN);
        ResTile const tl = tile_desc(it);
        uint32_t const nrows = static_cast<uint32_t>(tl.nrows);
        uint32_t const b = NB == 2u ? (it & 1u) : 0u;
        uint32_t const xs = xs0 + b * a.xs_bytes, xbar = xbar0 + 8u * b;
        uint32_t const xph = NB == 2u ? ((it >> 1) & 1u) : (it & 1u);
        uint32_t const itn = it + NB;                          // the tile that goes into this buffer next
        if (itn < my_tiles) prefetch_tile(itn);                // its halo row numbers: in flight during this tile
        bool xready = false;

        for (uint32_t r0 = 0; r0 < nrows; r0 += rpb) {
            if (tid == 0 && pi < my_tiles) produce();      // refill the stage consumed one iteration ago
            uint32_t const lrow = r0 + ty;                 // my row inside the tile
            bool const valid = active && lrow < nrows;
            uint32_t const ci = (static_cast<uint32_t>(tl.row0) + lrow) * cpr + tx;   // my chunk of y (launcher: < 2^32)

            mbar_wait(fb, cph);
            CH yv, xr, out;
            if constexpr (K > 0) {
                constexpr int NC = (K + 7) / 8;                          // 16-byte loads of codes
                constexpr int VPL = 16 / static_cast<int>(sizeof(T));    // values per 16-byte load
                constexpr int NV = (K + VPL - 1) / VPL;
                uint32_t code[NC * 8]; T v[NV * VPL];
                if (valid) {
                    yv = lds_chunk<CH>(sb + my_vec);
#pragma unroll
                    for (int q = 0; q < NC; ++q) {
                        uint4 const w = lds_u4(sb + my_code + 16u * q);
                        code[8 * q + 0] = w.x & 0xffffu; code[8 * q + 1] = w.x >> 16; code[8 * q + 2] = w.y & 0xffffu; code[8 * q + 3] = w.y >> 16;
                        code[8 * q + 4] = w.z & 0xffffu; code[8 * q + 5] = w.z >> 16; code[8 * q + 6] = w.w & 0xffffu; code[8 * q + 7] = w.w >> 16;
                    }
#pragma unroll
                    for (int q = 0; q < NV; ++q) {
                        Chunk<T, VPL> const c4 = lds_chunk<Chunk<T, VPL>>(sb + my_val + 16u * q);
#pragma unroll
                        for (int e = 0; e < VPL; ++e) v[q * VPL + e] = c4.e[e];
                    }
                }
                release_stage(fb + empty_off, tid);                      // this warp is done with the stage
                if (!xready) { mbar_wait(xbar, xph); xready = true; }
                if (valid) {
                    CH xg[K];
#pragma unroll
                    for (int s = 0; s < K; ++s) xg[s] = lds_chunk<CH>(xs + code[s] * row_bytes + tx * 16u);
                    xr = lds_chunk<CH>(xs + lrow * row_bytes + tx * 16u);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        T r = neg_(yv.e[e]);
#pragma unroll
                        for (int s = 0; s < K; ++s) r = fma_(v[s], xg[s].e[e], r);
                        out.e[e] = r;
                        sums_(acc + e * C, xr.e[e], r);
                    }
                    store_cs(y + ci, out);
                }
            } else {
                if (!xready) { mbar_wait(xbar, xph); xready = true; }
                if (valid) {
                    yv = lds_chunk<CH>(sb + my_vec);
#pragma unroll
                    for (int e = 0; e < V; ++e) out.e[e] = neg_(yv.e[e]);
                    for (uint32_t s0 = 0; s0 < kk; s0 += 4u) {
                        uint32_t code[4]; T v[4]; CH xg[4];
#pragma unroll
                        for (uint32_t j = 0; j < 4u; ++j) {
                            if (s0 + j < kk) { code[j] = lds_u16(sb + my_code + 2u * (s0 + j)); lds_val(sb + my_val + static_cast<uint32_t>(sizeof(T)) * (s0 + j), v[j]); }
                        }
#pragma unroll
                        for (uint32_t j = 0; j < 4u; ++j) { if (s0 + j < kk) xg[j] = lds_chunk<CH>(xs + code[j] * row_bytes + tx * 16u); }
#pragma unroll
                        for (uint32_t j = 0; j < 4u; ++j) {
                            if (s0 + j < kk) {
#pragma unroll
                                for (int e = 0; e < V; ++e) out.e[e] = fma_(v[j], xg[j].e[e], out.e[e]);
                            }
                        }
                    }
                    xr = lds_chunk<CH>(xs + lrow * row_bytes + tx * 16u);
                }
                release_stage(fb + empty_off, tid);
                if (valid) {
#pragma unroll
                    for (int e = 0; e < V; ++e) sums_(acc + e * C, xr.e[e], out.e[e]);
                    store_cs(y + ci, out);
                }
            }
            sb += a.stage_bytes; fb += 8u;
            if (sb == ring_end) { sb = ring0; fb = full0; cph ^= 1u; }
        }
        if (!xready) mbar_wait(xbar, xph);   // (a tile without rows cannot occur; keeps the phase bookkeeping exact anyway)
        // This buffer was read through the generic proxy and is refilled by the bulk-copy engine (async proxy): a proxy
        // fence by every reader, then the CTA barrier, orders the two.  With two buffers the refill (the tile after next)
        // overlaps the processing of the next tile, which is already resident or on its way.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (itn < my_tiles) issue_tile(b);
    }

    StepDev fin{};
    fin.R = a.R; fin.cpr = a.cpr; fin.rpb = a.rpb;
    fin.partials = a.partials; fin.counter = a.counter; fin.mom = a.mom; fin.m01 = a.m01; fin.M = a.M; fin.n = a.n; fin.fin = a.fin;
    finish_sums<C, NACC, TPB>(fin, acc, static_cast<int>(tx), static_cast<int>(ty));
}

// ------------------------------------------------------------------------------------------------
// Metadata: per tile the sorted list of outside rows its matrix elements reference (the halo) and, per matrix
// element, the 16-bit local code (own rows first, then the halo list).  One CTA per tile; the distinct outside
// columns are collected in a shared-memory hash set, sorted (bitonic) and looked up by binary search.
// ------------------------------------------------------------------------------------------------
constexpr int META_TPB = 256;
constexpr int META_HT = 16384;      // hash slots (64 KB): up to RES_MAX_HALO distinct outside rows per tile
constexpr int RES_MAX_HALO = 6144;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

/// collects the distinct outside columns of the tile in `table` (META_HT slots, -1 = empty); returns their number
/// (or RES_MAX_HALO + 1 when there are too many) in *count
__device__ void collect_halo(const int32_t* __restrict__ col, int64_t pitch, int k, ResTile tl, int32_t* table, int* count) {
    for (int i = threadIdx.x; i < META_HT; i += blockDim.x) table[i] = -1;
    if (threadIdx.x == 0) *count = 0;
    __syncthreads();
    int const total = tl.nrows * k;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int const s = i / tl.nrows, r = i - s * tl.nrows;
        int32_t const c = col[static_cast<int64_t>(s) * pitch + tl.row0 + r];
        if (c >= tl.row0 && c < tl.row0 + tl.nrows) continue;
        uint32_t h = hash32(static_cast<uint32_t>(c)) & (META_HT - 1);
        for (;;) {
            if (*reinterpret_cast<volatile int*>(count) > RES_MAX_HALO) break;
            int32_t const old = atomicCAS(table + h, -1, c);
            if (old == -1) { atomicAdd(count, 1); break; }
            if (old == c) break;
            h = (h + 1) & (META_HT - 1);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(META_TPB) res_count_kernel(const int32_t* __restrict__ col, int64_t pitch, int k,
                                                            const ResTile* __restrict__ tiles, int32_t* __restrict__ nh_out) {
    extern __shared__ int32_t meta_smem[];
    __shared__ int count;
    collect_halo(col, pitch, k, tiles[blockIdx.x], meta_smem, &count);
    if (threadIdx.x == 0) nh_out[blockIdx.x] = count;
}

template<class T>
__global__ void __launch_bounds__(META_TPB) res_fill_kernel(const T* __restrict__ val, const int32_t* __restrict__ col, int64_t pitch, int k,
                                                           const ResTile* __restrict__ tiles, int32_t* __restrict__ halo_rows,
                                                           unsigned char* __restrict__ codes, unsigned char* __restrict__ vals,
                                                           uint32_t cb, uint32_t kvb) {
    extern __shared__ int32_t meta_smem[];
    int32_t* table = meta_smem;                 // META_HT
    int32_t* list = meta_smem + META_HT;        // up to 8192 (power of two >= nh)
    __shared__ int count, filled;
    ResTile const tl = tiles[blockIdx.x];
    collect_halo(col, pitch, k, tl, table, &count);
    int const nh = count;
    int p2 = 1;
    while (p2 < nh) p2 <<= 1;
    if (threadIdx.x == 0) filled = 0;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) list[i] = 0x7fffffff;
    __syncthreads();
    for (int i = threadIdx.x; i < META_HT; i += blockDim.x) {
        int32_t const c = table[i];
        if (c >= 0) list[atomicAdd(&filled, 1)] = c;
    }
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1) {           // bitonic sort, ascending
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                int const j = i ^ stride;
                if (j > i) {
                    bool const up = (i & size) == 0;
                    int32_t const a = list[i], b = list[j];
                    if ((a > b) == up) { list[i] = b; list[j] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < nh; j += blockDim.x) halo_rows[tl.halo_off + j] = list[j];
    // codes and values, one thread per row of the tile
    uint32_t const kc = cb / 2u, kv = kvb / static_cast<uint32_t>(sizeof(T));
    for (int r = threadIdx.x; r < tl.nrows; r += blockDim.x) {
        int64_t const row = static_cast<int64_t>(tl.row0) + r;
        uint16_t* cp = reinterpret_cast<uint16_t*>(codes + row * cb);
        T* vp = reinterpret_cast<T*>(vals + row * kvb);
        for (int s = 0; s < k; ++s) {
            int32_t const c = col[static_cast<int64_t>(s) * pitch + row];
            int code;
            if (c >= tl.row0 && c < tl.row0 + tl.nrows) code = c - tl.row0;
            else {
                int lo = 0, hi = nh;
                while (lo < hi) { int const mid = (lo + hi) >> 1; if (list[mid] < c) lo = mid + 1; else hi = mid; }
                code = tl.nrows + lo;
            }
            cp[s] = static_cast<uint16_t>(code);
            vp[s] = val[static_cast<int64_t>(s) * pitch + row];
        }
        for (uint32_t s = k; s < kc; ++s) cp[s] = static_cast<uint16_t>(r);
        for (uint32_t s = k; s < kv; ++s) vp[s] = zero_(T{});
    }
}

using ResKernel = void (*)(ResDev);
constexpr int RES_TPB = 256;
constexpr int RES_MAX_DYN = 224 * 1024;   // + 2.2 KB of static shared memory (finish_sums) <= the 227 KB a CTA may own

cudaError_t raise_res_limit(ResKernel fn) {
    static std::mutex mutex;
    static std::map<ResKernel, bool> raised;
    std::lock_guard<std::mutex> lock(mutex);
    if (raised[fn]) return cudaSuccess;
    cudaError_t const err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_MAX_DYN);
    if (err == cudaSuccess) raised[fn] = true;
    return err;
}

template<class T, int V>
ResKernel res_kernel_k(int k) {
    switch (k) {
        case 3: return cheb_step_res<T, V, 3, RES_TPB, 3>;
        case 4: return cheb_step_res<T, V, 4, RES_TPB, 3>;
        case 7: return cheb_step_res<T, V, 7, RES_TPB, 3>;
        default: return cheb_step_res<T, V, 0, RES_TPB, 3>;
    }
}

template<class T>
cudaError_t launch_res_t(ResArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled) {
    constexpr int V = 16 / sizeof(T);
    *handled = false;
    if (a.R % V != 0) return cudaSuccess;
    int const cpr = a.R / V;
    if (static_cast<uint32_t>(a.R) * sizeof(T) != a.geo.row_bytes || cpr > RES_TPB) return cudaSuccess;
    int const rpb = RES_TPB / cpr;
    if ((a.nrows + rpb) * static_cast<int64_t>(cpr) >= (int64_t{1} << 32)) return cudaSuccess;
    ResKernel const fn = res_kernel_k<T, V>(a.k);
    cudaError_t err = raise_res_limit(fn);
    if (err != cudaSuccess) return err;
    uint32_t const stage_bytes = (static_cast<uint32_t>(rpb) * (a.geo.row_bytes + a.geo.cb + a.geo.kvb) + 127u) / 128u * 128u;
    int const dyn = static_cast<int>(a.geo.buffers * a.geo.xs_bytes + a.geo.stages * stage_bytes + 16u * a.geo.stages + 16u + 16u * RES_DESC_CAP);
    if (dyn > RES_MAX_DYN) return cudaSuccess;
    int grid = num_sms * a.geo.ctas_per_sm;
    if (grid > a.ntiles) grid = a.ntiles;
    if (grid > max_step_blocks(num_sms)) grid = max_step_blocks(num_sms);
    ResDev d{};
    d.tiles = a.tiles; d.ntiles = a.ntiles; d.halo_rows = a.halo_rows;
    d.codes = static_cast<const unsigned char*>(a.codes); d.vals = static_cast<const unsigned char*>(a.vals);
    d.x = a.x; d.y = a.y; d.R = a.R; d.cpr = cpr; d.rpb = rpb; d.k = a.k; d.stages = a.geo.stages;
    d.row_bytes = a.geo.row_bytes; d.cb = a.geo.cb; d.kvb = a.geo.kvb; d.xs_bytes = a.geo.xs_bytes; d.stage_bytes = stage_bytes;
    d.buffers = a.geo.buffers;
    { char const* v = std::getenv("PBK_RES_L2PF"); d.l2_prefetch = v ? std::atoi(v) : 0; }   // experiment knob, read per launch (off: the prefetches queue ahead of the ring copies in the same engine, 0.68 vs 0.78 of the roofline)
    d.partials = a.partials; d.counter = a.counter; d.mom = a.mom; d.m01 = a.m01; d.M = a.M; d.n = a.n; d.fin = a.fin;
    fn<<<grid, RES_TPB, dyn, stream>>>(d);
    *handled = true;
    if (info) { info->grid = grid; info->block = RES_TPB; info->V = V; info->K = a.k; info->bulk = a.geo.stages; info->res = 1; }
    return cudaGetLastError();
}

} // anonymous namespace

ResGeometry res_geometry(int dtype, int k, int lanes, int ctas_per_sm, int stages, int buffers) {
    ResGeometry g{};
    g.buffers = buffers >= 2 ? 2 : 1;
    uint32_t const s = static_cast<uint32_t>(dtype_size(dtype));
    g.row_bytes = static_cast<uint32_t>(lanes) * s;
    g.cb = (2u * static_cast<uint32_t>(k) + 15u) / 16u * 16u;
    g.kvb = (s * static_cast<uint32_t>(k) + 15u) / 16u * 16u;
    g.stages = stages < 2 ? 2 : (stages > 8 ? 8 : stages);
    g.ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
    uint32_t const cpr = g.row_bytes / 16u;
    uint32_t const rpb = cpr ? RES_TPB / cpr : 0;
    uint32_t const stage_bytes = (rpb * (g.row_bytes + g.cb + g.kvb) + 127u) / 128u * 128u;
    // shared memory of one SM (227 KB opt-in, 1 KB reserved per CTA) shared by the resident CTAs; finish_sums holds 2 KB statically
    uint32_t const per_cta = (227u * 1024u) / static_cast<uint32_t>(g.ctas_per_sm) - 1024u - 2304u;
    // the ring gives way to the resident rows: wide matrix records (large k, 16-byte scalars) fall back to a shallower ring
    // rather than leaving no room for a tile
    for (;; --g.stages) {
        uint32_t const fixed = g.stages * stage_bytes + 16u * g.stages + 16u + 16u * RES_DESC_CAP;
        g.xs_bytes = per_cta > fixed + 4096u ? (per_cta - fixed) / static_cast<uint32_t>(g.buffers) / 128u * 128u : 0u;
        g.cap_rows = g.row_bytes ? static_cast<int>(g.xs_bytes / g.row_bytes) : 0;
        if (g.stages == 2 || g.cap_rows >= 4 * static_cast<int>(rpb)) break;
    }
    if (g.cap_rows > 65535) { g.cap_rows = 65535; }     // 16-bit local codes
    g.rows_per_iteration = static_cast<int>(rpb);
    return g;
}

cudaError_t launch_res_count(EllDev const& h, const ResTile* tiles_dev, int ntiles, int32_t* nh_dev, cudaStream_t s) {
    static bool raised = false;
    size_t const smem = sizeof(int32_t) * META_HT;
    if (!raised) { cudaError_t e = cudaFuncSetAttribute(res_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); if (e != cudaSuccess) return e; raised = true; }
    res_count_kernel<<<ntiles, META_TPB, smem, s>>>(h.col, h.pitch, h.k, tiles_dev, nh_dev);
    return cudaGetLastError();
}

int res_max_halo() { return RES_MAX_HALO; }

cudaError_t launch_res_fill(int dtype, EllDev const& h, const ResTile* tiles_dev, int ntiles, int32_t* halo_rows, void* codes, void* vals,
                            ResGeometry const& g, cudaStream_t s) {
    size_t const smem = sizeof(int32_t) * (META_HT + 8192);
    auto* cd = static_cast<unsigned char*>(codes);
    auto* vd = static_cast<unsigned char*>(vals);
#define PBK_RES_FILL(T) { cudaError_t e = cudaFuncSetAttribute(res_fill_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); if (e != cudaSuccess) return e; \
        res_fill_kernel<T><<<ntiles, META_TPB, smem, s>>>(static_cast<const T*>(h.val), h.col, h.pitch, h.k, tiles_dev, halo_rows, cd, vd, g.cb, g.kvb); }
    switch (dtype) {
        case F32: PBK_RES_FILL(float) break;
        case C64: PBK_RES_FILL(float2) break;
        case F64: PBK_RES_FILL(double) break;
        case C128: PBK_RES_FILL(double2) break;
        default: return cudaErrorInvalidValue;
    }
#undef PBK_RES_FILL
    return cudaGetLastError();
}

cudaError_t launch_step_res(int dtype, ResArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled) {
    *handled = false;
    if (!a.tiles || a.ntiles <= 0 || a.nrows <= 0) return cudaSuccess;
    switch (dtype) {
        case F32: return launch_res_t<float>(a, num_sms, stream, info, handled);
        case C64: return launch_res_t<float2>(a, num_sms, stream, info, handled);
        case F64: return launch_res_t<double>(a, num_sms, stream, info, handled);
        case C128: return launch_res_t<double2>(a, num_sms, stream, info, handled);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
