// pair_v2_regs.cu -- REGISTER-PRESSURE STUDY for the next version of the two-step kernel (not part of libpbkpm.so).
//
// Same algorithm and arithmetic as pybinding_b200/csrc/kernels_pair.cu (`cheb_pair_bulk`), with the three source-level changes
// that the round-1 ncu capture called for (profiles/r01_ncu_pair_200nm_r32_v4.csv: 117 registers, 16 warps/SM, latency-bound):
//   1. the f64 sums of a phase live in registers only inside that phase; between phases they are parked in (volatile) local
//      memory -- one 96-byte store + load per tile and thread instead of 24 registers held for the whole kernel,
//   2. the producer's cursor (tile, phase, row, ring position) lives in shared memory: only thread 0 ever reads it,
//   3. no software prefetch of the next halo row's record; instead (PAIR_HALO_BULK=1, default) every thread bulk-copies one
//      halo row's a-chunk and H record into shared memory at the start of the tile (own mbarrier, proxy fence after the
//      previous tile's generic accesses), so the halo loop only waits for gathers and overwrites a with c in place.
// `make -C tools/probe pair_v2_regs` prints ptxas' register / spill numbers for complex64, ELL width 3:
//
//   variant (float2, K = 3, 256 threads)         __launch_bounds__(256, 2)   (256, 3)              (256, 4)
//   kernels_pair.cu (round 1)                    117 regs, no spills         80 regs, 208 B spills  64 regs, 388 B spills
//   + sums parked between phases                 117                         80, 60 B               64, 200 B
//   + no halo record prefetch                    107                         80, none               64, 184 B
//   + producer cursor in shared memory           101                         80, none               64, 32 B
//   + halo rows bulk-prefetched (this file)      103                         80, none               64, 44 B
//
// i.e. the kernel fits 3 CTAs/SM without spills and 4 CTAs/SM (the occupancy of the single-step kernel) with 32 bytes of
// spills.  Unverified on hardware: compile-only.  The file is a drop-in candidate for kernels_pair.cu (same PairArgs / launcher
// interface; shared memory: stage ring, barriers, halo buffer, staged halo records).
// Next step (round 2): swap it in, run the pair parity tests, and re-measure at R = 32 with 3 pipeline stages.
#include "bulk_common.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace pbk {

namespace {

struct PairDev {
    const unsigned char* packed;    // phase-1 records: global column ids
    const unsigned char* packed2;   // phase-2 records: column codes
    const int32_t* halo_ptr;        // [tiles + 1]
    const int32_t* halo_rows;       // global row ids of each tile's halo, ascending
    const void* a; const void* b; void* c; void* d;
    int nrows, ntiles, cpr, rpb, tile;   // tile: rows per locality cluster (any multiple of 1; block-iterations cover rpb rows)
    int R, stages;
    uint32_t rec, valoff, stage_bytes, halo_off;   // halo buffer starts halo_off bytes into dynamic shared memory
    uint32_t hrec_off;                             // PAIR_HALO_BULK: staged H records of the halo rows start here
    double* partials; unsigned* counter; double* mom; double* m01; int M; int n;
};

template<class CH> __device__ __forceinline__ void sts_chunk(uint32_t addr, CH const& v) {
    int4 const t = *reinterpret_cast<const int4*>(&v);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
}
template<class CH> __device__ __forceinline__ CH lds_chunk_sync(uint32_t addr) {  // data written by other threads of the CTA
    int4 t;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr) : "memory");
    return *reinterpret_cast<CH*>(&t);
}
/// coherent global load (ld.global, L1-allocating): c rows written earlier by this CTA, ordered by bar.sync
template<class CH> __device__ __forceinline__ CH load_coherent(const CH* p) {
    int4 t;
    asm volatile("ld.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p) : "memory");
    return *reinterpret_cast<CH*>(&t);
}
template<class CH> __device__ __forceinline__ void store_plain(CH* p, CH const& v) {  // stays in L2 for the phase-2 reads
    int4 const t = *reinterpret_cast<const int4*>(&v);
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
}
template<class T> __device__ __forceinline__ T ldg_val(const unsigned char* p) { return ldg_scalar(reinterpret_cast<const T*>(p)); }

#ifndef PAIR_HALO_BULK
#define PAIR_HALO_BULK 1   // 1: the halo rows' a-chunks and H records are bulk-prefetched at the start of the tile
#endif

template<class T, int V, int K, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) cheb_pair_bulk(PairDev a) {
    using CH = Chunk<T, V>;
    static_assert(sizeof(CH) == 16, "the staged kernel moves 16-byte chunks");
    constexpr int C = ST<T>::C;
    constexpr int NACC = V * C;
    constexpr uint32_t HALF = TPB * 16u;   // bytes of one staged vector operand
    constexpr uint32_t HOFF = 2u * HALF;   // H records follow the two vector slots of a stage
    extern __shared__ __align__(128) unsigned char dyn_smem[];

    const CH* __restrict__ va = static_cast<const CH*>(a.a);
    const CH* __restrict__ vb = static_cast<const CH*>(a.b);
    CH* vc = static_cast<CH*>(a.c);
    CH* __restrict__ vd = static_cast<CH*>(a.d);

    uint32_t const tid = threadIdx.x;
    uint32_t const cpr = a.cpr, rpb = a.rpb;
    uint32_t const tile_rows = static_cast<uint32_t>(a.tile);
    uint32_t const tx = tid % cpr, ty = tid / cpr;
    bool const active = ty < rpb;
    uint32_t const S = a.stages;
    uint32_t const stage_bytes = a.stage_bytes;
    uint32_t const smem0 = smem_u32(dyn_smem);
    uint32_t const ring_end = smem0 + S * stage_bytes;
    uint32_t const full0 = ring_end;          // full[S] then empty[S], 8 bytes each
    uint32_t const empty_off = 8u * S;
    uint32_t const halo0 = smem0 + a.halo_off;
    uint32_t const nrows = static_cast<uint32_t>(a.nrows);
    int const ntiles = a.ntiles;

    if (tid == 0) {
        for (uint32_t st = 0; st < S; ++st) { mbar_init(full0 + 8u * st, 1u); mbar_init(full0 + empty_off + 8u * st, TPB / 32); }
        mbar_init(full0 + 16u * S, 1u);   // halo prefetch barrier
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // rows of a tile (the last tile may be short); a tile is walked in block-iterations of rpb rows
    auto rows_of = [&](int tile) {
        uint32_t const left = nrows - static_cast<uint32_t>(tile) * tile_rows;
        return left < tile_rows ? left : tile_rows;
    };

    // ---- producer (thread 0): walks the same (tile, phase, iteration) sequence S - 1 stages ahead ----
    struct Prod { int ptile, pphase; uint32_t prow, prows, psb, pfb, pround; };
    __shared__ Prod prod_state;
    if (tid == 0) { prod_state.ptile = blockIdx.x; prod_state.pphase = 0; prod_state.prow = 0; prod_state.prows = static_cast<int>(blockIdx.x) < ntiles ? rows_of(blockIdx.x) : 0; prod_state.psb = smem0; prod_state.pfb = full0; prod_state.pround = 0; }
    auto produce = [&]() {
        int ptile = prod_state.ptile, pphase = prod_state.pphase;
        uint32_t prow = prod_state.prow, prows = prod_state.prows, psb = prod_state.psb, pfb = prod_state.pfb, pround = prod_state.pround;
        if (pround > 0) mbar_wait(pfb + empty_off, (pround - 1u) & 1u);
        uint32_t const row0 = static_cast<uint32_t>(ptile) * tile_rows + prow;
        uint32_t const pc0 = row0 * cpr;
        uint32_t const left = prows - prow;
        uint32_t const nr = left < rpb ? left : rpb;
        uint32_t const vbytes = nr * cpr * 16u;
        uint32_t const hbytes = nr * a.rec;
        size_t const rec_off = static_cast<size_t>(row0) * a.rec;
        if (pphase == 0) {
            mbar_expect_tx(pfb, 2u * vbytes + hbytes);
            bulk_g2s(psb, va + pc0, vbytes, pfb);
            bulk_g2s(psb + HALF, vb + pc0, vbytes, pfb);
            bulk_g2s(psb + HOFF, a.packed + rec_off, hbytes, pfb);
        } else {
            mbar_expect_tx(pfb, vbytes + hbytes);
            bulk_g2s(psb, vb + pc0, vbytes, pfb);
            bulk_g2s(psb + HOFF, a.packed2 + rec_off, hbytes, pfb);
        }
        prow += nr;
        if (prow == prows) {
            prow = 0;
            if (pphase == 0) { pphase = 1; }
            else { pphase = 0; ptile += gridDim.x; prows = ptile < ntiles ? rows_of(ptile) : 0; }
        }
        psb += stage_bytes; pfb += 8u;
        if (psb == ring_end) { psb = smem0; pfb = full0; ++pround; }
        prod_state.ptile = ptile; prod_state.pphase = pphase; prod_state.prow = prow; prod_state.prows = prows; prod_state.psb = psb; prod_state.pfb = pfb; prod_state.pround = pround;
    };
    if (tid == 0) {
        for (uint32_t i = 0; i + 1 < S && prod_state.ptile < ntiles; ++i) produce();
    }

    volatile double park1[NACC], park2[NACC];   // f64 sums live in registers only inside their own phase
#pragma unroll
    for (int q = 0; q < NACC; ++q) { park1[q] = 0.0; park2[q] = 0.0; }

    uint32_t sb = smem0, fb = full0, cph = 0;
    uint32_t hph = 0;   // phase of the halo prefetch barrier
    auto next_stage = [&]() {
        sb += stage_bytes; fb += 8u;
        if (sb == ring_end) { sb = smem0; fb = full0; cph ^= 1u; }
    };
    uint32_t const my_vec = tid * 16u;
    uint32_t const my_rec = HOFF + ty * a.rec;
    uint32_t const my_val = my_rec + a.valoff;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t const trows = rows_of(tile);
        uint32_t const tile_row0 = static_cast<uint32_t>(tile) * tile_rows;
        int const hp = __ldg(a.halo_ptr + tile);
        int const nh = __ldg(a.halo_ptr + tile + 1) - hp;
#if PAIR_HALO_BULK
        __syncthreads();   // every thread is done with the previous tile's halo buffer (generic-proxy reads and writes)
        {
            uint32_t const hbar = full0 + 16u * S;
            uint32_t const row_bytes = cpr * 16u;
            if (tid == 0) mbar_expect_tx(hbar, static_cast<uint32_t>(nh) * (row_bytes + a.rec));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // order the generic accesses before the async writes
            for (int j = static_cast<int>(tid); j < nh; j += TPB) {
                uint32_t const hrow = static_cast<uint32_t>(__ldg(a.halo_rows + hp + j));
                bulk_g2s(halo0 + static_cast<uint32_t>(j) * row_bytes, va + hrow * cpr, row_bytes, hbar);
                bulk_g2s(smem0 + a.hrec_off + static_cast<uint32_t>(j) * a.rec, a.packed + static_cast<size_t>(hrow) * a.rec, a.rec, hbar);
            }
        }
#endif

        // ---- phase 1, own rows: c = H b - a, sums |b|^2 and conj(c) b ----
        {
        double acc1[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc1[q] = park1[q];
        for (uint32_t w0 = 0; w0 < trows; w0 += rpb) {
            if (tid == 0 && prod_state.ptile < ntiles) produce();
            uint32_t const ci = (tile_row0 + w0 + ty) * cpr + tx;
            bool const valid = active && w0 + ty < trows;
            mbar_wait(fb, cph);
            CH yv, xr;
            int32_t c[K]; T v[K];
            if (valid) {
                yv = lds_chunk<CH>(sb + my_vec);
                xr = lds_chunk<CH>(sb + HALF + my_vec);
#pragma unroll
                for (int s = 0; s < K; ++s) { c[s] = lds_i32(sb + my_rec + 4u * s); lds_val(sb + my_val + static_cast<uint32_t>(sizeof(T)) * s, v[s]); }
            }
            __syncwarp();
            if ((tid & 31u) == 0) mbar_arrive(fb + empty_off);
            if (valid) {
                CH xg[K];
#pragma unroll
                for (int s = 0; s < K; ++s) xg[s] = load_nc(vb + (static_cast<uint32_t>(c[s]) * cpr + tx));
                CH out;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T r = neg_(yv.e[e]);
#pragma unroll
                    for (int s = 0; s < K; ++s) r = fma_(v[s], xg[s].e[e], r);
                    out.e[e] = r;
                    sums_(acc1 + e * C, xr.e[e], r);
                }
                store_plain(vc + ci, out);
            }
            next_stage();
        }

#pragma unroll
        for (int q = 0; q < NACC; ++q) park1[q] = acc1[q];
        }
#if !PAIR_HALO_BULK
        __syncthreads();   // every thread is done with the previous tile's halo buffer
#endif

        // ---- phase 1, halo rows: c into shared memory only ----
        {
#if PAIR_HALO_BULK
            mbar_wait(full0 + 16u * S, hph);
            hph ^= 1u;
#endif
            for (int j = static_cast<int>(ty); active && j < nh; j += static_cast<int>(rpb)) {
                int32_t c[K]; T v[K];
#if PAIR_HALO_BULK
                uint32_t const rp = smem0 + a.hrec_off + static_cast<uint32_t>(j) * a.rec;
                uint32_t const my = halo0 + (static_cast<uint32_t>(j) * cpr + tx) * 16u;
#pragma unroll
                for (int s = 0; s < K; ++s) { c[s] = lds_i32(rp + 4u * s); lds_val(rp + a.valoff + static_cast<uint32_t>(sizeof(T)) * s, v[s]); }
                CH const yv = lds_chunk_sync<CH>(my);   // a[halo row], overwritten in place by c below (this thread owns the chunk)
#else
                uint32_t const hrow = static_cast<uint32_t>(__ldg(a.halo_rows + hp + j));
                const unsigned char* rp = a.packed + static_cast<size_t>(hrow) * a.rec;
                uint32_t const my = halo0 + (static_cast<uint32_t>(j) * cpr + tx) * 16u;
#pragma unroll
                for (int s = 0; s < K; ++s) { c[s] = __ldg(reinterpret_cast<const int32_t*>(rp) + s); v[s] = ldg_val<T>(rp + a.valoff + sizeof(T) * s); }
                CH const yv = load_nc(va + (hrow * cpr + tx));
#endif
                CH xg[K];
#pragma unroll
                for (int s = 0; s < K; ++s) xg[s] = load_nc(vb + (static_cast<uint32_t>(c[s]) * cpr + tx));
                CH out;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T r = neg_(yv.e[e]);
#pragma unroll
                    for (int s = 0; s < K; ++s) r = fma_(v[s], xg[s].e[e], r);
                    out.e[e] = r;
                }
                sts_chunk(my, out);
            }
        }

        __syncthreads();   // c: own rows visible in global memory (CTA scope), halo rows in shared memory

        // ---- phase 2, own rows: d = H c - b, sums |c|^2 and conj(d) c ----
        {
        double acc2[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc2[q] = park2[q];
        for (uint32_t w0 = 0; w0 < trows; w0 += rpb) {
            if (tid == 0 && prod_state.ptile < ntiles) produce();
            uint32_t const ci = (tile_row0 + w0 + ty) * cpr + tx;
            bool const valid = active && w0 + ty < trows;
            CH xr;
            if (valid) xr = load_coherent(vc + ci);
            mbar_wait(fb, cph);
            CH yv;
            int32_t c[K]; T v[K];
            if (valid) {
                yv = lds_chunk<CH>(sb + my_vec);
#pragma unroll
                for (int s = 0; s < K; ++s) { c[s] = lds_i32(sb + my_rec + 4u * s); lds_val(sb + my_val + static_cast<uint32_t>(sizeof(T)) * s, v[s]); }
            }
            __syncwarp();
            if ((tid & 31u) == 0) mbar_arrive(fb + empty_off);
            if (valid) {
                CH xg[K];
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    if (c[s] >= 0) xg[s] = load_coherent(vc + (static_cast<uint32_t>(c[s]) * cpr + tx));
                    else xg[s] = lds_chunk_sync<CH>(halo0 + (static_cast<uint32_t>(-1 - c[s]) * cpr + tx) * 16u);
                }
                CH out;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T r = neg_(yv.e[e]);
#pragma unroll
                    for (int s = 0; s < K; ++s) r = fma_(v[s], xg[s].e[e], r);
                    out.e[e] = r;
                    sums_(acc2 + e * C, xr.e[e], r);
                }
                store_cs(vd + ci, out);
            }
            next_stage();
        }
#pragma unroll
        for (int q = 0; q < NACC; ++q) park2[q] = acc2[q];
        }
    }
    double acc1[NACC], acc2[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) { acc1[q] = park1[q]; acc2[q] = park2[q]; }

    StepDev fin{};
    fin.R = a.R; fin.cpr = a.cpr; fin.rpb = a.rpb;
    fin.partials = a.partials; fin.counter = a.counter; fin.mom = a.mom; fin.m01 = a.m01; fin.M = a.M; fin.n = a.n; fin.fin = FIN_STEP;
    finish_sums<C, NACC, TPB>(fin, acc1, static_cast<int>(tx), static_cast<int>(ty));
    __syncthreads();
    fin.partials = a.partials + static_cast<int64_t>(gridDim.x) * a.R * C;
    fin.counter = a.counter + 1;
    fin.n = a.n + 1;
    finish_sums<C, NACC, TPB>(fin, acc2, static_cast<int>(tx), static_cast<int>(ty));
}

using PairKernel = void (*)(PairDev);
constexpr int PAIR_MAX_DYN = 220 * 1024;
constexpr int PAIR_TPB = 256;

cudaError_t resident_pair_blocks(PairKernel fn, int block, int dyn_smem, int* out) {
    static std::mutex mutex;
    static std::map<std::tuple<PairKernel, int, int>, int> cache;
    static std::map<PairKernel, bool> raised;
    std::lock_guard<std::mutex> lock(mutex);
    if (!raised[fn]) {
        cudaError_t const err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_MAX_DYN);
        if (err != cudaSuccess) return err;
        raised[fn] = true;
    }
    auto const key = std::make_tuple(fn, block, dyn_smem);
    auto const it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return cudaSuccess; }
    int nb = 0;
    cudaError_t const err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, block, dyn_smem);
    if (err != cudaSuccess) return err;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    cache[key] = nb;
    *out = nb;
    return cudaSuccess;
}

template<class T, int V>
PairKernel pair_kernel_k(int k, int minb) {
    switch (k) {
        case 3: return minb >= 4 ? cheb_pair_bulk<T, V, 3, PAIR_TPB, 4> : minb >= 3 ? cheb_pair_bulk<T, V, 3, PAIR_TPB, 3> : cheb_pair_bulk<T, V, 3, PAIR_TPB, 2>;
        default: return nullptr;
    }
}

template<class T>
cudaError_t launch_pair_t(PairArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled) {
    constexpr int V = 16 / sizeof(T);
    *handled = false;
    if (a.R % V != 0) return cudaSuccess;
    int const cpr = a.R / V;
    if (cpr > PAIR_TPB) return cudaSuccess;
    int const rpb = PAIR_TPB / cpr;
    uint32_t rec = 0, valoff = 0;
    packed_record_layout(sizeof(T), a.k, &rec, &valoff);
    if (rec % 16u != 0) return cudaSuccess;
    if (a.nrows < 4 * a.tile || (a.nrows + a.tile) * cpr >= (int64_t{1} << 32) || a.nrows >= (int64_t{1} << 31) - (int64_t{1} << 24)) return cudaSuccess;
    int const stages = a.stages > 16 ? 16 : (a.stages < 2 ? 2 : a.stages);
    uint32_t const stage_bytes = 2u * PAIR_TPB * 16u + (static_cast<uint32_t>(rpb) * rec + 127u) / 128u * 128u;
    uint32_t const halo_off = (stages * stage_bytes + 16u * stages + 8u /* halo prefetch barrier */ + 127u) / 128u * 128u;
    int64_t const halo_bytes = static_cast<int64_t>(a.halo_max) * cpr * 16;
    int64_t const hrec_off = halo_off + halo_bytes;                       // staged H records of the halo rows (16-byte units)
    int64_t const dyn = hrec_off + static_cast<int64_t>(a.halo_max) * rec;
    if (dyn > PAIR_MAX_DYN) return cudaSuccess;
    // 3 resident CTAs (<= 80 registers, some spills) when their shared memory fits, else 2 (<= 128 registers)
    int minb = a.min_blocks;
    if (minb <= 0) minb = 3 * (dyn + 1024) <= 227 * 1024 ? 3 : 2;
    PairKernel const fn = pair_kernel_k<T, V>(a.k, minb);
    if (!fn) return cudaSuccess;
    int64_t const ntiles = (a.nrows + a.tile - 1) / a.tile;
    int resident = 0;
    cudaError_t const occ = resident_pair_blocks(fn, PAIR_TPB, static_cast<int>(dyn), &resident);
    if (occ != cudaSuccess) return occ;
    int const cap = num_sms * (a.blocks_per_sm > 0 && a.blocks_per_sm < resident ? a.blocks_per_sm : resident);
    int grid = static_cast<int>(ntiles < static_cast<int64_t>(cap) ? ntiles : cap);
    if (grid > max_step_blocks(num_sms)) grid = max_step_blocks(num_sms);

    PairDev d{};
    d.packed = static_cast<const unsigned char*>(a.packed);
    d.packed2 = static_cast<const unsigned char*>(a.packed2);
    d.halo_ptr = a.halo_ptr; d.halo_rows = a.halo_rows;
    d.a = a.a; d.b = a.b; d.c = a.c; d.d = a.d;
    d.nrows = static_cast<int>(a.nrows); d.ntiles = static_cast<int>(ntiles); d.cpr = cpr; d.rpb = rpb; d.tile = static_cast<int>(a.tile);
    d.R = a.R; d.stages = stages; d.rec = rec; d.valoff = valoff; d.stage_bytes = stage_bytes; d.halo_off = halo_off; d.hrec_off = static_cast<uint32_t>(hrec_off);
    d.partials = a.partials; d.counter = a.counter; d.mom = a.mom; d.m01 = a.m01; d.M = a.M; d.n = a.n;
    fn<<<grid, PAIR_TPB, static_cast<size_t>(dyn), stream>>>(d);
    *handled = true;
    if (info) { info->grid = grid; info->block = PAIR_TPB; info->V = V; info->K = a.k; info->bulk = stages; }
    return cudaGetLastError();
}

} // anonymous namespace

cudaError_t launch_step_pair(int dtype, PairArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled) {
    *handled = false;
    if (!a.packed || !a.packed2 || !a.halo_ptr || !a.halo_rows || a.tile <= 0 || a.nrows <= 0) return cudaSuccess;
    switch (dtype) {
        case C64: return launch_pair_t<float2>(a, num_sms, stream, info, handled);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
