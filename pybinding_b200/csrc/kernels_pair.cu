// kernels_pair.cu -- K1b `cheb_pair_bulk`: TWO fused Chebyshev steps per pass over the vectors (temporal blocking).
//
//     c = H~ b - a        (r_{n+1} from r_n = b and r_{n-1} = a)
//     d = H~ c - b        (r_{n+2})
//     sums of step n+1:  |b|^2, conj(c) b        sums of step n+2:  |c|^2, conj(d) c       (f64, four moments per launch)
//
// One step moves 3 vector streams per row (read x, read y, write y); two separate steps move 6.  Here a CTA owns
// one locality cluster (tile) at a time and, per tile,
//   phase 1   computes c on the tile's own rows (written to global memory, which the tile re-reads from L1/L2) *and*
//             on the tile's one-ring halo (rows of neighbouring tiles that the tile's rows reference; kept in shared
//             memory only -- the owner tile computes and stores them itself, bit-identically),
//   phase 2   computes d on the own rows from c(own: global, halo: shared memory),
// so DRAM sees read a, read b, write c, write d = 4 streams (+ the halo re-reads) for two steps.  The redundant work
// is the halo fraction of phase 1 (~28 % of the rows for 256-site clusters of a honeycomb lattice).
//
// Staging is the same as `cheb_step_bulk` (kernels_bulk.cu): thread 0 feeds a ring of shared-memory stages with
// cp.async.bulk copies of the contiguous operands -- phase 1: a[rows], b[rows], H records; phase 2: b[rows] and the
// phase-2 H records, whose column entries are *codes*: >= 0 a global row of c (own tile), < 0 a slot of the halo
// buffer.  Halo rows are scattered, so their a-row, H record and gathers are ordinary loads.
// Inputs (a, b, records) are never written by the launch; outputs (c, d) are separate buffers, so tiles are
// independent and the result does not depend on the tile -> CTA schedule.
//
// Replaces, from the reference: two consecutive iterations of calc_moments::basic (diagonal)
// (cppcore/include/kpm/calc_moments.hpp:36-51) = 2 x compute::kpm_spmv_diagonal (kernel_polynomial.hpp:288-323).
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
    int ptile = blockIdx.x, pphase = 0;
    uint32_t prow = 0, prows = ptile < ntiles ? rows_of(ptile) : 0;   // next row of the tile to stage, rows of the tile
    uint32_t psb = smem0, pfb = full0, pround = 0;
    auto produce = [&]() {
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
    };
    if (tid == 0) {
        for (uint32_t i = 0; i + 1 < S && ptile < ntiles; ++i) produce();
    }

    double acc1[NACC], acc2[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) { acc1[q] = 0.0; acc2[q] = 0.0; }

    uint32_t sb = smem0, fb = full0, cph = 0;
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

        // ---- phase 1, own rows: c = H b - a, sums |b|^2 and conj(c) b ----
        for (uint32_t w0 = 0; w0 < trows; w0 += rpb) {
            if (tid == 0 && ptile < ntiles) produce();
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

        __syncthreads();   // every thread is done with the previous tile's halo buffer

        // ---- phase 1, halo rows: c into shared memory only ----
        {
            int const hp = __ldg(a.halo_ptr + tile);
            int const nh = __ldg(a.halo_ptr + tile + 1) - hp;
            int j = static_cast<int>(ty);
            int32_t cn[K]; T vn[K]; uint32_t hrow_n = 0;
            auto fetch = [&](int jj) {
                hrow_n = static_cast<uint32_t>(__ldg(a.halo_rows + hp + jj));
                const unsigned char* rp = a.packed + static_cast<size_t>(hrow_n) * a.rec;
#pragma unroll
                for (int s = 0; s < K; ++s) { cn[s] = __ldg(reinterpret_cast<const int32_t*>(rp) + s); vn[s] = ldg_val<T>(rp + a.valoff + sizeof(T) * s); }
            };
            if (active && j < nh) fetch(j);
            while (active && j < nh) {
                int32_t c[K]; T v[K];
#pragma unroll
                for (int s = 0; s < K; ++s) { c[s] = cn[s]; v[s] = vn[s]; }
                uint32_t const hrow = hrow_n;
                CH xg[K];
#pragma unroll
                for (int s = 0; s < K; ++s) xg[s] = load_nc(vb + (static_cast<uint32_t>(c[s]) * cpr + tx));
                CH const yv = load_nc(va + (hrow * cpr + tx));
                int const jn = j + static_cast<int>(rpb);
                if (jn < nh) fetch(jn);
                CH out;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    T r = neg_(yv.e[e]);
#pragma unroll
                    for (int s = 0; s < K; ++s) r = fma_(v[s], xg[s].e[e], r);
                    out.e[e] = r;
                }
                sts_chunk(halo0 + (static_cast<uint32_t>(j) * cpr + tx) * 16u, out);
                j = jn;
            }
        }

        __syncthreads();   // c: own rows visible in global memory (CTA scope), halo rows in shared memory

        // ---- phase 2, own rows: d = H c - b, sums |c|^2 and conj(d) c ----
        for (uint32_t w0 = 0; w0 < trows; w0 += rpb) {
            if (tid == 0 && ptile < ntiles) produce();
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
    }

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
        case 3: return minb >= 3 ? cheb_pair_bulk<T, V, 3, PAIR_TPB, 3> : cheb_pair_bulk<T, V, 3, PAIR_TPB, 2>;
        case 4: return minb >= 3 ? cheb_pair_bulk<T, V, 4, PAIR_TPB, 3> : cheb_pair_bulk<T, V, 4, PAIR_TPB, 2>;
        case 7: return cheb_pair_bulk<T, V, 7, PAIR_TPB, 2>;
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
    uint32_t const halo_off = (stages * stage_bytes + 16u * stages + 127u) / 128u * 128u;
    int64_t const halo_bytes = static_cast<int64_t>(a.halo_max) * cpr * 16;
    int64_t const dyn = halo_off + halo_bytes;
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
    d.R = a.R; d.stages = stages; d.rec = rec; d.valoff = valoff; d.stage_bytes = stage_bytes; d.halo_off = halo_off;
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
        case F32: return launch_pair_t<float>(a, num_sms, stream, info, handled);
        case C64: return launch_pair_t<float2>(a, num_sms, stream, info, handled);
        case F64: return launch_pair_t<double>(a, num_sms, stream, info, handled);
        case C128: return launch_pair_t<double2>(a, num_sms, stream, info, handled);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
