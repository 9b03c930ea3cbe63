// kernels_bulk.cu -- K1, main variant: the fused Chebyshev step fed entirely by the bulk-copy (TMA) engine.
//
//     y[row, :] = sum_s val[row][s] * x[col[row][s], :]  -  y[row, :]
//     m2[r] += |x[row, r]|^2 ,  m3[r] += conj(y_new[row, r]) * x[row, r]          (f64 accumulators)
//
// Same arithmetic, reductions and moment bookkeeping as `cheb_step` (kernels.cu); what differs is how the
// operands reach the SM.  Every *streamed* operand of a block-iteration is contiguous in memory:
//   * y[rows]            -- rpb rows x R lanes, read once and overwritten in place,
//   * x[rows]            -- the CTA's own rows of the gathered vector (needed for the two dot products),
//   * H records of rows  -- (col[K], val[K]) of each row, from the *packed* copy of the ELL matrix: granules of four
//                           consecutive rows, int32 col[4][K] then T val[4][K] = 4 K (4 + sizeof T) bytes, always a
//                           multiple of 16 and exactly the algorithmic size (no padding: 36 bytes per row for c64, K = 3),
// so thread 0 moves them with three `cp.async.bulk` copies per iteration into a `stages`-deep ring of shared-memory
// buffers (mbarrier complete_tx signalling; SASS: UBLKCP + SYNCS).  Up to stages x ~8.5 KB per CTA are in flight
// towards HBM without holding a single register, the warps read y / x / H with conflict-free ld.shared, and the
// only global loads they issue are the gathers x[col] -- L1 / L2 hits inside the locality cluster -- plus one
// coalesced 16-byte store.  All indices are 32-bit (checked by the launcher).  Compared with the register-
// prefetching 64-bit general kernel this halves the instructions per row and removes the dependence of the DRAM
// queue depth on occupancy and SM clock.
//
// Replaces, from the reference (cppcore/): compute::kpm_spmv_diagonal (include/compute/kernel_polynomial.hpp:288-323)
// with the batching of DefaultCompute (src/kpm/default/Compute.cpp:52-88) and the Diagonal collectors.
#include "bulk_common.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace pbk {

namespace {

struct BulkDev {  // kernel parameters
    const unsigned char* packed;  // granules of 4 rows: int32 col[4][K], then T val[4][K]; `gran` bytes each
    const void* x; void* y;
    int nrows, cpr, rpb, ipt, tile_jump;  // tile_jump: rows to skip to reach this CTA's next tile
    int R, stages, k;
    int release;                          // when a warp hands a stage back: 0 right after reading it, 1 after the row is finished
    uint32_t gran, valoff, stage_bytes;   // valoff = 16 K: offset of the values inside a granule
    double* partials; unsigned* counter; double* mom; double* m01; int M; int n; int fin;
};

// K > 0: ELL width known at compile time (all K gathers of a row in flight together); K == 0: any width, the slots are
// walked four at a time (same order of the FMAs, so the vectors are bit-identical to the specialised kernels).
template<class T, int V, int K, bool XS, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) cheb_step_bulk(BulkDev a) {
    using CH = Chunk<T, V>;
    static_assert(sizeof(CH) == 16, "the staged kernel moves 16-byte chunks");
    constexpr int C = ST<T>::C;
    constexpr int NACC = V * C;
    constexpr uint32_t HALF = TPB * 16u;             // bytes of one staged vector operand
    constexpr uint32_t HOFF = XS ? 2u * HALF : HALF;  // H records follow the vector operand(s) inside a stage
    extern __shared__ __align__(128) unsigned char dyn_smem[];

    const CH* __restrict__ x = static_cast<const CH*>(a.x);
    CH* __restrict__ y = static_cast<CH*>(a.y);

    uint32_t const tid = threadIdx.x;
    uint32_t const cpr = a.cpr, rpb = a.rpb;
    int const ipt = a.ipt;
    uint32_t const tx = tid % cpr, ty = tid / cpr;
    bool const active = ty < rpb;
    uint32_t const S = a.stages;
    uint32_t const stage_bytes = a.stage_bytes;
    uint32_t const smem0 = smem_u32(dyn_smem);
    uint32_t const ring_end = smem0 + S * stage_bytes;
    uint32_t const full0 = ring_end;          // full[S] then empty[S], 8 bytes each
    uint32_t const empty_off = 8u * S;        // empty[st] = full[st] + empty_off
    // all positions are in 16-byte chunk units: chunk index of (row, tx) = row * cpr + tx  (< 2^32, checked by the launcher)
    uint32_t const lim = static_cast<uint32_t>(a.nrows) * cpr;  // first chunk past the last row
    uint32_t const step_c = rpb * cpr;                          // chunks per block-iteration
    uint32_t const jump_c = static_cast<uint32_t>(a.tile_jump) * cpr;  // extra chunks to this CTA's next tile

    if (tid == 0) {
        for (uint32_t st = 0; st < S; ++st) { mbar_init(full0 + 8u * st, 1u); mbar_init(full0 + empty_off + 8u * st, TPB / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    uint32_t const first_c = blockIdx.x * static_cast<uint32_t>(ipt) * step_c;  // the CTA's first tile

    // ---- producer (thread 0): runs S - 1 block-iterations ahead of the consumers ----
    uint32_t pc0 = first_c, psb = smem0, pfb = full0, pround = 0;
    int pw = 0;
    auto produce = [&]() {
        if (pround > 0) mbar_wait(pfb + empty_off, (pround - 1u) & 1u);  // every warp has read the stage's previous content
        uint32_t const left = lim - pc0;
        uint32_t const vbytes = (left < step_c ? left : step_c) * 16u;
        uint32_t const hbytes = (rpb >> 2) * a.gran;  // the packed copy is padded: always whole blocks of granules
        mbar_expect_tx(pfb, (XS ? 2u * vbytes : vbytes) + hbytes);
        bulk_g2s(psb, y + pc0, vbytes, pfb);
        if constexpr (XS) bulk_g2s(psb + HALF, x + pc0, vbytes, pfb);
        else bulk_prefetch_l2(x + pc0, vbytes);
        bulk_g2s(psb + HOFF, a.packed + static_cast<size_t>((pc0 / cpr) >> 2) * a.gran, hbytes, pfb);
        pc0 += step_c;
        if (++pw == ipt) { pw = 0; pc0 += jump_c; }
        psb += stage_bytes; pfb += 8u;
        if (psb == ring_end) { psb = smem0; pfb = full0; ++pround; }
    };
    if (tid == 0) {
        for (uint32_t i = 0; i + 1 < S && pc0 < lim; ++i) produce();
    }

    double acc[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = 0.0;

    // ---- consumers ----
    uint32_t c0 = first_c, sb = smem0, fb = full0, cph = 0;
    int w = 0;
    uint32_t const my_vec = tid * 16u;
    uint32_t const kk = K > 0 ? static_cast<uint32_t>(K) : static_cast<uint32_t>(a.k);
    uint32_t const my_gran = HOFF + (ty >> 2) * a.gran;
    uint32_t const my_rec = my_gran + (ty & 3u) * kk * 4u;                                        // col[ty & 3][0]
    uint32_t const my_val = my_gran + a.valoff + (ty & 3u) * kk * static_cast<uint32_t>(sizeof(T));  // val[ty & 3][0]

    while (c0 < lim) {
        if (tid == 0 && pc0 < lim) produce();  // refill the stage consumed one iteration ago

        uint32_t const ci = c0 + tid;          // my chunk of this block-iteration: row = ci / cpr
        bool const valid = active && ci < lim;
        CH yv, xr;
        if constexpr (!XS) { if (valid) xr = load_nc(x + ci); }

        mbar_wait(fb, cph);
        if constexpr (K > 0) {
            int32_t c[K]; T v[K];
            if (valid) {
                yv = lds_chunk<CH>(sb + my_vec);
                if constexpr (XS) xr = lds_chunk<CH>(sb + HALF + my_vec);
#pragma unroll
                for (int s = 0; s < K; ++s) { c[s] = lds_i32(sb + my_rec + 4u * s); lds_val(sb + my_val + static_cast<uint32_t>(sizeof(T)) * s, v[s]); }
            }
            if (a.release == 0) release_stage(fb + empty_off, tid);   // this warp is done with the stage

            if (valid) {
                CH xg[K];
#pragma unroll
                for (int s = 0; s < K; ++s) xg[s] = load_nc(x + (static_cast<uint32_t>(c[s]) * cpr + tx));
                CH out;
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
            if (a.release != 0) release_stage(fb + empty_off, tid);   // experiment: hand the stage back only after the row is finished
        } else {
            CH out;
            if (valid) {
                yv = lds_chunk<CH>(sb + my_vec);
                if constexpr (XS) xr = lds_chunk<CH>(sb + HALF + my_vec);
#pragma unroll
                for (int e = 0; e < V; ++e) out.e[e] = neg_(yv.e[e]);
                for (uint32_t s0 = 0; s0 < kk; s0 += 4u) {
                    int32_t c[4]; T v[4]; CH xg[4];
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) {
                        if (s0 + j < kk) { c[j] = lds_i32(sb + my_rec + 4u * (s0 + j)); lds_val(sb + my_val + static_cast<uint32_t>(sizeof(T)) * (s0 + j), v[j]); }
                    }
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) { if (s0 + j < kk) xg[j] = load_nc(x + (static_cast<uint32_t>(c[j]) * cpr + tx)); }
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) {
                        if (s0 + j < kk) {
#pragma unroll
                            for (int e = 0; e < V; ++e) out.e[e] = fma_(v[j], xg[j].e[e], out.e[e]);
                        }
                    }
                }
            }
            release_stage(fb + empty_off, tid);  // the records were read slot by slot: release the stage now
            if (valid) {
#pragma unroll
                for (int e = 0; e < V; ++e) sums_(acc + e * C, xr.e[e], out.e[e]);
                store_cs(y + ci, out);
            }
        }
        c0 += step_c;
        if (++w == ipt) { w = 0; c0 += jump_c; }
        sb += stage_bytes; fb += 8u;
        if (sb == ring_end) { sb = smem0; fb = full0; cph ^= 1u; }
    }

    StepDev fin{};
    fin.R = a.R; fin.cpr = a.cpr; fin.rpb = a.rpb;
    fin.partials = a.partials; fin.counter = a.counter; fin.mom = a.mom; fin.m01 = a.m01; fin.M = a.M; fin.n = a.n; fin.fin = a.fin;
    finish_sums<C, NACC, TPB>(fin, acc, static_cast<int>(tx), static_cast<int>(ty));
}

// ---- packing: slot-major ELL -> granules of four rows --------------------------------------------------
template<class T>
__global__ void pack_ell_kernel(const T* __restrict__ val, const int32_t* __restrict__ col, int64_t pitch, int k, int64_t rows,
                                int64_t padded_rows, unsigned char* __restrict__ out, uint32_t gran, uint32_t valoff) {
    int64_t const row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= padded_rows) return;
    unsigned char* const g = out + (row >> 2) * gran;
    int32_t* const cp = reinterpret_cast<int32_t*>(g) + (row & 3) * k;
    T* const vp = reinterpret_cast<T*>(g + valoff) + (row & 3) * k;
    for (int s = 0; s < k; ++s) {
        bool const in = row < rows;
        cp[s] = in ? col[s * pitch + row] : 0;
        vp[s] = in ? val[s * pitch + row] : zero_(T{});
    }
}

using BulkKernel = void (*)(BulkDev);

constexpr int BULK_MAX_DYN = 200 * 1024;

/// resident CTAs per SM for (kernel, dynamic shared memory); the opt-in shared-memory limit of a kernel is raised
/// once, to the largest size the launcher ever asks for
cudaError_t resident_bulk_blocks(BulkKernel fn, int block, int dyn_smem, int* out) {
    static std::mutex mutex;
    static std::map<std::tuple<BulkKernel, int, int>, int> cache;
    static std::map<BulkKernel, bool> raised;
    std::lock_guard<std::mutex> lock(mutex);
    if (!raised[fn]) {
        cudaError_t const err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, BULK_MAX_DYN);
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

constexpr int BULK_TPB = 256;

template<class T, int V, bool XS>
BulkKernel bulk_kernel_k(int k) {
    switch (k) {
        case 3: return cheb_step_bulk<T, V, 3, XS, BULK_TPB, 4>;
        case 4: return cheb_step_bulk<T, V, 4, XS, BULK_TPB, 4>;
        case 7: return cheb_step_bulk<T, V, 7, XS, BULK_TPB, 4>;
        default: return cheb_step_bulk<T, V, 0, XS, BULK_TPB, 4>;   // any other width (next-nearest neighbours, several orbitals ...)
    }
}

template<class T>
cudaError_t launch_bulk_t(StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled) {
    constexpr int V = 16 / sizeof(T);
    *handled = false;
    if (a.R % V != 0) return cudaSuccess;
    int const cpr = a.R / V;
    if (cpr > BULK_TPB) return cudaSuccess;
    int const rpb = (BULK_TPB / cpr) & ~3;   // whole granules of four rows per block-iteration (threads beyond rpb * cpr idle)
    uint32_t gran = 0, valoff = 0;
    packed_record_layout(sizeof(T), a.h.k, &gran, &valoff);
    if (rpb < 4 || rpb > PACKED_PAD_ROWS) return cudaSuccess;   // a block-iteration copies whole granules of four rows
    int ipt = 1;
    if (a.tile > rpb) ipt = static_cast<int>((a.tile + rpb - 1) / rpb);
    int64_t const tile_rows = static_cast<int64_t>(ipt) * rpb;
    if (a.nrows < 4 * tile_rows || a.nrows >= (int64_t{1} << 31) - (int64_t{1} << 24) ||
        (a.nrows + tile_rows) * cpr >= (int64_t{1} << 32)) return cudaSuccess;
    BulkKernel const fn = a.bulk_xstage ? bulk_kernel_k<T, V, true>(a.h.k) : bulk_kernel_k<T, V, false>(a.h.k);
    if (!fn) return cudaSuccess;

    int const stages = a.bulk_stages > 16 ? 16 : a.bulk_stages;
    uint32_t const hoff = BULK_TPB * 16u * (a.bulk_xstage ? 2u : 1u);
    uint32_t const stage_bytes = hoff + (static_cast<uint32_t>(rpb / 4) * gran + 127u) / 128u * 128u;
    int const dyn = static_cast<int>(stages * stage_bytes + 16u * stages);
    if (dyn > BULK_MAX_DYN) return cudaSuccess;
    int64_t const need = (a.nrows + tile_rows - 1) / tile_rows;
    int resident = 0;
    cudaError_t const occ = resident_bulk_blocks(fn, BULK_TPB, dyn, &resident);
    if (occ != cudaSuccess) return occ;
    int const cap = num_sms * (a.blocks_per_sm > 0 && a.blocks_per_sm < resident ? a.blocks_per_sm : resident);
    int grid = static_cast<int>(need < static_cast<int64_t>(cap) ? need : cap);
    if (grid > max_step_blocks(num_sms)) grid = max_step_blocks(num_sms);
    // the chunk cursor may run one grid-stride past the end before the loop stops: keep it inside uint32
    if ((a.nrows + (static_cast<int64_t>(grid) + 1) * tile_rows) * cpr >= (int64_t{1} << 32)) return cudaSuccess;

    BulkDev d{};
    d.packed = static_cast<const unsigned char*>(a.packed);
    d.x = a.x; d.y = a.y;
    d.nrows = static_cast<int>(a.nrows); d.cpr = cpr; d.rpb = rpb; d.ipt = ipt;
    d.tile_jump = static_cast<int>((grid - 1) * tile_rows);
    d.R = a.R; d.stages = stages; d.k = a.h.k; d.release = a.bulk_release; d.gran = gran; d.valoff = valoff; d.stage_bytes = stage_bytes;
    d.partials = a.partials; d.counter = a.counter; d.mom = a.mom; d.m01 = a.m01; d.M = a.M; d.n = a.n; d.fin = a.fin;
    fn<<<grid, BULK_TPB, dyn, stream>>>(d);
    *handled = true;
    if (info) { info->grid = grid; info->block = BULK_TPB; info->V = V; info->K = a.h.k; info->bulk = stages; }
    return cudaGetLastError();
}

} // anonymous namespace

void packed_record_layout(int scalar_bytes, int k, uint32_t* gran, uint32_t* valoff) {
    *valoff = 16u * static_cast<uint32_t>(k);                                       // int32 col[4][k]
    *gran = 4u * static_cast<uint32_t>(k) * (4u + static_cast<uint32_t>(scalar_bytes));   // + T val[4][k]: a multiple of 16 for every k
}

size_t packed_ell_bytes(int dtype, EllDev const& h) {
    uint32_t gran = 0, valoff = 0;
    packed_record_layout(dtype_size(dtype), h.k, &gran, &valoff);
    return static_cast<size_t>((h.rows + PACKED_PAD_ROWS + 3) / 4) * gran;
}

cudaError_t launch_pack_ell(int dtype, EllDev const& h, void* packed, cudaStream_t s) {
    uint32_t rec = 0, valoff = 0;
    packed_record_layout(dtype_size(dtype), h.k, &rec, &valoff);
    int64_t const padded = (h.rows + PACKED_PAD_ROWS + 3) / 4 * 4;
    int const grid = static_cast<int>((padded + 255) / 256);
    auto* out = static_cast<unsigned char*>(packed);
    switch (dtype) {
        case F32: pack_ell_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(h.val), h.col, h.pitch, h.k, h.rows, padded, out, rec, valoff); break;
        case C64: pack_ell_kernel<float2><<<grid, 256, 0, s>>>(static_cast<const float2*>(h.val), h.col, h.pitch, h.k, h.rows, padded, out, rec, valoff); break;
        case F64: pack_ell_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(h.val), h.col, h.pitch, h.k, h.rows, padded, out, rec, valoff); break;
        case C128: pack_ell_kernel<double2><<<grid, 256, 0, s>>>(static_cast<const double2*>(h.val), h.col, h.pitch, h.k, h.rows, padded, out, rec, valoff); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_step_bulk(int dtype, StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled) {
    *handled = false;
    if (a.bulk_stages < 2 || !a.packed || !a.subtract || !a.sums || a.y2 || a.tile <= 0 || a.nrows <= 0) return cudaSuccess;
    switch (dtype) {
        case F32: return launch_bulk_t<float>(a, num_sms, stream, info, handled);
        case C64: return launch_bulk_t<float2>(a, num_sms, stream, info, handled);
        case F64: return launch_bulk_t<double>(a, num_sms, stream, info, handled);
        case C128: return launch_bulk_t<double2>(a, num_sms, stream, info, handled);
        default: return cudaErrorInvalidValue;
    }
}

} // namespace pbk
