// kernels.cuh -- type-erased launch interface between the host engine and the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pbk {

enum DType : int { F32 = 0, C64 = 1, F64 = 2, C128 = 3 };
inline int dtype_size(int dt) { return dt == F32 ? 4 : (dt == C128 ? 16 : 8); }
inline bool dtype_complex(int dt) { return dt == C64 || dt == C128; }
inline int dtype_words(int dt) { return (dt == F32 || dt == C64) ? 1 : 2; }  // MT19937 words per draw

/// Device ELL matrix, slot-major: element (row, s) at s * pitch + row
struct EllDev {
    void* val = nullptr;     // scalar type of the Hamiltonian
    int32_t* col = nullptr;
    int64_t rows = 0;
    int64_t pitch = 0;
    int k = 0;
};

/// What the last block of a fused step kernel does with the reduced sums
enum FinMode : int { FIN_NONE = 0, FIN_INIT = 1, FIN_STEP = 2 };

struct StepArgs {
    EllDev h;
    const void* x = nullptr;   // N x R row-major block (gathered operand)
    void* y = nullptr;         // N x R block: read (SUB) and written
    void* y2 = nullptr;        // optional second destination, same layout as y
    int64_t y_block_stride = 0;   // != 0: y (y2) is a row of a k-blocked Kubo-Bastin stack -- consecutive 256-byte blocks of the
    int64_t y2_block_stride = 0;  //       vector lie this many bytes apart (general kernel only; a blocked y is write-only)
    int64_t nrows = 0;         // rows [0, nrows) are processed
    int R = 1;                 // vectors advanced together
    bool subtract = true;      // y = H*x - y   (else y = scale * H*x)
    bool sums = false;         // fused sum|x|^2 and sum conj(y_new)*x per vector
    double scale = 1.0;
    // fused moment bookkeeping (sums only)
    double* partials = nullptr;  // [grid][R*C]
    unsigned* counter = nullptr;
    double* mom = nullptr;       // c128 [R][M]
    double* m01 = nullptr;       // [R][3]: m0, re(m1), im(m1)
    int M = 0;
    int n = 0;                   // FIN_STEP: writes moments 2(n-1) and 2(n-1)+1
    int fin = FIN_NONE;
    // launch geometry
    int64_t tile = 0;            // consecutive rows owned by one CTA at a time (0: plain grid-stride)
    int tpb = 256;               // threads per block: 256, 512 or 1024
    int blocks_per_sm = 0;       // 0: as many as are resident
    int prefetch = 0;            // L2 prefetch distance in block-iterations (0: off)
    int prefetch_mask = 0;       // chunks with (index & mask) == 0 issue the prefetch
    // bulk-copy (TMA) staged variant, kernels_bulk.cu
    const void* packed = nullptr;  // row-major packed copy of h (launch_pack_ell); required by the staged variant
    int bulk_stages = 0;         // >= 2: use the staged variant with this pipeline depth where it applies
    bool bulk_xstage = true;     // stage the CTA's own x rows too (else: L2 bulk prefetch + read through L1)
    int bulk_release = 0;        // experiment knob: when a warp hands a stage back to the producer (see kernels_bulk.cu)
};

struct LaunchInfo { int grid = 0, block = 0, V = 0, K = 0, bulk = 0, res = 0; };

/// Fused Chebyshev step (K1/K2/K3).  Returns the launch geometry used (for stats / tests).
cudaError_t launch_step(int dtype, StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info);
/// Staged variant (kernels_bulk.cu): *handled == false means "not applicable, use the general kernel".
cudaError_t launch_step_bulk(int dtype, StepArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled);
/// Packed copy of an ELL matrix for the staged variant: granules of four consecutive rows, int32 col[4][k] followed by
/// T val[4][k] -- `gran` = 4 k (4 + sizeof T) bytes, a multiple of 16 for every k and scalar type, values at `valoff` = 16 k
/// inside the granule; PACKED_PAD_ROWS zero rows after the last row so whole blocks can always be copied.
constexpr int PACKED_PAD_ROWS = 256;
void packed_record_layout(int scalar_bytes, int k, uint32_t* gran, uint32_t* valoff);
size_t packed_ell_bytes(int dtype, EllDev const& h);
cudaError_t launch_pack_ell(int dtype, EllDev const& h, void* packed, cudaStream_t s);
// ---- resident-tile variant (kernels_res.cu): the x rows of a tile and of its halo live in shared memory ----
struct ResTile { int32_t row0, nrows, halo_off, nh; };   // rows [row0, row0 + nrows); halo_rows[halo_off .. + nh)
struct ResGeometry {
    uint32_t row_bytes = 0;        // bytes of one row of a vector block (lanes * sizeof scalar): the variant is built for one width
    uint32_t cb = 0, kvb = 0;      // bytes per row of the code records (uint16, padded to 16) and of the value records (padded to 16)
    uint32_t xs_bytes = 0;         // shared memory of ONE resident-tile buffer
    int cap_rows = 0;              // own + halo rows that fit in it
    int stages = 2, ctas_per_sm = 3, rows_per_iteration = 0;
    int buffers = 1;               // resident-tile buffers per CTA (2: the next tile loads while this one is processed)
};
ResGeometry res_geometry(int dtype, int k, int lanes, int ctas_per_sm, int stages, int buffers);
int res_max_halo();
/// nh_dev[t] = distinct rows outside tile t referenced by its rows (> res_max_halo(): too many to tell)
cudaError_t launch_res_count(EllDev const& h, const ResTile* tiles_dev, int ntiles, int32_t* nh_dev, cudaStream_t s);
/// halo lists (ascending) + per-row 16-bit local codes and values, for tiles whose halo_off / nh are final
cudaError_t launch_res_fill(int dtype, EllDev const& h, const ResTile* tiles_dev, int ntiles, int32_t* halo_rows, void* codes, void* vals,
                            ResGeometry const& g, cudaStream_t s);
struct ResArgs {
    const ResTile* tiles = nullptr; int ntiles = 0;
    const int32_t* halo_rows = nullptr; const void* codes = nullptr; const void* vals = nullptr;
    ResGeometry geo;
    const void* x = nullptr; void* y = nullptr;
    int64_t nrows = 0; int R = 1, k = 0;
    double* partials = nullptr; unsigned* counter = nullptr; double* mom = nullptr; double* m01 = nullptr; int M = 0, n = 0, fin = FIN_NONE;
};
/// y = H x - y with the fused sums (the step of the diagonal recursion); *handled == false: not applicable
cudaError_t launch_step_res(int dtype, ResArgs const& a, int num_sms, cudaStream_t stream, LaunchInfo* info, bool* handled);

// ---- persistent variant (kernels_persist.cu): the whole diagonal recursion of ONE vector of a small system in one launch ----
struct PersistArgs {
    EllDev h;
    int64_t nrows = 0;
    void* buf0 = nullptr; void* buf1 = nullptr;   // buf0 = r0 on entry; both are overwritten
    int steps = 0;                                // M / 2
    double* table = nullptr; int table_ctas = 0;  // [steps][table_ctas][3] doubles of scratch
    unsigned* barrier = nullptr;                  // one word
    double* mom = nullptr; int M = 0;             // c128 [M]
};
/// *handled == false: not applicable (system too large for one resident grid, unsupported ELL width)
cudaError_t launch_persistent_diagonal(int dtype, PersistArgs const& a, int num_sms, cudaStream_t stream, bool* handled);

/// Upper bound of blocks launch_step may use (size of the partials buffer = this * R * 3 doubles)
int max_step_blocks(int num_sms);

// ---- device-side construction of a layout from the resident CSR (build.cu) -----------------------
enum BuildMode : int { BUILD_PLAIN = 0, BUILD_SCALED = 1, BUILD_VELOCITY = 2 };
struct BuildArgs {
    const int32_t* indptr = nullptr; const int32_t* indices = nullptr;   // resident CSR (original order, unscaled)
    int64_t n = 0;
    const int32_t* queue = nullptr;    // row of the layout -> original site, or null (caller's order)
    const int32_t* perm = nullptr;     // original site -> row of the layout (with queue)
    int mode = BUILD_PLAIN;
    double f = 1.0, sb = 0.0;          // BUILD_SCALED: H~ = (H - sb) * f with the reference's association (see build.cu)
    const float* positions = nullptr;  // BUILD_VELOCITY: one coordinate per original site
    int k = 0; int64_t pitch = 0;      // ELL width (launch_row_width) and pitch of the output
    int32_t* col = nullptr;            // output columns, slot-major; the values go to `val` of launch_csr_to_ell
};
int build_max_width();
/// *out_dev = widest row (+1 where a diagonal entry has to be created by a non-zero b offset)
cudaError_t launch_row_width(const int32_t* indptr, const int32_t* indices, int64_t n, bool insert_diag, int* out_dev, cudaStream_t s);
cudaError_t launch_invert_order(const int32_t* queue, int64_t n, int32_t* perm, cudaStream_t s);
cudaError_t launch_csr_to_ell(int dtype, BuildArgs const& a, const void* data, void* val, cudaStream_t s);

// ---- small helpers -------------------------------------------------------------------------------
/// dst[i * R + lane] = (i == src[lane]) for lane < nsrc, 0 elsewhere (UnitStarter); dst is zero-filled first
cudaError_t launch_unit_starter(int dtype, void* dst, int64_t n, int R, const int32_t* src_dev, int nsrc, cudaStream_t s);
/// out[i * stride_out + n] = v[idx[i] * R + lane0]  (MultiUnitCollector); mom is c128
cudaError_t launch_gather_moment(int dtype, const void* v, int R, const int32_t* idx_dev, int nidx, double* mom,
                                 int64_t M, int n, double scale, cudaStream_t s);
/// mom[n] = scale * sum_i conj(beta[i]) * v[i]   (GenericCollector), deterministic two-pass
cudaError_t launch_dot_moment(int dtype, const void* beta, const void* v, int64_t n_rows, double* mom, int n, double scale,
                              double* scratch, unsigned* counter, int num_sms, cudaStream_t s);
/// acc[n] += sum_{lane < R} mom[lane][n]  (BatchAccumulator)
cudaError_t launch_accumulate_lanes(const double* mom, int R, int M, double* acc, cudaStream_t s);
/// dst[perm ? perm[i] : i][lane] = cast(src c128 [i]) -- ConstantStarter upload (src on device, c128)
cudaError_t launch_scatter_block(int dtype, const double* src_c128, int64_t n, int R, int lane, const int32_t* perm_dev,
                                 void* dst, cudaStream_t s);
/// out c128[i] = v[(perm ? perm[i] : i) * R + lane]
cudaError_t launch_extract_lane(int dtype, const void* v, int64_t n, int R, int lane, const int32_t* perm_dev, double* out_c128, cudaStream_t s);

// ---- light-cone sub-systems of LDOS (kernels.cu) ------------------------------------------------
/// gmap[resident row of site queue[i]] = set ? i : -1   (perm_dev: site -> resident row, or null)
cudaError_t launch_cone_mark(const int32_t* queue_dev, int64_t count, const int32_t* perm_dev, int32_t* gmap, bool set, cudaStream_t s);
/// Sub-ELL of the first `rows` sites of `queue` (slot-major, pitch `out_pitch`, columns = positions in `queue`) cut out
/// of the resident scaled ELL `h`; rows [rows, out_pitch) are zero padding.
cudaError_t launch_cone_extract(int dtype, EllDev const& h, const int32_t* queue_dev, const int32_t* perm_dev, const int32_t* gmap,
                                int64_t rows, void* out_val, int32_t* out_col, int64_t out_pitch, cudaStream_t s);

/// One light-cone sub-system of a group (device-resident table read by cone_group_step_kernel)
struct ConeSlot {
    const void* val; const int32_t* col; int64_t pitch;   // sub-ELL (slot-major)
    void* buf[2];                                          // r_even / r_odd
    int64_t nvec;                                          // vector length (sites of the ball)
    const int32_t* rows;                                   // rows[k]: rows processed at step k, k = 1 .. M/2
    double* mom;                                           // c128 [M]
    double* m01;                                           // [3]
};
/// unit starters of a group: buf[0] = e_0, buf[1] = 0
cudaError_t launch_cone_group_start(int dtype, const ConeSlot* slots_dev, int nslots, int64_t max_nvec, cudaStream_t s);
/// step k (1 = initial step) for all sub-systems of a group in one launch; partials: [nslots][blocks][C] doubles,
/// counters: nslots zeroed unsigned
cudaError_t launch_cone_group_step(int dtype, const ConeSlot* slots_dev, int nslots, int k, int kell, int M, int64_t max_rows,
                                   double* partials, unsigned* counters, int blocks_per_slot_cap, cudaStream_t s);

// ---- random starters (mt19937.cu) ------------------------------------------------------------
constexpr int MT_N = 624;
/// state_dev: 624 words + 1 position word.  Seeds std::mt19937's default state (seed 5489, position 624).
cudaError_t launch_mt_seed(uint32_t* state_dev, cudaStream_t s);
/// Write the next `count` tempered outputs of the stream to `out` (out == nullptr: just skip them)
cudaError_t launch_mt_generate(uint32_t* state_dev, uint32_t* out, int64_t count, cudaStream_t s);
/// Segment-parallel variant (mt_jump.cu): tempered outputs [position, position + count) of the default-seeded
/// stream, generated by up to `max_segments` CTAs whose start states come from GF(2) jump-ahead.
int64_t mt_stream_scratch_words(int max_segments);
cudaError_t launch_mt_stream(uint32_t* states_dev, int max_segments, uint64_t position, int64_t count, uint32_t* out,
                             cudaStream_t s, int* launches);
/// host-only reference of the jump: the 624-word window W_{position-1} (words 1.. = raw outputs from `position`)
bool mt_jump_window_host(uint64_t position, uint32_t* window);
/// raw words [lanes_filled][n*w] -> N x R block of +-1 (real) or exp(i*2*pi_f*u) (complex), RandomStarter;
/// lanes >= lanes_filled are zeroed.  perm_dev (optional): destination row of site i (reorder map).
cudaError_t launch_random_transform(int dtype, const uint32_t* raw, int64_t n, int R, int lanes_filled, const int32_t* perm_dev,
                                    void* dst, cudaStream_t s);

// ---- Kubo-Bastin contraction (kubo.cu) -------------------------------------------------------
/// The two M x N stacks are "k-blocked" (kubo.cu): block kb holds bytes [256 kb, 256 kb + 256) of every moment row, rows
/// `row_stride` bytes apart (payload + pad), blocks `block_stride` = M * row_stride bytes apart.  Row m of the stack is
/// written by the step kernel through StepArgs::y_block_stride / y2_block_stride with base = stack + m * row_stride.
/// The bytes of the last block past the end of the vector must be zero.
struct KuboStackLayout { int64_t blocks; int64_t block_stride; uint32_t row_stride; size_t bytes; };
KuboStackLayout kubo_stack_layout(int dtype, int M, size_t vector_bytes);
/// C (M x M, c128 row-major) += A (M x N) * B^H (N x M) for two k-blocked stacks of the Hamiltonian's scalar type
/// (N = sites x lanes).  `workspace` holds the split-K partial tiles (kubo_gemm_workspace_bytes).
size_t kubo_gemm_workspace_bytes(int M, int64_t blocks, int num_sms);
cudaError_t launch_kubo_gemm(int dtype, const void* A, const void* B, int M, int64_t N, KuboStackLayout const& layout, double* C_c128,
                             double* workspace, size_t workspace_bytes, int num_sms, cudaStream_t s, double* flops);

// ---- reconstruction (reconstruct.cu) ----------------------------------------------------------
/// out[c * ne + i] = k / sqrt(1 - E_i^2) * sum_q Re(mu[q * n_stride + c * col_stride]) cos(q acos E_i); E already scaled
cudaError_t launch_spectral_density(const double* mu_c128, int M, int cols, int64_t col_stride, int64_t n_stride, const double* scaled_energy,
                                    int ne, double k, double* out, cudaStream_t s);
/// out[c * ne + i] = -2i / (a sqrt(1 - E_i^2)) * sum_q mu[c * M + q] exp(-i q acos E_i); out is c128
cudaError_t launch_greens(const double* mu_c128, int M, int cols, const double* scaled_energy, int ne, double inv_a, double* out_c128,
                          cudaStream_t s);

// ---- Lanczos helpers (bounds) ----------------------------------------------------------------
/// v0 = t - b_prev*v0 - a*v1 (a read from a_dev[0]) ; out[0] = |v0|^2
cudaError_t launch_lanczos_update(int dtype, int64_t n, const void* t, const void* v1, void* v0, double b_prev, const double* a_dev,
                                  double* out, double* scratch, unsigned* counter, int num_sms, cudaStream_t s);
/// v *= 1/sqrt(norm2[0])
cudaError_t launch_scale_inv_sqrt(int dtype, int64_t n, void* v, const double* norm2_dev, cudaStream_t s);

} // namespace pbk
