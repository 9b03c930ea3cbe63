// gemm_loop_probe.cu -- where does the FP64 DMMA GEMM main loop (csrc/kubo.cu) lose time?  The 32 x 32 warp tile of
// kubo_gemm_kernel (4 A fragments x 4 B fragments = 16 DMMA.8x8x4 per k4-step) is run in four settings, 16 warps per CTA,
// one CTA per SM, no global memory traffic at all:
//   regs      operands stay in registers (rotated so that no two consecutive DMMAs see the same pair)
//   lds       fragments are read from a shared-memory stage with the GEMM's padded layout, no barriers
//   lds+bar   ... plus a CTA barrier every `kper` k4-steps (the cp.async ring's stage hand-over)
//   lds+mbar  ... the hand-over is a per-warp mbarrier arrive / wait on a ring of stages instead (warps may drift apart)
// Build: make -C tools/probe ; run on the GPU box.  Prints one JSON line per setting.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

constexpr int STRIDE = 160;              // bytes between rows of a stage: 16 doubles + 32 bytes of padding (ROW_BYTES = 128)
constexpr int TILE = 128 * STRIDE;       // one operand
constexpr int STAGE = 2 * TILE;
constexpr int NST = 4;

// MODE 0: regs, 1: lds, 2: lds + __syncthreads, 3: lds + mbarrier ring
template<int MODE, int WM, int WNB, int THREADS>      // warp tile = 8 WM x 8 WNB
__global__ void __launch_bounds__(THREADS, 1) loop_probe(double* out, int steps, int kper) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t const s0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const nwarps = blockDim.x >> 5;
    int const wn_count = 128 / (8 * WNB);
    int const wm = warp / wn_count, wn = warp % wn_count;
    for (int i = tid; i < NST * STAGE / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1e-3 * (i % 97);
    uint32_t const bars = s0 + NST * STAGE;
    if (tid == 0) for (int s = 0; s < NST; ++s) mbar_init(bars + 8 * s, nwarps);
    __syncthreads();
    int const fr = lane >> 2, fk = lane & 3;
    uint32_t const offA = (wm * 8 * WM + fr) * STRIDE + fk * 8;
    uint32_t const offB = TILE + (wn * 8 * WNB + fr) * STRIDE + fk * 8;
    double c[WM][WNB][2];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WNB; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
    double a[WM], b[WNB];
#pragma unroll
    for (int i = 0; i < WM; ++i) a[i] = 1e-3 * (lane + i);
#pragma unroll
    for (int j = 0; j < WNB; ++j) b[j] = 1e-3 * (lane - j);
    int slot = 0, h = 0;
    for (int st = 0; st < steps; st += kper, ++h) {
        uint32_t const base = s0 + slot * STAGE;
        if (MODE == 2) __syncthreads();
        if (MODE == 3 && h >= NST - 1) {
            // wait until every warp has finished hand-over h - (NST - 1): a warp may run at most NST - 1 stages ahead of the slowest
            int const hp = h - (NST - 1);
            mbar_wait(bars + 8 * (hp % NST), static_cast<uint32_t>(hp / NST) & 1u);
        }
#pragma unroll 4
        for (int kk = 0; kk < kper; ++kk) {
            if (MODE != 0) {
#pragma unroll
                for (int i = 0; i < WM; ++i) a[i] = lds64(base + offA + i * 8 * STRIDE + (kk & 3) * 32);
#pragma unroll
                for (int j = 0; j < WNB; ++j) b[j] = lds64(base + offB + j * 8 * STRIDE + (kk & 3) * 32);
            }
#pragma unroll
            for (int i = 0; i < WM; ++i)
#pragma unroll
                for (int j = 0; j < WNB; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
        if (MODE == 3) { __syncwarp(); if (lane == 0) mbar_arrive(bars + 8 * slot); }
        if (++slot == NST) slot = 0;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WNB; ++j) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + tid] = s;
}

template<int MODE, int WM, int WNB, int threads> void run(const char* name, int kper) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int const smem = NST * STAGE + 64;
    auto k = loop_probe<MODE, WM, WNB, threads>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    double* out; cudaMalloc(&out, sizeof(double) * sms * threads);
    int const steps = 40000 / kper * kper;
    k<<<sms, threads, smem>>>(out, 400 / kper * kper + kper, kper);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms, threads, smem>>>(out, steps, kper);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double const flops = double(sms) * (threads / 32) * steps * WM * WNB * 512.0;
    printf("{\"setting\": \"%s\", \"warps\": %d, \"warp_tile\": \"%dx%d\", \"k4_steps_per_handover\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n",
           name, threads / 32, 8 * WM, 8 * WNB, kper, ms, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    // 16 warps, 32 x 32 warp tiles (the shipped layout)
    run<0, 4, 4, 512>("regs", 4);
    run<1, 4, 4, 512>("lds", 4);
    run<2, 4, 4, 512>("lds+bar", 4);
    run<2, 4, 4, 512>("lds+bar", 8);
    run<3, 4, 4, 512>("lds+mbar", 4);
    run<3, 4, 4, 512>("lds+mbar", 8);
    // 8 warps, 32 x 64 warp tiles
    run<0, 4, 8, 256>("regs", 4);
    run<1, 4, 8, 256>("lds", 4);
    run<2, 4, 8, 256>("lds+bar", 8);
    // 8 warps, 64 x 32
    run<1, 8, 4, 256>("lds", 4);
    run<3, 8, 4, 256>("lds+mbar", 8);
    run<3, 4, 8, 256>("lds+mbar", 8);
    return 0;
}
