// dmma_probe.cu -- FP64 throughput of the B200 SM by instruction shape (register-resident operands, no memory):
// DFMA vs mma.sync f64 m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16.  Build: make -C tools/probe ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template<int SHAPE> __device__ __forceinline__ void op(double (&c)[4], double (&a)[8], double (&b)[4]) {
    if constexpr (SHAPE == 0) {  // 4 DFMA
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = fma(a[i], b[i], c[i]);
    } else if constexpr (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
    } else if constexpr (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
    } else if constexpr (SHAPE == 3) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
}

template<int SHAPE, int CHAINS>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
    double a[8], b[4], c[CHAINS][4];
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = seed * 0.5 + i;
    for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) c[j][i] = j + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) op<SHAPE>(c[j], a, b);
    }
    double s = 0;
    for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int SHAPE, int CHAINS> void run(const char* name, double flops_per_warp_op, int blocks_per_sm) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int const grid = sms * blocks_per_sm, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * grid * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<SHAPE, CHAINS><<<grid, 256>>>(out, 100, 1.0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<SHAPE, CHAINS><<<grid, 256>>>(out, iters, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double const warps = double(grid) * 8;
    double const tf = warps * iters * CHAINS * flops_per_warp_op / (ms * 1e-3) / 1e12;
    printf("{\"op\": \"%s\", \"chains\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", name, CHAINS, blocks_per_sm, ms, tf,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int bps : {1, 2, 4}) {
        run<0, 8>("dfma x4/thread", 4.0 * 32 * 2, bps);
        run<1, 8>("dmma m8n8k4", 2.0 * 8 * 8 * 4, bps);
        run<2, 8>("dmma m16n8k4", 2.0 * 16 * 8 * 4, bps);
        run<3, 8>("dmma m16n8k8", 2.0 * 16 * 8 * 8, bps);
        run<4, 8>("dmma m16n8k16", 2.0 * 16 * 8 * 16, bps);
    }
    run<4, 4>("dmma m16n8k16", 2.0 * 16 * 8 * 16, 2);
    run<4, 2>("dmma m16n8k16", 2.0 * 16 * 8 * 16, 4);
    run<1, 16>("dmma m8n8k4", 2.0 * 8 * 8 * 4, 2);
    return 0;
}
