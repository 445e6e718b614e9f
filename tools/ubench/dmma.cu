// FP64 tensor-core probe: mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64 versus plain DFMA on B200.
// Prints FMA/clk/SM (one FMA = one multiply-add of doubles) for each, and a DFMA+DMMA mix (same pipe or not).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {   // m16n8k8: a[4], b[2], c[4]
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {  // m16n8k16: a[8], b[4], c[4]
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template <int OP> __global__ void k(double* out, double seed) {
    double c[8][4], a[8], b[4], d[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x + i; d[i] = a[i] * 0.5; for (int j = 0; j < 4; ++j) c[i][j] = 0.0; }
    for (int j = 0; j < 4; ++j) b[j] = seed * (j + 1);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { d[i] = fma(d[i], 1.0000001, 0.25); }
            else if (OP == 1) mma884(c[i][0], c[i][1], a[i], b[0]);
            else if (OP == 2) mma1688(c[i], a, b);
            else if (OP == 3) mma16816(c[i], a, b);
            else if (OP == 4) { mma884(c[i][0], c[i][1], a[i], b[0]); d[i] = fma(d[i], 1.0000001, 0.25); }
        }
    }
    double r = 0; for (int i = 0; i < 8; ++i) { r += d[i]; for (int j = 0; j < 4; ++j) r += c[i][j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> void run(const char* name, double fma_per_thread_per_inner) {
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    k<OP><<<sms * 8, 256>>>(out, 1.5); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(s); k<OP><<<sms * 8, 256>>>(out, 1.5 + r); cudaEventRecord(e); cudaEventSynchronize(e); float ms; cudaEventElapsedTime(&ms, s, e); if (ms < best) best = ms; }
    double fmas = (double)sms * 8 * 256 * ITERS * 8 * fma_per_thread_per_inner;
    printf("%-28s %8.3f ms  %8.1f GFMA/s  %6.1f FMA/clk/SM (at %d MHz nominal)  err=%d\n", name, best, fmas / best / 1e6, fmas / (best * 1e-3) / (clk * 1e3) / sms, clk / 1000, (int)cudaGetLastError());
    cudaFree(out);
}
int main() {
    run<0>("DFMA", 1);
    run<1>("DMMA m8n8k4", 8.0 * 8 * 4 / 32);
    run<2>("DMMA m16n8k8", 16.0 * 8 * 8 / 32);
    run<3>("DMMA m16n8k16", 16.0 * 8 * 16 / 32);
    run<4>("DMMA m8n8k4 + DFMA (sum)", 8.0 * 8 * 4 / 32 + 1);
    return 0;
}
