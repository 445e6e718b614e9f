// Does round-to-integer of a double (FRND.F64 / F2F) issue on the FP64 pipe or beside it?
// Compares 8 DFMA per step with 7 DFMA + 1 rint per step, and rint alone.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(double* out, double seed) {
    double d[8]; for (int i = 0; i < 8; ++i) d[i] = threadIdx.x + i + seed;
    const double m = 1.0000001, c = 0.25;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { double x = d[i];
#pragma unroll
                for (int r = 0; r < 8; ++r) x = fma(x, m, c); d[i] = x; }
            else if (MODE == 1) { double x = d[i];
#pragma unroll
                for (int r = 0; r < 7; ++r) x = fma(x, m, c); d[i] = rint(x); }
            else if (MODE == 2) { d[i] = rint(d[i] * m); }
            else if (MODE == 3) { double x = d[i];
#pragma unroll
                for (int r = 0; r < 6; ++r) x = fma(x, m, c); d[i] = rint(x) + 6755399441055744.0; }
        }
    }
    double r = 0; for (int i = 0; i < 8; ++i) r += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double fp64ops, double rnd) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8 * sms * 8 * 256);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    k<MODE><<<sms * 8, 256>>>(out, 1.5); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(s); k<MODE><<<sms * 8, 256>>>(out, 1.5 + r); cudaEventRecord(e); cudaEventSynchronize(e); float ms; cudaEventElapsedTime(&ms, s, e); if (ms < best) best = ms; }
    double steps = (double)sms * 8 * 256 * ITERS * 8;
    printf("%-28s %8.3f ms   fp64 %.1f /clk/SM   rint %.1f /clk/SM (at 1965 MHz)\n", name, best,
           steps * fp64ops / (best * 1e-3) / 1.965e9 / sms, steps * rnd / (best * 1e-3) / 1.965e9 / sms);
}
int main() {
    run<0>("8 DFMA", 8, 0); run<1>("7 DFMA + 1 rint", 7, 1); run<2>("DMUL + rint", 1, 1); run<3>("6 DFMA + rint + DADD", 7, 1);
    return 0;
}
