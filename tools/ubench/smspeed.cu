// Is SM speed position-dependent?  Per-SM duration of (0) a pure DFMA loop, (1) a streaming read of
// 512 KiB rows from a large buffer, (2) DFMA + reads together.  One line per group of 16 smids.
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE> __global__ void k(u64* rec, const double* buf, size_t nrows, double* sink) {
    u64 t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    double d[8]; for (int i = 0; i < 8; ++i) d[i] = threadIdx.x + i + 0.5;
    double acc = 0;
    for (int it = 0; it < 64; ++it) {
        if (MODE != 0) {
            size_t row = ((size_t)blockIdx.x * 64 + it) * 7919 % nrows;
            const double* p = buf + row * 65536;
            for (int j = 0; j < 16; ++j) acc += __ldg(p + (size_t)j * 4096 + threadIdx.x);
        }
        if (MODE != 1) {
#pragma unroll 1
            for (int r = 0; r < 64; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = fma(d[i], 1.0000001, 0.25);
        }
    }
    for (int i = 0; i < 8; ++i) acc += d[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    u64 t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) { rec[3 * blockIdx.x] = smid; rec[3 * blockIdx.x + 1] = t0; rec[3 * blockIdx.x + 2] = t1; }
}
template <int MODE> void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 6, nrows = 1024;
    u64* rec; cudaMalloc(&rec, 24 * grid);
    double* buf; cudaMalloc(&buf, (size_t)nrows * 65536 * 8); cudaMemset(buf, 0, (size_t)nrows * 65536 * 8);
    double* sink; cudaMalloc(&sink, 8 * grid * 128);
    for (int r = 0; r < 3; ++r) k<MODE><<<grid, 128>>>(rec, buf, nrows, sink);
    cudaDeviceSynchronize();
    std::vector<u64> h(3 * grid); cudaMemcpy(h.data(), rec, 24 * grid, cudaMemcpyDeviceToHost);
    std::vector<double> sum(sms, 0); std::vector<int> cnt(sms, 0);
    for (int b = 0; b < grid; ++b) { sum[h[3 * b]] += (h[3 * b + 2] - h[3 * b + 1]) / 1e3; cnt[h[3 * b]]++; }
    printf("%-14s", name);
    for (int s0 = 0; s0 < sms; s0 += 16) {
        double m = 0; int c = 0;
        for (int s = s0; s < std::min(sms, s0 + 16); ++s) if (cnt[s]) { m += sum[s] / cnt[s]; ++c; }
        printf(" %7.1f", c ? m / c : 0.0);
    }
    double mn = 1e30, mx = 0; for (int s = 0; s < sms; ++s) if (cnt[s]) { mn = std::min(mn, sum[s] / cnt[s]); mx = std::max(mx, sum[s] / cnt[s]); }
    printf("   us per CTA by smid/16;  per-SM min %.1f max %.1f\n", mn, mx);
    cudaFree(rec); cudaFree(buf); cudaFree(sink);
}
int main() { run<0>("dfma"); run<1>("reads"); run<2>("dfma+reads"); return 0; }
