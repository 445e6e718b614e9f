// Throughput probe for the arithmetic pipes that matter to 64-bit modular arithmetic on B200:
// IMAD (32-bit), IMAD.WIDE (32x32->64), IADD3, DFMA, DADD, DMUL.  Prints ops/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define ITERS 4096
template<int OP> __global__ void k(u64* out, u64 seed) {
    u32 a[8]; u64 w[8]; double d[8];
    for (int i=0;i<8;++i){ a[i]=(u32)(seed+threadIdx.x*7+i); w[i]=seed*(i+3)+threadIdx.x; d[i]=(double)(threadIdx.x+i)+0.5; }
    u32 m=(u32)seed|1; double dm = 1.0000001, dc=0.25;
    for (int it=0; it<ITERS; ++it) {
#pragma unroll
        for (int i=0;i<8;++i) {
            if (OP==0) a[i] = a[i]*m + a[(i+1)&7];                                   // IMAD
            else if (OP==1) w[i] = (u64)(u32)w[i]*m + w[i];                           // IMAD.WIDE
            else if (OP==2) a[i] = a[i] + m + (a[(i+1)&7]^it);                        // IADD3/LOP3
            else if (OP==3) d[i] = fma(d[i], dm, dc);                                 // DFMA
            else if (OP==4) d[i] = d[i] + dc;                                         // DADD
            else if (OP==5) d[i] = d[i] * dm;                                         // DMUL
            else if (OP==6) w[i] = __umul64hi(w[i], seed) + w[i];                     // mul.hi.u64
            else if (OP==7) { d[i] = fma(d[i], dm, dc); a[i] = a[i]*m + a[(i+1)&7]; } // DFMA + IMAD together
            else if (OP==8) { d[i] = fma(d[i], dm, dc); a[i] = a[i] + m + (a[(i+1)&7]^it); } // DFMA + ALU
        }
    }
    u64 r=0; for (int i=0;i<8;++i) r += a[i] + w[i] + (u64)d[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int OP> void run(const char* name, int opsPerIter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    u64* out; cudaMalloc(&out, sizeof(u64)*sms*8*256);
    cudaEvent_t s,e; cudaEventCreate(&s); cudaEventCreate(&e);
    k<OP><<<sms*8,256>>>(out, 12345); cudaDeviceSynchronize();
    float best=1e9;
    for (int r=0;r<5;++r){ cudaEventRecord(s); k<OP><<<sms*8,256>>>(out, 12345+r); cudaEventRecord(e); cudaEventSynchronize(e); float ms; cudaEventElapsedTime(&ms,s,e); if(ms<best)best=ms; }
    double ops = (double)sms*8*256*ITERS*8*opsPerIter;
    printf("%-22s %8.3f ms  %7.1f Gop/s  %6.1f ops/clk/SM (at %d MHz nominal)\n", name, best, ops/best/1e6, ops/(best*1e-3)/(clk*1e3)/sms, clk/1000);
    cudaFree(out);
}
int main(){
    run<0>("IMAD", 1); run<1>("IMAD.WIDE", 1); run<2>("IADD3+LOP3", 2); run<3>("DFMA", 1); run<4>("DADD", 1); run<5>("DMUL", 1);
    run<6>("mul.hi.u64", 1); run<7>("DFMA+IMAD (pairs)", 1); run<8>("DFMA+ALU (pairs)", 1);
    return 0;
}
