// compat.h -- one source, two builds.
//
//  * nvcc (product): plain CUDA for sm_100a.
//  * g++ -DFHE_EMU (tests only, tests/emu/): a functional simulator that runs the SAME
//    kernel bodies on the host so index arithmetic / modular arithmetic can be validated in
//    a container without a GPU.  Every CUDA thread of a block is a ucontext fiber,
//    __syncthreads() yields to the block scheduler, __shared__ is thread_local static
//    storage, blocks are spread over host threads.  It is never built into, loaded by or
//    selected from the product package.
#pragma once
#include <cstdint>
#include <cstddef>
#include <utility>

typedef unsigned long long u64;
typedef unsigned int u32;

#ifndef FHE_EMU
// ----------------------------------------------------------------------------- CUDA
#include <cuda_runtime.h>
#define FHE_HD __host__ __device__ __forceinline__
#define FHE_D __device__ __forceinline__
#define FHE_SHARED __shared__
#define FHE_DYN_SHARED(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; type* name = reinterpret_cast<type*>(name##_raw)

template <typename... KArgs, typename... Args>
inline void fhe_launch(void (*k)(KArgs...), dim3 g, dim3 b, size_t smem, cudaStream_t s, Args&&... args) {
    k<<<g, b, smem, s>>>(std::forward<Args>(args)...);
}

FHE_D u64 umulhi64(u64 a, u64 b) { return __umul64hi(a, b); }
FHE_D u32 brev32(u32 x) { return __brev(x); }
// FP64 tensor-core product of one warp: C[8x8] += A[8x4] B[4x8].  With g = lane >> 2, t = lane & 3 a lane holds
// a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].  Same 64 FMA/clk/SM as DFMA (tools/ubench/dmma.cu)
// but eight FMAs per lane per instruction with both operands in registers.
FHE_D void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#else
// ----------------------------------------------------------------------------- emulation
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define FHE_HD inline
#define FHE_D inline
#define FHE_SHARED static thread_local
#define FHE_DYN_SHARED(type, name) type* name = reinterpret_cast<type*>(fhe_emu_dyn_smem())
char* fhe_emu_dyn_smem();
void fhe_emu_set_dyn_smem(size_t bytes);

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct ulonglong2 { u64 x, y; };
inline ulonglong2 make_ulonglong2(u64 x, u64 y) { ulonglong2 r; r.x = x; r.y = y; return r; }

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
void __syncthreads();

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

void fhe_emu_launch(dim3 g, dim3 b, const std::function<void()>& body);

template <typename... KArgs, typename... Args>
inline void fhe_launch(void (*k)(KArgs...), dim3 g, dim3 b, size_t smem, cudaStream_t, Args&&... args) {
    fhe_emu_set_dyn_smem(smem);
    fhe_emu_launch(g, b, [=]() { k(args...); });
}

inline u64 umulhi64(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
inline u32 brev32(u32 x) {
    u32 r = 0;
    for (int i = 0; i < 32; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
// warp-collective 8x8x4 product: lanes exchange their fragments through a per-block buffer between two barriers
// (every thread of the block must execute the same sequence of dmma884 calls, which the kernels guarantee)
#include <cmath>
inline void dmma884(double& c0, double& c1, double a, double b) {
    static thread_local double ea[1024], eb[1024];
    const int tid = (int)threadIdx.x, wb = tid & ~31, g = (tid & 31) >> 2, t = tid & 3;
    ea[tid] = a; eb[tid] = b;
    __syncthreads();
    for (int k = 0; k < 4; ++k) {
        c0 = std::fma(ea[wb + g * 4 + k], eb[wb + (2 * t) * 4 + k], c0);
        c1 = std::fma(ea[wb + g * 4 + k], eb[wb + (2 * t + 1) * 4 + k], c1);
    }
    __syncthreads();
}
template <typename T> inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T max(T a, T b) { return a > b ? a : b; }
#endif
