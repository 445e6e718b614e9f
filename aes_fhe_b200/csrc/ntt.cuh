// ntt.cuh -- negacyclic NTT / iNTT over Z_q[X]/(X^N+1), N = 2^log_n (12..16), 64-bit
// residues, bit-reversed spectrum order (the layout every other kernel assumes).
//
// Decomposition (B200: 148 SMs, 227 KB smem, HBM-resident 512 KiB limbs):
//   a limb is viewed as R x 256 (R = N/256).  Two kernels per transform:
//     pass A : the log2(R) stages whose butterfly stride is >= 256 -- a CTA owns a tile of
//              COLS columns (all R rows), 16 elements per thread in registers, one
//              shared-memory exchange between two radix-16 rounds;
//     pass B : the last 8 stages on 16 contiguous 256-element rows per CTA.
//   Every global access is a full 128-byte segment per half-warp.  Twiddles are Shoup pairs
//   (w, floor(w 2^64/q)) read as one 16-byte word; butterflies are Harvey-lazy ([0,4q)
//   forward, [0,2q) inverse) so the only conditional corrections are at the very end.
//   The first load and the last store go through functors, which is how rescale, ModDown
//   and the base conversions are fused into the transform instead of being extra passes.
#pragma once
#include "modarith.cuh"

struct DevTables {
    const Modulus* mod;            // [n_q + n_p]
    const ShoupConst* tw_fwd;      // [n_q + n_p][N]   psi^bitrev(k)
    const ShoupConst* tw_inv;      // [n_q + n_p][N]   psi^-bitrev(k)
    const ShoupConst* inv_final;   // [n_q + n_p][2]   {N^-1, psi^-bitrev(1) * N^-1}
    int log_n;
    int n_q, n_p;
};

// Which modulus a row belongs to, and where it lives.  Rows are grouped in "polys" of
// rows_per_poly limbs; limb slot j = j0 + row % rows_per_poly maps to modulus j (< nq) or
// p_base + (j - nq).  skip_alpha > 0 marks the ModUp layout [batch][digit][nq + n_p]: the rows of a
// digit's own limbs are not transformed.
struct RowMap {
    int rows_per_poly;
    int j0;
    int nq;
    int p_base;
    int skip_alpha;
    int digits;                    // beta, when skip_alpha > 0 (rows are [batch][digit][nq + n_p])
    int log_n;
    FHE_D int mod_id(int row) const {
        int j = j0 + row % rows_per_poly;
        if (skip_alpha > 0) {
            int lo = ((row / rows_per_poly) % digits) * skip_alpha;
            if (j >= lo && j < lo + skip_alpha && j < nq) return -1;
        }
        return j < nq ? j : p_base + (j - nq);
    }
};

// ------------------------------------------------------------------ butterflies
FHE_D void ct_bfly(u64& a, u64& b, const ShoupConst w, u64 q, u64 two_q) {
    u64 u = a >= two_q ? a - two_q : a;
    u64 t = mul_shoup_lazy(b, w.w, w.ws, q);
    a = u + t;
    b = u - t + two_q;
}
FHE_D void gs_bfly(u64& a, u64& b, const ShoupConst w, u64 q, u64 two_q) {
    u64 s = a + b;
    u64 d = a - b + two_q;
    a = s >= two_q ? s - two_q : s;
    b = mul_shoup_lazy(d, w.w, w.ws, q);
}

FHE_D ShoupConst ld_tw(const ShoupConst* p) {
#ifndef FHE_EMU
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    ShoupConst r; r.w = v.x; r.ws = v.y; return r;
#else
    return *p;
#endif
}

// twiddle index of global stage S (1-based) for the butterfly group holding element gidx
FHE_D u32 tw_index(int S, u32 gidx, int log_n) { return (1u << (S - 1)) + (gidx >> (log_n - S + 1)); }

// Forward: levels with half = 8, 4, ... down to HALF_END on the 16 register values.
// S0 is the global stage of the first level; gidx(i) the global element index of x[i].
template <int HALF_END, class GIdx>
FHE_D void ct_radix16(u64 (&x)[16], int S0, GIdx gidx, const ShoupConst* tw, int log_n, u64 q, u64 two_q) {
    int S = S0;
#pragma unroll
    for (int half = 8; half >= HALF_END; half >>= 1) {
#pragma unroll
        for (int base = 0; base < 16; base += 2 * half) {
            ShoupConst w = ld_tw(tw + tw_index(S, gidx(base), log_n));
#pragma unroll
            for (int k = 0; k < half; ++k) ct_bfly(x[base + k], x[base + k + half], w, q, two_q);
        }
        ++S;
    }
}
// Inverse: levels with half = HALF_START, ..., 8; S0 is the global stage of the first level.
template <int HALF_START, bool LAST_IS_FINAL, class GIdx>
FHE_D void gs_radix16(u64 (&x)[16], int S0, GIdx gidx, const ShoupConst* tw, int log_n, u64 q, u64 two_q,
                      ShoupConst fin0, ShoupConst fin1) {
    int S = S0;
#pragma unroll
    for (int half = HALF_START; half <= 8; half <<= 1) {
        if (LAST_IS_FINAL && half == 8) {
            // global stage 1 merged with the N^-1 (and any caller-supplied) scaling
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                u64 s = x[k] + x[k + 8];
                u64 d = x[k] - x[k + 8] + two_q;
                u64 r0 = mul_shoup_lazy(s, fin0.w, fin0.ws, q);
                u64 r1 = mul_shoup_lazy(d, fin1.w, fin1.ws, q);
                x[k] = r0 >= q ? r0 - q : r0;
                x[k + 8] = r1 >= q ? r1 - q : r1;
            }
        } else {
#pragma unroll
            for (int base = 0; base < 16; base += 2 * half) {
                ShoupConst w = ld_tw(tw + tw_index(S, gidx(base), log_n));
#pragma unroll
                for (int k = 0; k < half; ++k) gs_bfly(x[base + k], x[base + k + half], w, q, two_q);
            }
        }
        --S;
    }
}

FHE_D int pad16(int a) { return a + (a >> 4); }

// ------------------------------------------------------------------ load / store functors
// (each functor carries its own poly stride so source, scratch and destination may have
//  different layouts; the RowMap only says which modulus a row uses)
struct LoadPlain {
    const u64* src; long long poly_stride;
    FHE_D u64 operator()(const RowMap& map, int row, u32 idx, int, const Modulus&) const {
        return src[(size_t)(row / map.rows_per_poly) * (size_t)poly_stride + ((size_t)(row % map.rows_per_poly) << map.log_n) + idx];
    }
};
struct StorePlain {
    u64* dst; long long poly_stride;
    FHE_D void operator()(const RowMap& map, int row, u32 idx, u64 v, int, const Modulus&) const {
        dst[(size_t)(row / map.rows_per_poly) * (size_t)poly_stride + ((size_t)(row % map.rows_per_poly) << map.log_n) + idx] = v;
    }
};

// ------------------------------------------------------------------ forward kernels
template <int LOG_R, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(256) ntt_fwd_pass_a(DevTables T, RowMap map, LoadOp ld, StoreOp st) {
    constexpr int R = 1 << LOG_R, G = R / 16, COLS = 256 / G, LEV1 = LOG_R - 4;
    FHE_SHARED u64 sm[4096];
    const int row = blockIdx.y;
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const Modulus M = T.mod[mid];
    const u64 q = M.q, two_q = 2 * M.q;
    const int log_n = T.log_n;
    const ShoupConst* tw = T.tw_fwd + ((size_t)mid << log_n);
    const int tid = threadIdx.x, cc = tid % COLS, g = tid / COLS;
    const u32 c = blockIdx.x * COLS + cc;
    u64 x[16];
    if (LEV1 > 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ld(map, row, (u32)((g + G * i) << 8) + c, mid, M);
        ct_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1))>(x, 1, [&](int i) { return (u32)((g + G * i) << 8) + c; }, tw, log_n, q, two_q);
#pragma unroll
        for (int i = 0; i < 16; ++i) sm[(g + G * i) * COLS + cc] = x[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = sm[(16 * g + i) * COLS + cc];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ld(map, row, (u32)((16 * g + i) << 8) + c, mid, M);
    }
    ct_radix16<1>(x, LEV1 + 1, [&](int i) { return (u32)((16 * g + i) << 8) + c; }, tw, log_n, q, two_q);
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, row, (u32)((16 * g + i) << 8) + c, x[i], mid, M);
}

template <class LoadOp, class StoreOp>
__global__ void __launch_bounds__(256) ntt_fwd_pass_b(DevTables T, RowMap map, LoadOp ld, StoreOp st) {
    FHE_SHARED u64 sm[4096 + 256];
    const int row = blockIdx.y;
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const Modulus M = T.mod[mid];
    const u64 q = M.q, two_q = 2 * M.q;
    const int log_n = T.log_n;
    const ShoupConst* tw = T.tw_fwd + ((size_t)mid << log_n);
    const int tid = threadIdx.x, l16 = tid & 15, rr = tid >> 4;
    const u32 base = (u32)(blockIdx.x * 16 + rr) << 8;
    u64 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ld(map, row, base + l16 + 16 * i, mid, M);
    ct_radix16<1>(x, log_n - 7, [&](int i) { return base + l16 + 16 * i; }, tw, log_n, q, two_q);
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + l16 + 16 * i)] = x[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[pad16(rr * 256 + 16 * l16 + i)];
    ct_radix16<1>(x, log_n - 3, [&](int i) { return base + 16 * l16 + i; }, tw, log_n, q, two_q);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        u64 v = x[i];
        v = v >= two_q ? v - two_q : v;
        v = v >= q ? v - q : v;
        sm[pad16(rr * 256 + 16 * l16 + i)] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, row, base + l16 + 16 * i, sm[pad16(rr * 256 + l16 + 16 * i)], mid, M);
}

// ------------------------------------------------------------------ inverse kernels
// pass B' : stages log_n .. log_n-7 (strides 1..128) on 16 contiguous rows of 256.
template <class LoadOp, class StoreOp>
__global__ void __launch_bounds__(256) ntt_inv_pass_b(DevTables T, RowMap map, LoadOp ld, StoreOp st) {
    FHE_SHARED u64 sm[4096 + 256];
    const int row = blockIdx.y;
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const Modulus M = T.mod[mid];
    const u64 q = M.q, two_q = 2 * M.q;
    const int log_n = T.log_n;
    const ShoupConst* tw = T.tw_inv + ((size_t)mid << log_n);
    const int tid = threadIdx.x, l16 = tid & 15, rr = tid >> 4;
    const u32 base = (u32)(blockIdx.x * 16 + rr) << 8;
    u64 x[16];
    ShoupConst dummy; dummy.w = 0; dummy.ws = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + l16 + 16 * i)] = ld(map, row, base + l16 + 16 * i, mid, M);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[pad16(rr * 256 + 16 * l16 + i)];
    gs_radix16<1, false>(x, log_n, [&](int i) { return base + 16 * l16 + i; }, tw, log_n, q, two_q, dummy, dummy);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + 16 * l16 + i)] = x[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[pad16(rr * 256 + l16 + 16 * i)];
    gs_radix16<1, false>(x, log_n - 4, [&](int i) { return base + l16 + 16 * i; }, tw, log_n, q, two_q, dummy, dummy);
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, row, base + l16 + 16 * i, x[i], mid, M);
}

// pass A' : stages log_n-8 .. 1 (row strides 1..R/2); the final stage carries the scaling
// constants scale[row % rows_per_poly] = {c * N^-1, c * psi^-bitrev(1) * N^-1} (c = 1 when
// scale == nullptr).
template <int LOG_R, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(256) ntt_inv_pass_a(DevTables T, RowMap map, LoadOp ld, StoreOp st,
                                                      const ShoupConst* scale) {
    constexpr int R = 1 << LOG_R, G = R / 16, COLS = 256 / G, LEV1 = LOG_R - 4;
    FHE_SHARED u64 sm[4096];
    const int row = blockIdx.y;
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const Modulus M = T.mod[mid];
    const u64 q = M.q, two_q = 2 * M.q;
    const int log_n = T.log_n;
    const ShoupConst* tw = T.tw_inv + ((size_t)mid << log_n);
    const ShoupConst* fin = scale ? scale + 2 * (size_t)(row % map.rows_per_poly) : T.inv_final + 2 * (size_t)mid;
    const ShoupConst fin0 = fin[0], fin1 = fin[1];
    const int tid = threadIdx.x, cc = tid % COLS, g = tid / COLS;
    const u32 c = blockIdx.x * COLS + cc;
    u64 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ld(map, row, (u32)((16 * g + i) << 8) + c, mid, M);
    if (LEV1 > 0) {
        gs_radix16<1, false>(x, LOG_R, [&](int i) { return (u32)((16 * g + i) << 8) + c; }, tw, log_n, q, two_q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) sm[(16 * g + i) * COLS + cc] = x[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = sm[(g + G * i) * COLS + cc];
        gs_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1)), true>(x, LEV1, [&](int i) { return (u32)((g + G * i) << 8) + c; }, tw, log_n, q, two_q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, row, (u32)((g + G * i) << 8) + c, x[i], mid, M);
    } else {
        gs_radix16<1, true>(x, LOG_R, [&](int i) { return (u32)((16 * g + i) << 8) + c; }, tw, log_n, q, two_q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, row, (u32)((16 * g + i) << 8) + c, x[i], mid, M);
    }
}

// ------------------------------------------------------------------ host launchers
// `work` ([.., work_stride] layout) receives the lazy intermediate between the two passes; it
// may alias the destination, or the source for in-place use.
template <class LoadOp, class StoreOp>
inline void ntt_forward(const DevTables& T, const RowMap& map, int rows, LoadOp ld, u64* work,
                        long long work_stride, StoreOp st, cudaStream_t s) {
    const int log_r = T.log_n - 8;
    dim3 grid(1u << (log_r - 4), rows), block(256);
    StorePlain sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadPlain lp; lp.src = work; lp.poly_stride = work_stride;
    switch (log_r) {
        case 4: fhe_launch(ntt_fwd_pass_a<4, LoadOp, StorePlain>, grid, block, 0, s, T, map, ld, sp); break;
        case 5: fhe_launch(ntt_fwd_pass_a<5, LoadOp, StorePlain>, grid, block, 0, s, T, map, ld, sp); break;
        case 6: fhe_launch(ntt_fwd_pass_a<6, LoadOp, StorePlain>, grid, block, 0, s, T, map, ld, sp); break;
        case 7: fhe_launch(ntt_fwd_pass_a<7, LoadOp, StorePlain>, grid, block, 0, s, T, map, ld, sp); break;
        default: fhe_launch(ntt_fwd_pass_a<8, LoadOp, StorePlain>, grid, block, 0, s, T, map, ld, sp); break;
    }
    fhe_launch(ntt_fwd_pass_b<LoadPlain, StoreOp>, grid, block, 0, s, T, map, lp, st);
}

template <class LoadOp, class StoreOp>
inline void ntt_inverse(const DevTables& T, const RowMap& map, int rows, LoadOp ld, u64* work,
                        long long work_stride, StoreOp st, const ShoupConst* scale, cudaStream_t s) {
    const int log_r = T.log_n - 8;
    dim3 grid(1u << (log_r - 4), rows), block(256);
    StorePlain sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadPlain lp; lp.src = work; lp.poly_stride = work_stride;
    fhe_launch(ntt_inv_pass_b<LoadOp, StorePlain>, grid, block, 0, s, T, map, ld, sp);
    switch (log_r) {
        case 4: fhe_launch(ntt_inv_pass_a<4, LoadPlain, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 5: fhe_launch(ntt_inv_pass_a<5, LoadPlain, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 6: fhe_launch(ntt_inv_pass_a<6, LoadPlain, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 7: fhe_launch(ntt_inv_pass_a<7, LoadPlain, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        default: fhe_launch(ntt_inv_pass_a<8, LoadPlain, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
    }
}
