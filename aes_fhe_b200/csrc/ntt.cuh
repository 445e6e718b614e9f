// ntt.cuh -- negacyclic NTT / iNTT over Z_q[X]/(X^N+1), N = 2^log_n (12..16), residues
// stored as canonical 64-bit words, bit-reversed spectrum order (the layout every other
// kernel assumes), arithmetic on the FP64 pipe (modarith.cuh).
//
// Decomposition (B200: 148 SMs, 227 KB smem, HBM-resident 512 KiB limbs):
//   a limb is viewed as R x 256 (R = N/256).  Two kernels per transform:
//     pass A : the log2(R) stages whose butterfly stride is >= 256 -- a CTA owns a tile of
//              COLS columns (all R rows), 16 elements per thread in registers, one
//              shared-memory exchange between two radix-16 rounds;
//     pass B : the last 8 stages on 16 contiguous 256-element rows per CTA.
//   Every global access is a full 128-byte segment per half-warp.  Twiddle tables hold w
//   only (w/q is one multiply by RN(1/q)); the per-thread twiddles of the last four
//   stages are fetched with 16-byte loads.  Values are signed and lazy: the forward
//   transform never corrects (|x| < 12q), the inverse reduces once per radix-16 round.
//   The lazy intermediate between the two passes is written as raw doubles.
//   The first load and the last store go through functors, which is how rescale, ModDown
//   and the base conversions are fused into the transform instead of being extra passes.
#pragma once
#include "modarith.cuh"

#ifndef FHE_PASSA_MINB
#define FHE_PASSA_MINB 1
#endif
#ifndef FHE_PASSB_MINB
#define FHE_PASSB_MINB 1
#endif

struct DevTables {
    const Modulus* mod;            // [n_q + n_p]
    const double* tw_fwd;          // [n_q + n_p][N]   psi^bitrev(k)   (w only: w/q = w * RN(1/q)
    const double* tw_inv;          // [n_q + n_p][N]   psi^-bitrev(k)   is one DMUL, half the bytes)
    const ConstF* inv_final;       // [n_q + n_p][2]   {N^-1, psi^-bitrev(1) * N^-1}
    int log_n;
    int n_q, n_p;
};

// Which modulus a row belongs to.  Rows are grouped in blocks of rows_per_poly limbs; limb
// slot j = j0 + row % rows_per_poly maps to modulus j (< nq) or p_base + (j - nq).
// skip_alpha > 0 marks the ModUp layout [batch][digit][nq + n_p]: the rows of a digit's own
// limbs are not transformed.
struct RowRef { int blk, j; };     // row = blk * rows_per_poly + j: poly block and limb slot
struct RowMap {
    int rows_per_poly;
    int j0;
    int nq;
    int p_base;
    int skip_alpha;
    int digits;                    // beta, when skip_alpha > 0
    int log_n;
    int n_blocks;                  // rows / rows_per_poly: launch order is limb-major, block-minor,
                                   // so the CTAs sharing one twiddle table are co-resident (L2 reuse)
    FHE_D int launch_row(int y) const { return (y % n_blocks) * rows_per_poly + y / n_blocks; }
    FHE_D RowRef ref(int row) const { RowRef r; r.blk = row / rows_per_poly; r.j = row - r.blk * rows_per_poly; return r; }
    // the same classification from (limb slot j, digit index of the block); no divisions
    FHE_D int mod_id_of(int j_slot, int dig) const {
        const int j = j0 + j_slot;
        if (skip_alpha > 0) {
            const int lo = dig * skip_alpha;
            if (j >= lo && j < lo + skip_alpha && j < nq) return -1;
        }
        return j < nq ? j : p_base + (j - nq);
    }
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
FHE_D void ct_bfly(double& a, double& b, const ConstF w, double q) {
    const double t = mulmod_const(b, w, q);
    b = d_add(a, -t);
    a = d_add(a, t);
}
FHE_D void gs_bfly(double& a, double& b, const ConstF w, double q) {
    const double s = d_add(a, b);
    const double d = d_add(a, -b);
    a = s;
    b = mulmod_const(d, w, q);
}

FHE_D void prefetch_l2(const void* p) {
#ifndef FHE_EMU
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}
FHE_D ConstF mk_tw(double w, double qinv) { ConstF r; r.w = w; r.wq = d_mul(w, qinv); return r; }
FHE_D double ld_d(const double* p) {
#ifndef FHE_EMU
    return __ldg(p);
#else
    return *p;
#endif
}
// 15 twiddles of the four stride-8/4/2/1 stages of one thread's 16 contiguous elements, read
// with 16-byte loads: stage with half h uses w[8/h - 1 + group].
struct Tw15 { double w[15]; };
FHE_D Tw15 ld_tw15(const double* tw, u32 first_elem, int log_n) {
    Tw15 t;
    const u32 e = first_elem >> 4;                       // 16-element block index within the limb
    const u32 n16 = 1u << (log_n - 4);
    t.w[0] = ld_d(tw + n16 + e);
#ifndef FHE_EMU
    const double2 a = __ldg(reinterpret_cast<const double2*>(tw + 2 * (n16 + e)));
    t.w[1] = a.x; t.w[2] = a.y;
    const double2* p4 = reinterpret_cast<const double2*>(tw + 4 * (n16 + e));
    const double2 b0 = __ldg(p4), b1 = __ldg(p4 + 1);
    t.w[3] = b0.x; t.w[4] = b0.y; t.w[5] = b1.x; t.w[6] = b1.y;
    const double2* p8 = reinterpret_cast<const double2*>(tw + 8 * (n16 + e));
#pragma unroll
    for (int j = 0; j < 4; ++j) { const double2 c = __ldg(p8 + j); t.w[7 + 2 * j] = c.x; t.w[8 + 2 * j] = c.y; }
#else
    for (int j = 0; j < 2; ++j) t.w[1 + j] = tw[2 * (n16 + e) + j];
    for (int j = 0; j < 4; ++j) t.w[3 + j] = tw[4 * (n16 + e) + j];
    for (int j = 0; j < 8; ++j) t.w[7 + j] = tw[8 * (n16 + e) + j];
#endif
    return t;
}

// twiddle index of global stage S (1-based) for the butterfly group holding element gidx
FHE_D u32 tw_index(int S, u32 gidx, int log_n) { return (1u << (S - 1)) + (gidx >> (log_n - S + 1)); }

// Forward: levels with half = 8, 4, ... down to HALF_END on the 16 register values.
// S0 is the global stage of the first level; gidx(i) the global element index of x[i].
// twp(S, half, base) returns the twiddle of the group starting at register `base`.
template <int HALF_END, class TwP>
FHE_D void ct_radix16(double (&x)[16], int S0, TwP twp, double q) {
    int S = S0;
#pragma unroll
    for (int half = 8; half >= HALF_END; half >>= 1) {
#pragma unroll
        for (int base = 0; base < 16; base += 2 * half) {
            const ConstF w = twp(S, half, base);
#pragma unroll
            for (int k = 0; k < half; ++k) ct_bfly(x[base + k], x[base + k + half], w, q);
        }
        ++S;
    }
}
// Inverse: levels with half = HALF_START, ..., 8; S0 is the global stage of the first level.
// Inputs must satisfy |x| <= q; outputs are <= 16q (sum path) -- callers reduce per round.
template <int HALF_START, bool LAST_IS_FINAL, class TwP>
FHE_D void gs_radix16(double (&x)[16], int S0, TwP twp, double q, ConstF fin0, ConstF fin1) {
    int S = S0;
#pragma unroll
    for (int half = HALF_START; half <= 8; half <<= 1) {
        if (LAST_IS_FINAL && half == 8) {
            // global stage 1 merged with the N^-1 (and any caller-supplied) scaling
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double s = d_add(x[k], x[k + 8]);
                const double d = d_add(x[k], -x[k + 8]);
                x[k] = canon(mulmod_const(s, fin0, q), q);
                x[k + 8] = canon(mulmod_const(d, fin1, q), q);
            }
        } else {
#pragma unroll
            for (int base = 0; base < 16; base += 2 * half) {
                const ConstF w = twp(S, half, base);
#pragma unroll
                for (int k = 0; k < half; ++k) gs_bfly(x[base + k], x[base + k + half], w, q);
            }
        }
        --S;
    }
}

FHE_D int pad16(int a) { return a + (a >> 4); }

// ------------------------------------------------------------------ load / store functors
// A LoadOp returns the element as a (small, signed) double; a StoreOp receives the canonical
// value in [0, q) as a double.  Each functor carries its own poly stride so source, scratch
// and destination may have different layouts.
FHE_D size_t row_off(const RowMap& map, RowRef row, long long poly_stride) {
    return (size_t)row.blk * (size_t)poly_stride + ((size_t)row.j << map.log_n);
}
struct LoadPlain {          // canonical residues
    const u64* src; long long poly_stride;
    FHE_D double operator()(const RowMap& map, RowRef row, int idx, int, const Modulus&) const {
        return u64_to_f(src[row_off(map, row, poly_stride) + idx]);
    }
    // hint: `count` consecutive elements starting at idx will be loaded soon (fused NTT, L2 prefetch)
    FHE_D void prefetch(const RowMap& map, RowRef row, int idx, int count) const {
        const u64* p = src + row_off(map, row, poly_stride) + idx;
        for (int k = 0; k < count; k += 16) prefetch_l2(p + k);
    }
};
struct StorePlain {
    u64* dst; long long poly_stride;
    FHE_D void operator()(const RowMap& map, RowRef row, int idx, double v, int, const Modulus&) const {
        dst[row_off(map, row, poly_stride) + idx] = f_to_u64(v);
    }
};
struct LoadRaw {            // lazy doubles written by the other pass
    const u64* src; long long poly_stride;
    FHE_D double operator()(const RowMap& map, RowRef row, int idx, int, const Modulus&) const {
        return bits_to_f(src[row_off(map, row, poly_stride) + idx]);
    }
};
struct StoreRaw {
    u64* dst; long long poly_stride;
    FHE_D void operator()(const RowMap& map, RowRef row, int idx, double v, int, const Modulus&) const {
        dst[row_off(map, row, poly_stride) + idx] = f_to_bits(v);
    }
};

// ------------------------------------------------------------------ forward kernels
template <int LOG_R, class LoadOp>
__global__ void __launch_bounds__(256, FHE_PASSA_MINB) ntt_fwd_pass_a(DevTables T, RowMap map, LoadOp ld, StoreRaw st) {
    constexpr int R = 1 << LOG_R, G = R / 16, COLS = 256 / G, LEV1 = LOG_R - 4;
    FHE_SHARED double sm[4096];
    const int row = map.launch_row(blockIdx.y);
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const RowRef rref = map.ref(row);
    const Modulus M = T.mod[mid];
    const double q = M.qd;
    const int log_n = T.log_n;
    const double* tw = T.tw_fwd + ((size_t)mid << log_n);
    const double qinv = M.qinv;
    const int tid = threadIdx.x, cc = tid % COLS, g = tid / COLS;
    const u32 c = blockIdx.x * COLS + cc;
    double x[16];
    if (LEV1 > 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, (u32)((g + G * i) << 8) + c, mid, M);
        ct_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1))>(x, 1, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, (u32)((g + G * i) << 8) + c, log_n)), qinv); }, q);
#pragma unroll
        for (int i = 0; i < 16; ++i) sm[(g + G * i) * COLS + cc] = x[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = sm[(16 * g + i) * COLS + cc];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, (u32)((16 * g + i) << 8) + c, mid, M);
    }
    ct_radix16<1>(x, LEV1 + 1, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, (u32)((16 * g + i) << 8) + c, log_n)), qinv); }, q);
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, rref, (u32)((16 * g + i) << 8) + c, x[i], mid, M);
}

template <class StoreOp>
__global__ void __launch_bounds__(256, FHE_PASSB_MINB) ntt_fwd_pass_b(DevTables T, RowMap map, LoadRaw ld, StoreOp st) {
    FHE_SHARED double sm[4096 + 256];
    const int row = map.launch_row(blockIdx.y);
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const RowRef rref = map.ref(row);
    const Modulus M = T.mod[mid];
    const double q = M.qd;
    const int log_n = T.log_n;
    const double* tw = T.tw_fwd + ((size_t)mid << log_n);
    const double qinv = M.qinv;
    const int tid = threadIdx.x, l16 = tid & 15, rr = tid >> 4;
    const u32 base = (u32)(blockIdx.x * 16 + rr) << 8;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, base + l16 + 16 * i, mid, M);
    const Tw15 t2 = ld_tw15(tw, base + 16 * l16, log_n);          // issued early, used in round 2
    ct_radix16<1>(x, log_n - 7, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, base + l16 + 16 * i, log_n)), qinv); }, q);
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + l16 + 16 * i)] = x[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[pad16(rr * 256 + 16 * l16 + i)];
    ct_radix16<1>(x, log_n - 3, [&](int, int half, int i) { return mk_tw(t2.w[8 / half - 1 + i / (2 * half)], qinv); }, q);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + 16 * l16 + i)] = reduce_canon(x[i], q, M.qinv);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, rref, base + l16 + 16 * i, sm[pad16(rr * 256 + l16 + 16 * i)], mid, M);
}

// ------------------------------------------------------------------ inverse kernels
// pass B' : stages log_n .. log_n-7 (strides 1..128) on 16 contiguous rows of 256.
template <class LoadOp>
__global__ void __launch_bounds__(256, FHE_PASSB_MINB) ntt_inv_pass_b(DevTables T, RowMap map, LoadOp ld, StoreRaw st) {
    FHE_SHARED double sm[4096 + 256];
    const int row = map.launch_row(blockIdx.y);
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const RowRef rref = map.ref(row);
    const Modulus M = T.mod[mid];
    const double q = M.qd, qinv = M.qinv;
    const int log_n = T.log_n;
    const double* tw = T.tw_inv + ((size_t)mid << log_n);
    const int tid = threadIdx.x, l16 = tid & 15, rr = tid >> 4;
    const u32 base = (u32)(blockIdx.x * 16 + rr) << 8;
    double x[16];
    ConstF dummy; dummy.w = 0; dummy.wq = 0;
    const Tw15 t1 = ld_tw15(tw, base + 16 * l16, log_n);
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + l16 + 16 * i)] = ld(map, rref, base + l16 + 16 * i, mid, M);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[pad16(rr * 256 + 16 * l16 + i)];
    gs_radix16<1, false>(x, log_n, [&](int, int half, int i) { return mk_tw(t1.w[8 / half - 1 + i / (2 * half)], qinv); }, q, dummy, dummy);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[pad16(rr * 256 + 16 * l16 + i)] = reduce_sym(x[i], q, qinv);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm[pad16(rr * 256 + l16 + 16 * i)];
    gs_radix16<1, false>(x, log_n - 4, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, base + l16 + 16 * i, log_n)), qinv); }, q, dummy, dummy);
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, rref, base + l16 + 16 * i, reduce_sym(x[i], q, qinv), mid, M);
}

// pass A' : stages log_n-8 .. 1 (row strides 1..R/2); the final stage carries the scaling
// constants scale[row % rows_per_poly] = {c * N^-1, c * psi^-bitrev(1) * N^-1} (c = 1 when
// scale == nullptr).
template <int LOG_R, class StoreOp>
__global__ void __launch_bounds__(256, FHE_PASSA_MINB) ntt_inv_pass_a(DevTables T, RowMap map, LoadRaw ld, StoreOp st,
                                                      const ConstF* scale) {
    constexpr int R = 1 << LOG_R, G = R / 16, COLS = 256 / G, LEV1 = LOG_R - 4;
    FHE_SHARED double sm[4096];
    const int row = map.launch_row(blockIdx.y);
    const int mid = map.mod_id(row);
    if (mid < 0) return;
    const RowRef rref = map.ref(row);
    const Modulus M = T.mod[mid];
    const double q = M.qd, qinv = M.qinv;
    const int log_n = T.log_n;
    const double* tw = T.tw_inv + ((size_t)mid << log_n);
    const ConstF* fin = scale ? scale + 2 * (size_t)rref.j : T.inv_final + 2 * (size_t)mid;
    const ConstF fin0 = fin[0], fin1 = fin[1];
    const int tid = threadIdx.x, cc = tid % COLS, g = tid / COLS;
    const u32 c = blockIdx.x * COLS + cc;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, (u32)((16 * g + i) << 8) + c, mid, M);
    if (LEV1 > 0) {
        gs_radix16<1, false>(x, LOG_R, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, (u32)((16 * g + i) << 8) + c, log_n)), qinv); }, q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) sm[(16 * g + i) * COLS + cc] = reduce_sym(x[i], q, qinv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = sm[(g + G * i) * COLS + cc];
        gs_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1)), true>(x, LEV1, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, (u32)((g + G * i) << 8) + c, log_n)), qinv); }, q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, rref, (u32)((g + G * i) << 8) + c, x[i], mid, M);
    } else {
        gs_radix16<1, true>(x, LOG_R, [&](int S, int, int i) { return mk_tw(ld_d(tw + tw_index(S, (u32)((16 * g + i) << 8) + c, log_n)), qinv); }, q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, rref, (u32)((16 * g + i) << 8) + c, x[i], mid, M);
    }
}

// ------------------------------------------------------------------ host launchers
// `work` ([.., work_stride] layout) receives the lazy intermediate between the two passes; it
// may alias the destination, or the source for in-place use.
template <class LoadOp, class StoreOp>
inline void ntt_forward(const DevTables& T, const RowMap& map_in, int rows, LoadOp ld, u64* work,
                        long long work_stride, StoreOp st, cudaStream_t s) {
    const int log_r = T.log_n - 8;
    dim3 grid(1u << (log_r - 4), rows), block(256);
    StoreRaw sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadRaw lp; lp.src = work; lp.poly_stride = work_stride;
    RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
    switch (log_r) {
        case 4: fhe_launch(ntt_fwd_pass_a<4, LoadOp>, grid, block, 0, s, T, map, ld, sp); break;
        case 5: fhe_launch(ntt_fwd_pass_a<5, LoadOp>, grid, block, 0, s, T, map, ld, sp); break;
        case 6: fhe_launch(ntt_fwd_pass_a<6, LoadOp>, grid, block, 0, s, T, map, ld, sp); break;
        case 7: fhe_launch(ntt_fwd_pass_a<7, LoadOp>, grid, block, 0, s, T, map, ld, sp); break;
        default: fhe_launch(ntt_fwd_pass_a<8, LoadOp>, grid, block, 0, s, T, map, ld, sp); break;
    }
    fhe_launch(ntt_fwd_pass_b<StoreOp>, grid, block, 0, s, T, map, lp, st);
}

template <class LoadOp, class StoreOp>
inline void ntt_inverse(const DevTables& T, const RowMap& map_in, int rows, LoadOp ld, u64* work,
                        long long work_stride, StoreOp st, const ConstF* scale, cudaStream_t s) {
    const int log_r = T.log_n - 8;
    dim3 grid(1u << (log_r - 4), rows), block(256);
    StoreRaw sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadRaw lp; lp.src = work; lp.poly_stride = work_stride;
    RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
    fhe_launch(ntt_inv_pass_b<LoadOp>, grid, block, 0, s, T, map, ld, sp);
    switch (log_r) {
        case 4: fhe_launch(ntt_inv_pass_a<4, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 5: fhe_launch(ntt_inv_pass_a<5, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 6: fhe_launch(ntt_inv_pass_a<6, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 7: fhe_launch(ntt_inv_pass_a<7, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        default: fhe_launch(ntt_inv_pass_a<8, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
    }
}
