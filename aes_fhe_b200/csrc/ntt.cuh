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
#ifdef FHE_NTT_DIAG_NOMEM      /* measurement variant only (tools/ntt_ab.py): arithmetic and exchanges without global data traffic */
        return u64_to_f((u64)(idx * 2654435761u + row.j));
#else
        return u64_to_f(src[row_off(map, row, poly_stride) + idx]);
#endif
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
#ifdef FHE_NTT_DIAG_NOMEM
        if (v == -12345.0) dst[row_off(map, row, poly_stride) + idx] = f_to_u64(v);
#else
        dst[row_off(map, row, poly_stride) + idx] = f_to_u64(v);
#endif
    }
    // a StoreOp may read other operands in its epilogue: warm(idx) is called once per thread at the
    // start of the last pass with the first of 16 consecutive elements the thread's row will need
    FHE_D void warm(const RowMap&, RowRef, int) const {}
    // ... and fetch(idx) returns that operand ahead of the store loop (all 16 loads of a thread in flight at
    // once; inside operator() each load would have to wait for the previous store, the compiler cannot prove
    // that `out` and `in` do not alias); put() is operator() with the fetched value
    FHE_D double fetch(const RowMap&, RowRef, int) const { return 0.0; }
    FHE_D void put(const RowMap& map, RowRef row, int idx, double v, int mid, const Modulus& M, double) const {
        (*this)(map, row, idx, v, mid, M);
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

// ------------------------------------------------------------------ two-pass kernels
// Grid: (tiles, n_blocks, rows_per_poly) -- blockIdx.y is the poly block, blockIdx.z the limb slot,
// so no division is needed to find the row; launch order is limb-major (z outermost), which
// keeps the CTAs sharing one twiddle table co-resident (L2 reuse).  LOG_N, the tile shape and
// the stage are compile-time, indices are signed ints: after unrolling every access is
// `base register + immediate`.
//
// Twiddle index algebra (element e, stage S, LOG_N = LOG_R + 8): tw_index = 2^(S-1) + (e >> (LOG_N-S+1)).
//   column pass, first round  (S = 1..LEV1, registers hold rows g + G i):   2^(S-1) + (i >> (5-S))          (thread-independent)
//   column pass, second round (S = LEV1+t, rows 16 g + i):                  2^(S-1) + (g << (t-1)) + (i >> (5-t))
//   row pass (S = LOG_R+s, 256-element row `grow`, local element e):        2^(S-1) + (grow << (s-1)) + (e >> (9-s))
FHE_D int ntt_dig(const RowMap& map, int blk) { return map.skip_alpha > 0 ? blk % map.digits : 0; }

#ifndef FHE_EMU
FHE_D void ntt_sync_warp() { __syncwarp(); }
#else
inline void ntt_sync_warp() { __syncthreads(); }       // control flow is CTA-uniform: equivalent
#endif
// 16 contiguous doubles <-> registers with 16-byte shared-memory accesses
FHE_D void lds16(double (&x)[16], const double* p) {
#ifndef FHE_EMU
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 v = reinterpret_cast<const double2*>(p)[i]; x[2 * i] = v.x; x[2 * i + 1] = v.y; }
#else
    for (int i = 0; i < 16; ++i) x[i] = p[i];
#endif
}
FHE_D void sts16(double* p, const double (&x)[16]) {
#ifndef FHE_EMU
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<double2*>(p)[i] = make_double2(x[2 * i], x[2 * i + 1]);
#else
    for (int i = 0; i < 16; ++i) p[i] = x[i];
#endif
}
#define FHE_ROW_STRIDE 288     /* a 256-element row as 16 blocks of 16 (+2 pad): LDS.128, conflict-free both ways */

template <int LOG_R>
FHE_D double tw_col_r1(const double* tw, int S, int i) { return ld_d(tw + (1 << (S - 1)) + (i >> (5 - S))); }
template <int LOG_R>
FHE_D double tw_col_r2(const double* tw, int g, int S, int i) {
    constexpr int LEV1 = LOG_R - 4;
    const int t = S - LEV1;
    return ld_d(tw + (1 << (S - 1)) + (g << (t - 1)) + (i >> (5 - t)));
}
// the 15 twiddles of the last four stages for the 16 contiguous elements of block `blk16` of the limb
template <int LOG_N>
FHE_D Tw15 ld_tw15c(const double* tw, int blk16) {
    Tw15 t;
    constexpr int n16 = 1 << (LOG_N - 4);
    t.w[0] = ld_d(tw + n16 + blk16);
#ifndef FHE_EMU
    const double2 a = __ldg(reinterpret_cast<const double2*>(tw + 2 * n16) + blk16);
    t.w[1] = a.x; t.w[2] = a.y;
    const double2* p4 = reinterpret_cast<const double2*>(tw + 4 * n16) + 2 * blk16;
    const double2 b0 = __ldg(p4), b1 = __ldg(p4 + 1);
    t.w[3] = b0.x; t.w[4] = b0.y; t.w[5] = b1.x; t.w[6] = b1.y;
    const double2* p8 = reinterpret_cast<const double2*>(tw + 8 * n16) + 4 * blk16;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const double2 c = __ldg(p8 + j); t.w[7 + 2 * j] = c.x; t.w[8 + 2 * j] = c.y; }
#else
    for (int j = 0; j < 2; ++j) t.w[1 + j] = tw[2 * (n16 + blk16) + j];
    for (int j = 0; j < 4; ++j) t.w[3 + j] = tw[4 * (n16 + blk16) + j];
    for (int j = 0; j < 8; ++j) t.w[7 + j] = tw[8 * (n16 + blk16) + j];
#endif
    return t;
}

// ------------------------------------------------------------------ forward
template <int LOG_R, class LoadOp, class StoreOp>
FHE_D void fwd_pass_a_body(const DevTables& T, const RowMap& map, RowRef rref, int mid, int tile, const LoadOp& ld,
                           const StoreOp& st, double* sm) {
    constexpr int R = 1 << LOG_R, G = R / 16, COLS = 256 / G, LEV1 = LOG_R - 4, LOG_N = LOG_R + 8;
    const Modulus M = T.mod[mid];
    const double q = M.qd, qinv = M.qinv;
    const double* tw = T.tw_fwd + ((size_t)mid << LOG_N);
    const int tid = threadIdx.x, cc = tid % COLS, g = tid / COLS;
    const int c = tile * COLS + cc;
    double x[16];
    if (LEV1 > 0) {
        const int e0 = (g << 8) + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, e0 + i * (G << 8), mid, M);
        ct_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1))>(x, 1, [&](int S, int, int i) { return mk_tw(tw_col_r1<LOG_R>(tw, S, i), qinv); }, q);
#ifndef FHE_NTT_DIAG_NOSMEM     /* measurement variant only: the forward transform without its exchanges (wrong results) */
        double* smw = sm + g * COLS + cc;
#pragma unroll
        for (int i = 0; i < 16; ++i) smw[i * (G * COLS)] = x[i];
        __syncthreads();
        const double* smr = sm + 16 * g * COLS + cc;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = smr[i * COLS];
#endif
    } else {
        const int e0 = (g << 12) + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, e0 + (i << 8), mid, M);
    }
    ct_radix16<1>(x, LEV1 + 1, [&](int S, int, int i) { return mk_tw(tw_col_r2<LOG_R>(tw, g, S, i), qinv); }, q);
    const int o0 = (g << 12) + c;
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, rref, o0 + (i << 8), x[i], mid, M);
}
template <int LOG_R, class LoadOp>
__global__ void __launch_bounds__(256, FHE_PASSA_MINB) ntt_fwd_pass_a(DevTables T, RowMap map, LoadOp ld, StoreRaw st) {
    FHE_SHARED double sm[4096];
    RowRef rref; rref.blk = blockIdx.y; rref.j = blockIdx.z;
    const int mid = map.mod_id_of(rref.j, ntt_dig(map, rref.blk));
    if (mid < 0) return;
    fwd_pass_a_body<LOG_R>(T, map, rref, mid, blockIdx.x, ld, st, sm);
}

template <int LOG_N, class LoadOp, class StoreOp>
FHE_D void fwd_pass_b_body(const DevTables& T, const RowMap& map, RowRef rref, int mid, int tile, const LoadOp& ld,
                           const StoreOp& st, double* sm) {
    constexpr int LOG_R = LOG_N - 8, RS = FHE_ROW_STRIDE;
    const Modulus M = T.mod[mid];
    const double q = M.qd, qinv = M.qinv;
    const double* tw = T.tw_fwd + ((size_t)mid << LOG_N);
    const int tid = threadIdx.x, l16 = tid & 15, rr = tid >> 4;
    const int grow = tile * 16 + rr;                          // 256-element row of the limb
    const int rbase = (grow << 8) + l16;
    double* smS = sm + rr * RS + l16;                         // strided view    [18 * i]
    double* smC = sm + rr * RS + 18 * l16;                    // contiguous view [i]
    double x[16];
    st.warm(map, rref, (grow << 8) + 16 * l16);               // epilogue operands -> L2 while the butterflies run
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, rbase + 16 * i, mid, M);
    const Tw15 t2 = ld_tw15c<LOG_N>(tw, grow * 16 + l16);     // issued early, used in round 2
    {
        const double* t1 = tw + (1 << LOG_R) + grow, *t2p = tw + (2 << LOG_R) + 2 * grow,
                    *t3 = tw + (4 << LOG_R) + 4 * grow, *t4 = tw + (8 << LOG_R) + 8 * grow;
        ct_radix16<1>(x, 1, [&](int s, int, int i) {
            return mk_tw(ld_d(s == 1 ? t1 : s == 2 ? t2p + (i >> 3) : s == 3 ? t3 + (i >> 2) : t4 + (i >> 1)), qinv); }, q);
    }
#ifndef FHE_NTT_DIAG_NOSMEM
#pragma unroll
    for (int i = 0; i < 16; ++i) smS[18 * i] = x[i];
    ntt_sync_warp();                                          // a row is owned by 16 lanes of one warp
    lds16(x, smC);
#endif
    ct_radix16<1>(x, 5, [&](int, int half, int i) { return mk_tw(t2.w[8 / half - 1 + i / (2 * half)], qinv); }, q);
    ntt_sync_warp();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = reduce_canon(x[i], q, qinv);
#ifndef FHE_NTT_DIAG_NOSMEM
    sts16(smC, x);
#endif
    double ep[16];                                            // epilogue operands (zeros for a plain store)
#pragma unroll
    for (int i = 0; i < 16; ++i) ep[i] = st.fetch(map, rref, rbase + 16 * i);
    ntt_sync_warp();
#ifndef FHE_NTT_DIAG_NOSMEM
#pragma unroll
    for (int i = 0; i < 16; ++i) st.put(map, rref, rbase + 16 * i, smS[18 * i], mid, M, ep[i]);
#else
#pragma unroll
    for (int i = 0; i < 16; ++i) st.put(map, rref, rbase + 16 * i, x[i], mid, M, ep[i]);
    (void)smS; (void)smC;
#endif
}
template <int LOG_N, class StoreOp>
__global__ void __launch_bounds__(256, FHE_PASSB_MINB) ntt_fwd_pass_b(DevTables T, RowMap map, LoadRaw ld, StoreOp st) {
    FHE_SHARED __align__(16) double sm[16 * FHE_ROW_STRIDE];
    RowRef rref; rref.blk = blockIdx.y; rref.j = blockIdx.z;
    const int mid = map.mod_id_of(rref.j, ntt_dig(map, rref.blk));
    if (mid < 0) return;
    fwd_pass_b_body<LOG_N>(T, map, rref, mid, blockIdx.x, ld, st, sm);
}

// ------------------------------------------------------------------ inverse
// pass B' : stages LOG_N .. LOG_N-7 (strides 1..128) on 16 contiguous rows of 256.
template <int LOG_N, class LoadOp, class StoreOp>
FHE_D void inv_pass_b_body(const DevTables& T, const RowMap& map, RowRef rref, int mid, int tile, const LoadOp& ld,
                           const StoreOp& st, double* sm) {
    constexpr int LOG_R = LOG_N - 8, RS = FHE_ROW_STRIDE;
    const Modulus M = T.mod[mid];
    const double q = M.qd, qinv = M.qinv;
    const double* tw = T.tw_inv + ((size_t)mid << LOG_N);
    const int tid = threadIdx.x, l16 = tid & 15, rr = tid >> 4;
    const int grow = tile * 16 + rr;
    const int rbase = (grow << 8) + l16;
    double* smS = sm + rr * RS + l16;
    double* smC = sm + rr * RS + 18 * l16;
    double x[16];
    ConstF dummy; dummy.w = 0; dummy.wq = 0;
    const Tw15 t1 = ld_tw15c<LOG_N>(tw, grow * 16 + l16);
#pragma unroll
    for (int i = 0; i < 16; ++i) smS[18 * i] = ld(map, rref, rbase + 16 * i, mid, M);
    ntt_sync_warp();
    lds16(x, smC);
    gs_radix16<1, false>(x, 8, [&](int, int half, int i) { return mk_tw(t1.w[8 / half - 1 + i / (2 * half)], qinv); }, q, dummy, dummy);
    ntt_sync_warp();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = reduce_sym(x[i], q, qinv);
    sts16(smC, x);
    ntt_sync_warp();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = smS[18 * i];
    {
        const double* t1p = tw + (1 << LOG_R) + grow, *t2p = tw + (2 << LOG_R) + 2 * grow,
                    *t3 = tw + (4 << LOG_R) + 4 * grow, *t4 = tw + (8 << LOG_R) + 8 * grow;
        gs_radix16<1, false>(x, 4, [&](int s, int, int i) {
            return mk_tw(ld_d(s == 1 ? t1p : s == 2 ? t2p + (i >> 3) : s == 3 ? t3 + (i >> 2) : t4 + (i >> 1)), qinv); }, q, dummy, dummy);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) st(map, rref, rbase + 16 * i, reduce_sym(x[i], q, qinv), mid, M);
}
template <int LOG_N, class LoadOp>
__global__ void __launch_bounds__(256, FHE_PASSB_MINB) ntt_inv_pass_b(DevTables T, RowMap map, LoadOp ld, StoreRaw st) {
    FHE_SHARED __align__(16) double sm[16 * FHE_ROW_STRIDE];
    RowRef rref; rref.blk = blockIdx.y; rref.j = blockIdx.z;
    const int mid = map.mod_id_of(rref.j, ntt_dig(map, rref.blk));
    if (mid < 0) return;
    inv_pass_b_body<LOG_N>(T, map, rref, mid, blockIdx.x, ld, st, sm);
}

// pass A' : stages LOG_R .. 1 (row strides 1..R/2); the final stage carries the scaling
// constants scale[limb slot] = {c * N^-1, c * psi^-bitrev(1) * N^-1} (c = 1 when scale == nullptr).
template <int LOG_R, class LoadOp, class StoreOp>
FHE_D void inv_pass_a_body(const DevTables& T, const RowMap& map, RowRef rref, int mid, int tile, const LoadOp& ld,
                           const StoreOp& st, const ConstF* scale, double* sm) {
    constexpr int R = 1 << LOG_R, G = R / 16, COLS = 256 / G, LEV1 = LOG_R - 4, LOG_N = LOG_R + 8;
    const Modulus M = T.mod[mid];
    const double q = M.qd, qinv = M.qinv;
    const double* tw = T.tw_inv + ((size_t)mid << LOG_N);
    const ConstF* fin = scale ? scale + 2 * (size_t)rref.j : T.inv_final + 2 * (size_t)mid;
    const ConstF fin0 = fin[0], fin1 = fin[1];
    const int tid = threadIdx.x, cc = tid % COLS, g = tid / COLS;
    const int c = tile * COLS + cc;
    double x[16];
    const int e1 = (g << 12) + c;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ld(map, rref, e1 + (i << 8), mid, M);
    if (LEV1 > 0) {
        gs_radix16<1, false>(x, LOG_R, [&](int S, int, int i) { return mk_tw(tw_col_r2<LOG_R>(tw, g, S, i), qinv); }, q, fin0, fin1);
        double* smw = sm + 16 * g * COLS + cc;
#pragma unroll
        for (int i = 0; i < 16; ++i) smw[i * COLS] = reduce_sym(x[i], q, qinv);
        __syncthreads();
        const double* smr = sm + g * COLS + cc;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = smr[i * (G * COLS)];
        gs_radix16<(16 >> (LEV1 > 0 ? LEV1 : 1)), true>(x, LEV1, [&](int S, int, int i) { return mk_tw(tw_col_r1<LOG_R>(tw, S, i), qinv); }, q, fin0, fin1);
        const int e0 = (g << 8) + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, rref, e0 + i * (G << 8), x[i], mid, M);
    } else {
        gs_radix16<1, true>(x, LOG_R, [&](int S, int, int i) { return mk_tw(tw_col_r2<LOG_R>(tw, g, S, i), qinv); }, q, fin0, fin1);
#pragma unroll
        for (int i = 0; i < 16; ++i) st(map, rref, e1 + (i << 8), x[i], mid, M);
    }
}
template <int LOG_R, class StoreOp>
__global__ void __launch_bounds__(256, FHE_PASSA_MINB) ntt_inv_pass_a(DevTables T, RowMap map, LoadRaw ld, StoreOp st,
                                                      const ConstF* scale) {
    FHE_SHARED double sm[4096];
    RowRef rref; rref.blk = blockIdx.y; rref.j = blockIdx.z;
    const int mid = map.mod_id_of(rref.j, ntt_dig(map, rref.blk));
    if (mid < 0) return;
    inv_pass_a_body<LOG_R>(T, map, rref, mid, blockIdx.x, ld, st, scale, sm);
}

// ------------------------------------------------------------------ host launchers
// `work` ([.., work_stride] layout) receives the lazy intermediate between the two passes; it
// may alias the destination, or the source for in-place use.
template <class LoadOp, class StoreOp>
inline void ntt_forward(const DevTables& T, const RowMap& map_in, int rows, LoadOp ld, u64* work,
                        long long work_stride, StoreOp st, cudaStream_t s) {
    const int log_r = T.log_n - 8;
    RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
    dim3 grid(1u << (log_r - 4), map.n_blocks, map.rows_per_poly), block(256);
    StoreRaw sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadRaw lp; lp.src = work; lp.poly_stride = work_stride;
    switch (log_r) {
        case 4: fhe_launch(ntt_fwd_pass_a<4, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_fwd_pass_b<12, StoreOp>, grid, block, 0, s, T, map, lp, st); break;
        case 5: fhe_launch(ntt_fwd_pass_a<5, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_fwd_pass_b<13, StoreOp>, grid, block, 0, s, T, map, lp, st); break;
        case 6: fhe_launch(ntt_fwd_pass_a<6, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_fwd_pass_b<14, StoreOp>, grid, block, 0, s, T, map, lp, st); break;
        case 7: fhe_launch(ntt_fwd_pass_a<7, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_fwd_pass_b<15, StoreOp>, grid, block, 0, s, T, map, lp, st); break;
        default: fhe_launch(ntt_fwd_pass_a<8, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                 fhe_launch(ntt_fwd_pass_b<16, StoreOp>, grid, block, 0, s, T, map, lp, st); break;
    }
}

template <class LoadOp, class StoreOp>
inline void ntt_inverse(const DevTables& T, const RowMap& map_in, int rows, LoadOp ld, u64* work,
                        long long work_stride, StoreOp st, const ConstF* scale, cudaStream_t s) {
    const int log_r = T.log_n - 8;
    RowMap map = map_in; map.n_blocks = rows / map.rows_per_poly;
    dim3 grid(1u << (log_r - 4), map.n_blocks, map.rows_per_poly), block(256);
    StoreRaw sp; sp.dst = work; sp.poly_stride = work_stride;
    LoadRaw lp; lp.src = work; lp.poly_stride = work_stride;
    switch (log_r) {
        case 4: fhe_launch(ntt_inv_pass_b<12, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_inv_pass_a<4, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 5: fhe_launch(ntt_inv_pass_b<13, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_inv_pass_a<5, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 6: fhe_launch(ntt_inv_pass_b<14, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_inv_pass_a<6, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        case 7: fhe_launch(ntt_inv_pass_b<15, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                fhe_launch(ntt_inv_pass_a<7, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
        default: fhe_launch(ntt_inv_pass_b<16, LoadOp>, grid, block, 0, s, T, map, ld, sp);
                 fhe_launch(ntt_inv_pass_a<8, StoreOp>, grid, block, 0, s, T, map, lp, st, scale); break;
    }
}
