// kernels.cuh -- streaming (HBM-bound) kernels around the NTT: ciphertext arithmetic,
// Galois permutation, hybrid key-switch phases (base conversion, key inner product) and
// the fused load/store functors used by rescale and ModDown.
//
// Layout convention: a "poly block" is [rows_per_poly][N] u64, blocks are contiguous unless
// a stride is given.  One thread handles one coefficient index of one row (or of all rows
// when a contraction over limbs is involved); consecutive threads touch consecutive
// addresses, so every access is a full 128-byte line per half-warp.
#pragma once
#include "ntt_fused.cuh"

#define FHE_MAX_SRC 16      // max limbs in one base-conversion source basis (alpha, K)
#define FHE_MAX_DST 48      // max target limbs of one base conversion
#define FHE_MAX_LIMBS 48

FHE_D u64 ld_u64(const u64* p) {
#ifndef FHE_EMU
    return __ldg(p);
#else
    return *p;
#endif
}

#ifndef FHE_EMU
#define FHE_GRID_CONSTANT __grid_constant__
#else
#define FHE_GRID_CONSTANT
#endif

struct LimbConsts {          // per-limb-slot constants passed by value
    u64 a[FHE_MAX_LIMBS];
    u64 b[FHE_MAX_LIMBS];
};

// ---------------------------------------------------------------- elementwise
// op: 0 add, 1 sub, 2 mul.  Rows are [poly][batch][limb]; each operand has its own poly and
// batch strides (0 = broadcast), so plaintexts / keys are never replicated in memory.
struct Strides { long long poly, batch; };
template <int OP>
__global__ void __launch_bounds__(256) k_binary(DevTables T, RowMap map, int batch, u64* out, const u64* a, const u64* b,
                                                Strides so, Strides sa, Strides sb) {
    const int row = blockIdx.y;
    const int mid = map.mod_id(row);
    const Modulus M = T.mod[mid];
    // four elements 256 apart per thread, all eight loads issued before the first store (HBM-bound: bytes in flight)
    const u32 idx = blockIdx.x * 1024 + threadIdx.x;
    const int blk = row / map.rows_per_poly, j = row % map.rows_per_poly;
    const size_t poly = blk / batch, bt = blk % batch;
    const size_t lo = ((size_t)j << map.log_n) + idx;
    const u64* ap = a + poly * sa.poly + bt * sa.batch + lo;
    const u64* bp = b + poly * sb.poly + bt * sb.batch + lo;
    u64 x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = ld_u64(ap + 256 * i); y[i] = ld_u64(bp + 256 * i); }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u64 r;
        if (OP == 0) r = add_mod(x[i], y[i], M.q);
        else if (OP == 1) r = sub_mod(x[i], y[i], M.q);
        else r = f_to_u64(canon(mulmod_var(u64_to_f(x[i]), u64_to_f(y[i]), M.qd, M.qinv), M.qd));
        out[poly * so.poly + bt * so.batch + lo + 256 * i] = r;
    }
}

__global__ void __launch_bounds__(256) k_neg(DevTables T, RowMap map, u64* out, const u64* a) {
    const int row = blockIdx.y;
    const Modulus M = T.mod[map.mod_id(row)];
    const size_t o = ((size_t)row << map.log_n) + blockIdx.x * 256 + threadIdx.x;
    out[o] = neg_mod(a[o], M.q);
}

// d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1      rows: batch * nq limbs, poly stride ps
__global__ void __launch_bounds__(256) k_tensor(DevTables T, int nq, u64* d, const u64* a, const u64* b, long long ps) {
    const int j = blockIdx.y % nq;
    const Modulus M = T.mod[j];
    const size_t n = (size_t)1 << T.log_n;
    const size_t o = (size_t)blockIdx.y * n + blockIdx.x * 256 + threadIdx.x;
    const double q = M.qd, qi = M.qinv;
    const double a0 = u64_to_f(a[o]), a1 = u64_to_f(a[o + ps]), b0 = u64_to_f(b[o]), b1 = u64_to_f(b[o + ps]);
    d[o] = f_to_u64(canon(mulmod_var(a0, b0, q, qi), q));
    const double m = d_add(mulmod_var(a0, b1, q, qi), mulmod_var(a1, b0, q, qi));      // (-1.1q, 1.1q)
    d[o + ps] = f_to_u64(reduce_canon(m, q, qi));
    d[o + 2 * ps] = f_to_u64(canon(mulmod_var(a1, b1, q, qi), q));
}

// out = a * (first half of spectrum ? c.a[j] : c.b[j])   (complex constant, see encoding.py)
// ADD != 0: out = a + constant instead.
template <int ADD>
__global__ void __launch_bounds__(256) k_const(DevTables T, RowMap map, u64* out, const u64* a, LimbConsts c) {
    const int row = blockIdx.y;
    const int mid = map.mod_id(row);
    const Modulus M = T.mod[mid];
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const int j = row % map.rows_per_poly;
    const u64 k = (idx >> (map.log_n - 1)) ? c.b[j] : c.a[j];
    const size_t o = ((size_t)row << map.log_n) + idx;
    out[o] = ADD ? add_mod(a[o], k, M.q) : f_to_u64(canon(mulmod_var(u64_to_f(a[o]), u64_to_f(k), M.qd, M.qinv), M.qd));
}

// NTT-domain automorphism X -> X^g on bit-reversed spectra: out[p] = in[perm(p)]
// A thread moves four elements 256 apart (four gathers in flight before the first store): the kernel is pure
// memory traffic and one 8-byte load per thread left the memory system short of requests.   grid: (N/1024, rows)
__global__ void __launch_bounds__(256) k_automorphism(int log_n, u64* out, const u64* in, u64 g) {
    const size_t ro = (size_t)blockIdx.y << log_n;
    const u32 mask = (2u << log_n) - 1;
    const u32 p0 = blockIdx.x * 1024 + threadIdx.x;
    u64 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u32 k = brev32(p0 + 256 * i) >> (32 - log_n);
        const u32 kk = (u32)((g * (u64)(2 * k + 1)) & mask) >> 1;
        v[i] = ld_u64(in + ro + (brev32(kk) >> (32 - log_n)));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[ro + p0 + 256 * i] = v[i];
}

// signed coefficients [batch][N] -> residues of every row's modulus, rows [batch][limbs]
__global__ void __launch_bounds__(256) k_from_i64(DevTables T, RowMap map, u64* out, const long long* coeffs) {
    const int row = blockIdx.y;
    const Modulus M = T.mod[map.mod_id(row)];
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const long long v = coeffs[((size_t)(row / map.rows_per_poly) << map.log_n) + idx];
    u64 r;
    if (v >= 0) r = reduce_u64((u64)v, M);
    else r = neg_mod(reduce_u64((u64)(-v), M), M.q);
    out[((size_t)row << map.log_n) + idx] = r;
}

// centred CRT of limbs (0,1) (or limb 0 alone) to double; x: [batch][limbs][N], out [batch][N]
__global__ void __launch_bounds__(256) k_crt_centered(DevTables T, double* out, const u64* x, int limbs, u64 q0inv_mod_q1) {
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const Modulus M0 = T.mod[0];
    x += ((size_t)blockIdx.y * limbs) << T.log_n;
    out += (size_t)blockIdx.y << T.log_n;
    const u64 x0 = x[idx];
    if (limbs < 2) {
        out[idx] = x0 > (M0.q >> 1) ? -(double)(M0.q - x0) : (double)x0;
        return;
    }
    const Modulus M1 = T.mod[1];
    const u64 x1 = x[((size_t)1 << T.log_n) + idx];
    const u64 t = mul_mod(sub_mod(x1, reduce_u64(x0, M1), M1.q), q0inv_mod_q1, M1);
    // v = x0 + q0 * t  in [0, q0 q1);  compare with Q/2 and negate in 128 bits
    u128t v = mul_wide(M0.q, t);
    v.lo += x0; v.hi += (v.lo < x0);
    u128t Q = mul_wide(M0.q, M1.q);
    u128t h; h.lo = (Q.lo >> 1) | (Q.hi << 63); h.hi = Q.hi >> 1;
    const bool neg = v.hi > h.hi || (v.hi == h.hi && v.lo > h.lo);
    if (neg) { u128t w; w.lo = Q.lo - v.lo; w.hi = Q.hi - v.hi - (Q.lo < v.lo); v = w; }
    const double d = (double)v.hi * 18446744073709551616.0 + (double)v.lo;
    out[idx] = neg ? -d : d;
}

// ---------------------------------------------------------------- LUT inner sums (K7)
// out_m = sum_t c[m][t] (.) in_t  (+ c0[m] on polynomial 0),   m < M, t < T <= 16
// The reference encodes every LUT coefficient as a full-slot constant plaintext and does one
// multiply + rescale + add per term (xor_service.py:283-285, sbox_service.py:124-136).  A
// constant a+bi is the 2-term polynomial a + b X^(N/2), whose NTT image takes one value on the
// first half of the bit-reversed spectrum and another on the second half -- so the whole
// inner sum is an elementwise weighted sum.  Each input ciphertext is read ONCE and all M
// outputs are produced from registers; no rescale happens here (the constants already carry
// the scale that makes the later product land on the level's scale).
#define FHE_LC_MAX_T 16
struct LinCombIn {
    const u64* ptr[FHE_LC_MAX_T];        // in_t : [2][batch][nq_t][N]
    long long poly_stride[FHE_LC_MAX_T];
    long long batch_stride[FHE_LC_MAX_T];
};
// consts: [M][T][nq][2] ConstF (first / second half of the spectrum); c0: [M][nq][2] u64 or null
// The sums are a warp-level matrix product on the FP64 tensor cores: rows = 8 coefficients,
// K = the T inputs, columns = 8 outputs.  Exact Karatsuba dot products (modarith.cuh, split23 / dot3_finish):
// inputs and constants are split at 2^23 into (h, l, s = h + l); the three sums  sum h h', sum s s', sum l l'
// are DMMA accumulations of integers below 2^51, so every FMA inside the tensor core is exact, and each output
// is reduced once.  Three FMAs per term instead of the seven of mulmod_const + add, no operand fetch per FMA:
// the constants of an 8-output tile are a B fragment in registers (one conflict-free LDS.64 per 8x8x4 block).
// The scalar form of this kernel (constants broadcast from shared memory into DFMAs) ran at 53 % of the FP64
// peak because a broadcast LDS delivers 4 bytes per cycle per SM; DMMA runs at the same 64 FMA/clk/SM as DFMA
// (tools/ubench/dmma.cu) but needs no operand traffic: 5.70 -> 3.56 ms for the 22 x 11 sums of SubBytes at
// batch 16 (profiles/r01_kernel_breakdown_*.md).  The same formulation for the base conversion was measured
// and not kept: k_bconv already runs at 89 % of the FP64 peak and the reduction per target eats the gain.
//   consts staged per CTA as three shared-memory planes [ceil8(M)][STRIDE];  grid: (N / 512, 2 * batch * nq)
#define FHE_LCM_CHUNK 512
#define FHE_LCM_STRIDE(KT) (4 * (KT) + ((KT) % 2 == 0 ? 4 : 0))
#define FHE_LCM_SMEM(M, KT) ((size_t)3 * (((M) + 7) & ~7) * FHE_LCM_STRIDE(KT) * sizeof(double))
#ifndef FHE_LCM_MT
#define FHE_LCM_MT 2          // row blocks (of 8 coefficients) per B fragment load
#define FHE_LCM_MINB 3        // measured on B200 (22 x 11 sums, batch 16): (MT, CTAs/SM) = (2,2) 3.56 ms, (1,3) 3.52,
                              // (1,4) 3.53, (2,3) 3.33, (4,1) 4.45 -- the kernel is bound by dependent-issue latency
#endif
template <int KT>
__global__ void __launch_bounds__(256, FHE_LCM_MINB) k_lincomb_mma(DevTables Tb, int nq, int batch, int M, int T, LinCombIn in,
                                                        const ConstF* consts, const u64* c0, u64* out) {
    constexpr int TP = 4 * KT, STRIDE = FHE_LCM_STRIDE(KT), MT = FHE_LCM_MT;
    FHE_DYN_SHARED(double, s_c);
    const int row = blockIdx.y;
    const int j = row % nq, b = (row / nq) % batch, poly = row / (nq * batch);
    const Modulus Mo = Tb.mod[j];
    const double q = Mo.qd, qi = Mo.qinv;
    const int half = (blockIdx.x * FHE_LCM_CHUNK) >> (Tb.log_n - 1);      // uniform over the CTA
    const int MP = (M + 7) & ~7, plane = MP * STRIDE;
    for (int i = threadIdx.x; i < MP * TP; i += 256) {
        const int m = i / TP, t = i % TP;
        const Split3 c = split23(m < M && t < T ? consts[(((size_t)m * T + t) * nq + j) * 2 + half].w : 0.0);
        s_c[m * STRIDE + t] = c.h; s_c[plane + m * STRIDE + t] = c.l; s_c[2 * plane + m * STRIDE + t] = c.s;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const u64* p[KT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        const int t = 4 * kt + tig;
        p[kt] = t < T ? in.ptr[t] + (size_t)poly * in.poly_stride[t] + (size_t)b * in.batch_stride[t] + ((size_t)j << Tb.log_n)
                      : nullptr;
    }
    __syncthreads();
    const size_t out_ct = (size_t)2 * batch * nq << Tb.log_n;
    u64* o = out + (((size_t)poly * batch + b) * nq << Tb.log_n) + ((size_t)j << Tb.log_n);
    const bool add_c0 = c0 != nullptr && poly == 0;
    const u32 base = blockIdx.x * FHE_LCM_CHUNK + warp * (FHE_LCM_CHUNK / 8);
    for (int it = 0; it < FHE_LCM_CHUNK / 8 / (8 * MT); ++it) {
        Split3 a[MT][KT];
        u32 ci[MT];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
            ci[mi] = base + (it * MT + mi) * 8 + gid;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) a[mi][kt] = split23(p[kt] != nullptr ? u64_to_f(p[kt][ci[mi]]) : 0.0);
        }
        for (int n = 0; n < MP / 8; ++n) {
            const int bo = (8 * n + gid) * STRIDE + tig;
            double acc[MT][3][2];
#pragma unroll
            for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                for (int u = 0; u < 3; ++u) acc[mi][u][0] = acc[mi][u][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const double bh = s_c[bo + 4 * kt], bl = s_c[plane + bo + 4 * kt], bs = s_c[2 * plane + bo + 4 * kt];
#pragma unroll
                for (int mi = 0; mi < MT; ++mi) {
                    dmma884(acc[mi][0][0], acc[mi][0][1], a[mi][kt].h, bh);
                    dmma884(acc[mi][1][0], acc[mi][1][1], a[mi][kt].s, bs);
                    dmma884(acc[mi][2][0], acc[mi][2][1], a[mi][kt].l, bl);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = 8 * n + 2 * tig + e;
                if (m < M) {
                    const double k = add_c0 ? u64_to_f(c0[((size_t)m * nq + j) * 2 + half]) : 0.0;
#pragma unroll
                    for (int mi = 0; mi < MT; ++mi) {
                        double r = dot3_finish(acc[mi][0][e], acc[mi][1][e], acc[mi][2][e], q, qi);
                        if (add_c0) { r = d_add(r, k); r = r >= q ? d_add(r, -q) : r; }
                        o[(size_t)m * out_ct + ci[mi]] = f_to_u64(r);
                    }
                }
            }
        }
    }
}

// One or two outputs (M <= 2: the constant-weighted sums of the Chebyshev leaves of EvalMod, level adjustments): the
// 8-column tensor-core tile would be seven eighths padding, so these run as plain FP64 code -- mulmod_const per term,
// constants of the (limb, half) in shared memory, two coefficients 256 apart per thread.  HBM-bound.
//   grid: (N / 512, 2 * batch * nq)
template <int T_MAX>
__global__ void __launch_bounds__(256) k_lincomb_few(DevTables Tb, int nq, int batch, int M, int T, LinCombIn in,
                                                     const ConstF* consts, const u64* c0, u64* out) {
    FHE_SHARED ConstF s_c[2 * FHE_LC_MAX_T];                       // [M <= 2][T]
    const int row = blockIdx.y;
    const int j = row % nq, b = (row / nq) % batch, poly = row / (nq * batch);
    const Modulus Mo = Tb.mod[j];
    const double q = Mo.qd, qi = Mo.qinv;
    const u32 idx = blockIdx.x * 512 + threadIdx.x;
    const int half = (blockIdx.x * 512) >> (Tb.log_n - 1);        // uniform over the CTA
    for (int i = threadIdx.x; i < M * T; i += 256) s_c[i] = consts[((size_t)i * nq + j) * 2 + half];
    const size_t lo = ((size_t)j << Tb.log_n) + idx;
    double x0[T_MAX], x1[T_MAX];
#pragma unroll
    for (int t = 0; t < T_MAX; ++t) {
        if (t < T) {
            const u64* p = in.ptr[t] + (size_t)poly * in.poly_stride[t] + (size_t)b * in.batch_stride[t] + lo;
            x0[t] = u64_to_f(ld_u64(p)); x1[t] = u64_to_f(ld_u64(p + 256));
        } else { x0[t] = 0.0; x1[t] = 0.0; }
    }
    __syncthreads();
    const size_t out_ct = (size_t)2 * batch * nq << Tb.log_n;
    u64* o = out + (((size_t)poly * batch + b) * nq << Tb.log_n) + lo;
    for (int m = 0; m < M; ++m) {
        const ConstF* cm = s_c + m * T;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int t = 0; t < T_MAX; ++t)
            if (t < T) {
                const ConstF c = cm[t];
                a0 = d_add(a0, mulmod_const(x0[t], c, q));
                a1 = d_add(a1, mulmod_const(x1[t], c, q));
            }
        double r0 = reduce_canon(a0, q, qi), r1 = reduce_canon(a1, q, qi);
        if (c0 != nullptr && poly == 0) {
            const double k = u64_to_f(c0[((size_t)m * nq + j) * 2 + half]);
            r0 = d_add(r0, k); r0 = r0 >= q ? d_add(r0, -q) : r0;
            r1 = d_add(r1, k); r1 = r1 >= q ? d_add(r1, -q) : r1;
        }
        o[(size_t)m * out_ct] = f_to_u64(r0);
        o[(size_t)m * out_ct + 256] = f_to_u64(r1);
    }
}

// acc3 (+)= a (x) b  : accumulate the three tensor components of G products in one pass.
//   a_g : [2][batch][nq_a][N]  (first nq limbs used),  b : G x [2][batch][nq][N] contiguous
//   acc : [3][batch][nq][N]    grid: (N/256, batch * nq)
struct TensorAccIn {
    const u64* a[FHE_LC_MAX_T];
    long long a_poly_stride[FHE_LC_MAX_T];
    long long a_batch_stride[FHE_LC_MAX_T];
};
// acc_ps: polynomial stride of acc in words (the accumulator may be a batch slice of a larger [3][B][nq][N] tensor);
// init (optional, [2][batch][nq][N]): a 2-polynomial term the sum starts from (the constant-outer row of a LUT)
__global__ void __launch_bounds__(256) k_tensor_acc(DevTables Tb, int nq, int batch, int G, TensorAccIn in,
                                                    const u64* bsrc, int b_batch, u64* acc, int accumulate,
                                                    long long acc_ps, const u64* init) {
    const int row = blockIdx.y;
    const int j = row % nq, b = row / nq;
    const Modulus Mo = Tb.mod[j];
    const double q = Mo.qd, qi = Mo.qinv;
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const size_t lo = ((size_t)j << Tb.log_n) + idx;
    const size_t ps = (size_t)acc_ps;                            // poly stride of acc
    const size_t bo = ((size_t)b * nq << Tb.log_n) + lo;
    const size_t bps = (size_t)b_batch * nq << Tb.log_n;         // poly stride of b_g (b_batch = batch or 1)
    const size_t bbo = ((size_t)(b % b_batch) * nq << Tb.log_n) + lo;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    if (accumulate) { d0 = u64_to_f(acc[bo]); d1 = u64_to_f(acc[bo + ps]); d2 = u64_to_f(acc[bo + 2 * ps]); }
    if (init) { d0 = d_add(d0, u64_to_f(init[bo])); d1 = d_add(d1, u64_to_f(init[bo + ((size_t)batch * nq << Tb.log_n)])); }
    for (int g = 0; g < G; ++g) {
        const u64* ap = in.a[g] + (size_t)b * in.a_batch_stride[g] + lo;
        const double a0 = u64_to_f(ap[0]), a1 = u64_to_f(ap[in.a_poly_stride[g]]);
        const u64* bp = bsrc + (size_t)g * 2 * bps + bbo;
        const double b0 = u64_to_f(bp[0]), b1 = u64_to_f(bp[bps]);
        d0 = d_add(d0, mulmod_var(a0, b0, q, qi));
        d1 = d_add(d1, d_add(mulmod_var(a0, b1, q, qi), mulmod_var(a1, b0, q, qi)));
        d2 = d_add(d2, mulmod_var(a1, b1, q, qi));
    }
    acc[bo] = f_to_u64(reduce_canon(d0, q, qi));
    acc[bo + ps] = f_to_u64(reduce_canon(d1, q, qi));
    acc[bo + 2 * ps] = f_to_u64(reduce_canon(d2, q, qi));
}

// out (+)= sum_t a_t (.) p_t : the rotate-mask-add / diagonal-matrix pattern (shiftrows_service.py:41-50
// and the BSGS linear transforms of bootstrapping) in ONE pass.  p_t are plaintexts [nq][N] shared by
// both polynomials and the whole batch: a thread owns (limb, index), keeps its T plaintext words in
// registers and walks (poly, batch) with all T ciphertext loads of an iteration in flight.
//   a_t : [2][batch][nq_t >= nq][N],  out : [2][batch][nq][N]      grid: (N/256, nq)
struct PlainSumIn {
    const u64* a[FHE_LC_MAX_T];
    long long a_poly_stride[FHE_LC_MAX_T];
    long long a_batch_stride[FHE_LC_MAX_T];
    const u64* p[FHE_LC_MAX_T];
};
template <int T_MAX>
__global__ void __launch_bounds__(256, 2) k_mul_plain_sum(DevTables Tb, int nq, int batch, int T, PlainSumIn in, u64* out,
                                                           int accumulate) {
    const int j = blockIdx.y;
    const Modulus Mo = Tb.mod[j];
    const double q = Mo.qd, qi = Mo.qinv;
    const size_t lo = ((size_t)j << Tb.log_n) + blockIdx.x * 256 + threadIdx.x;
    double p[T_MAX];
#pragma unroll
    for (int t = 0; t < T_MAX; ++t) p[t] = t < T ? u64_to_f(ld_u64(in.p[t] + lo)) : 0.0;
    const size_t ops = (size_t)batch * nq << Tb.log_n, obs = (size_t)nq << Tb.log_n;
    for (int poly = 0; poly < 2; ++poly)
        for (int b = 0; b < batch; ++b) {
            u64 a[T_MAX];
#pragma unroll
            for (int t = 0; t < T_MAX; ++t)
                if (t < T) a[t] = ld_u64(in.a[t] + (size_t)poly * in.a_poly_stride[t] + (size_t)b * in.a_batch_stride[t] + lo);
            u64* o = out + (size_t)poly * ops + (size_t)b * obs + lo;
            double acc = accumulate ? u64_to_f(*o) : 0.0;
#pragma unroll
            for (int t = 0; t < T_MAX; ++t)
                if (t < T) acc = d_add(acc, mulmod_var(u64_to_f(a[t]), p[t], q, qi));
            *o = f_to_u64(reduce_canon(acc, q, qi));
        }
}

// out_g = sum_t a_t (.) p_{g,t} for g < G in ONE pass over the ciphertexts: all the diagonal sums of a baby-step/
// giant-step linear transform (one per giant step, same baby rotations a_t) -- every a_t is read once instead of G
// times (once per sweep of GC = 32 / T_MAX output rows, whose plaintext words a thread keeps in registers).
// p_{g,t} == nullptr: that diagonal is absent.   out : [G][2][batch][nq][N]      grid: (N/256, nq)
#define FHE_PM_MAX_G 8
struct PlainMultiIn {
    const u64* a[FHE_LC_MAX_T];
    long long a_poly_stride[FHE_LC_MAX_T];
    long long a_batch_stride[FHE_LC_MAX_T];
    const u64* p[FHE_PM_MAX_G][FHE_LC_MAX_T];
};
template <int T_MAX>
__global__ void __launch_bounds__(256, 2) k_mul_plain_multi(DevTables Tb, int nq, int batch, int T, int G, PlainMultiIn in,
                                                             u64* out) {
    constexpr int GC = 32 / T_MAX;                              // output rows per sweep: GC * T_MAX plaintext words in registers
    const int j = blockIdx.y;
    const Modulus Mo = Tb.mod[j];
    const double q = Mo.qd, qi = Mo.qinv;
    const size_t lo = ((size_t)j << Tb.log_n) + blockIdx.x * 256 + threadIdx.x;
    const size_t ops = (size_t)batch * nq << Tb.log_n, obs = (size_t)nq << Tb.log_n;
    for (int g0 = 0; g0 < G; g0 += GC) {
        double p[GC][T_MAX];
#pragma unroll
        for (int c = 0; c < GC; ++c)
#pragma unroll
            for (int t = 0; t < T_MAX; ++t) {
                const u64* pp = (g0 + c < G && t < T) ? in.p[g0 + c][t] : nullptr;
                p[c][t] = pp != nullptr ? u64_to_f(ld_u64(pp + lo)) : 0.0;
            }
        for (int poly = 0; poly < 2; ++poly)
            for (int b = 0; b < batch; ++b) {
                double a[T_MAX];
#pragma unroll
                for (int t = 0; t < T_MAX; ++t)
                    a[t] = t < T ? u64_to_f(ld_u64(in.a[t] + (size_t)poly * in.a_poly_stride[t] + (size_t)b * in.a_batch_stride[t] + lo))
                                 : 0.0;
#pragma unroll
                for (int c = 0; c < GC; ++c) {
                    if (g0 + c < G) {
                        double acc = 0.0;
#pragma unroll
                        for (int t = 0; t < T_MAX; ++t)
                            if (t < T) acc = d_add(acc, mulmod_var(a[t], p[c][t], q, qi));      // absent term: p = 0
                        out[(size_t)(g0 + c) * 2 * ops + (size_t)poly * ops + (size_t)b * obs + lo] =
                            f_to_u64(reduce_canon(acc, q, qi));
                    }
                }
            }
    }
}

// ---------------------------------------------------------------- double-hoisted baby steps of a BSGS transform
// All baby-step rotations of a baby-step/giant-step linear transform, their plaintext products and the sums over
// the baby steps, in ONE pass and entirely in the extended basis Q u P ("double hoisting": no ModDown per baby
// rotation, one per giant step):
//
//     out_g = sum_bb  pt[g][bb] (.) Z_bb,      Z_bb = < sigma_bb(ModUp(c1)), key_bb >  +  P (sigma_bb(c0), 0)
//
// sigma_bb is the Galois automorphism of baby step bb; on the bit-reversed spectrum it is the gather x[perm(p)]
// (k_automorphism), and it commutes with the base extension bit for bit, so ONE ModUp output `ext` serves every
// baby step.  The keys are read at p (coalesced), ext / c0 are gathered at perm(p): the automorphism maps every
// aligned block of 2^k spectrum indices onto an aligned block, so the 32 gathers of a warp cover exactly two
// 128-byte lines, and all gathers of one limb (batch x beta rows, 24 MB at batch 16) hit in L2 after the first
// baby step.  Neither the rotated ciphertexts nor the products ever exist in memory: HBM sees ext once, the keys,
// the plaintexts and the G output sums.
//   key[bb]  : [dnum][2][n_q + n_p][N] switching key of baby step bb; nullptr = the identity (Z = P (c0, c1))
//   pt[g][bb]: [nq + n_p][N] plaintext in the extended basis; nullptr = absent diagonal
//   ext      : [batch][beta][nq + n_p][N] (own-digit rows are read from c1);  ct: [2][batch][ct_nq][N]
//   out      : [G][2][batch][nq + n_p][N]
// grid: (N/256 * chunks, nq + n_p), chunks = ceil(batch / BB): the batch chunks of one (tile, limb) are neighbours in
// launch order, so the keys and plaintexts they share are read from HBM once and from L2 afterwards
#define FHE_BSGS_MAX_BABY 16
#define FHE_BSGS_MAX_G 4
struct BsgsIn {
    const u64* key[FHE_BSGS_MAX_BABY];
    u64 galois[FHE_BSGS_MAX_BABY];
    const u64* pt[FHE_BSGS_MAX_G][FHE_BSGS_MAX_BABY];
};
// BETA (digits), GN (giant steps) and BB (ciphertexts per thread) are compile-time and the batch tail is handled by
// clamping, so the body has no data-dependent branch inside a baby step: the 14 + 8 independent modular products of
// every ciphertext of a chunk are one basic block the scheduler can interleave (a first version with runtime
// `j < beta` / `g < G` tests compiled to ~7-instruction blocks and ran at 25 % of the FP64 pipe).  All row addresses
// are formed once per thread; a baby step only adds its gather index.
// SMACC: the GN x 2 x BB lazy accumulators of a thread live in shared memory ([slot][thread], conflict-free) instead
// of registers.
template <int BETA, int GN, int BB, bool SMACC>
__global__ void __launch_bounds__(256, (SMACC ? 3 : 2)) k_bsgs_inner(DevTables T, int nq, int alpha, int batch, int nb,
                                                                  BsgsIn in, const u64* ext, const u64* ct, int ct_nq,
                                                                  const ConstF* p_mod_q, u64* out, int accumulate) {
    const int t = blockIdx.y;
    const int ne = nq + T.n_p;
    const int id = t < nq ? t : T.n_q + (t - nq);
    const Modulus M = T.mod[id];
    const double q = M.qd, qi = M.qinv;
    const int log_n = T.log_n;
    const int chunks = (batch + BB - 1) / BB;
    const u32 p = (blockIdx.x / chunks) * 256 + threadIdx.x;
    const int b0 = (blockIdx.x % chunks) * BB;
    const size_t tot = (size_t)(T.n_q + T.n_p);
    const int own = t < nq ? t / alpha : -1;
    const bool qlimb = t < nq;
    ConstF pc; pc.w = 0; pc.wq = 0;
    if (qlimb) pc = p_mod_q[t];
    const u32 mask = (2u << log_n) - 1;
    const u32 kbr = brev32(p) >> (32 - log_n);
    const size_t c1_off = ((size_t)batch * ct_nq) << log_n;             // polynomial 1 of ct
    // row bases, fixed for the whole kernel (a clamped tail ciphertext recomputes the last one and is not stored)
    const u64* erow[BB][BETA];
    const u64* crow[BB];
#pragma unroll
    for (int u = 0; u < BB; ++u) {
        const int b = min(b0 + u, batch - 1);
        crow[u] = ct + ((((size_t)b * ct_nq) + (qlimb ? t : 0)) << log_n);
#pragma unroll
        for (int j = 0; j < BETA; ++j)
            erow[u][j] = j == own ? crow[u] + c1_off : ext + ((((size_t)b * BETA + j) * ne + t) << log_n);
    }
    const size_t key_row = ((size_t)id << log_n) + p, key_poly = tot << log_n, pt_row = ((size_t)t << log_n) + p;
    double acc[SMACC ? 1 : GN][2][SMACC ? 1 : BB];
    FHE_DYN_SHARED(double, sacc);                                       // SMACC: [GN * 2 * BB][256]
#define FHE_ACC(g, pl, u) (*(SMACC ? &sacc[(((g) * 2 + (pl)) * BB + (u)) * 256 + threadIdx.x] : &acc[SMACC ? 0 : (g)][pl][SMACC ? 0 : (u)]))
#pragma unroll
    for (int g = 0; g < GN; ++g)
#pragma unroll
        for (int u = 0; u < BB; ++u) { FHE_ACC(g, 0, u) = 0.0; FHE_ACC(g, 1, u) = 0.0; }
    for (int bb = 0; bb < nb; ++bb) {
        double pv[GN];
#pragma unroll
        for (int g = 0; g < GN; ++g) {
            const u64* pp = in.pt[g][bb];
            pv[g] = pp != nullptr ? u64_to_f(ld_u64(pp + pt_row)) : 0.0;
        }
        const u64* key = in.key[bb];
        double z0[BB], z1[BB];
        if (key != nullptr) {
            double k0[BETA], k1[BETA];
#pragma unroll
            for (int j = 0; j < BETA; ++j) {
                const u64* kp = key + (size_t)j * 2 * key_poly + key_row;
                k0[j] = u64_to_f(ld_u64(kp)); k1[j] = u64_to_f(ld_u64(kp + key_poly));
            }
            const u32 kk = (u32)((in.galois[bb] * (u64)(2 * kbr + 1)) & mask) >> 1;
            const u32 pp = brev32(kk) >> (32 - log_n);
            u64 e[BB][BETA], l0[BB];
#pragma unroll
            for (int u = 0; u < BB; ++u) {
#pragma unroll
                for (int j = 0; j < BETA; ++j) e[u][j] = ld_u64(erow[u][j] + pp);
                l0[u] = ld_u64(crow[u] + pp);
            }
#pragma unroll
            for (int u = 0; u < BB; ++u) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int j = 0; j < BETA; ++j) {
                    const double ev = u64_to_f(e[u][j]);
                    a0 = d_add(a0, mulmod_var(ev, k0[j], q, qi));
                    a1 = d_add(a1, mulmod_var(ev, k1[j], q, qi));
                }
                z0[u] = d_add(a0, mulmod_const(u64_to_f(l0[u]), pc, q));      // pc = 0 on the special limbs
                z1[u] = a1;
            }
        } else {
#pragma unroll
            for (int u = 0; u < BB; ++u) {
                z0[u] = mulmod_const(u64_to_f(ld_u64(crow[u] + p)), pc, q);
                z1[u] = mulmod_const(u64_to_f(ld_u64(crow[u] + c1_off + p)), pc, q);
            }
        }
        // |z| <= (BETA + 1) 0.51 q < 2^47: a valid mulmod_var operand; the sums over the baby steps stay lazy
#pragma unroll
        for (int u = 0; u < BB; ++u)
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                FHE_ACC(g, 0, u) = d_add(FHE_ACC(g, 0, u), mulmod_var(z0[u], pv[g], q, qi));
                FHE_ACC(g, 1, u) = d_add(FHE_ACC(g, 1, u), mulmod_var(z1[u], pv[g], q, qi));
            }
    }
    const size_t ops = ((size_t)batch * ne) << log_n;
#pragma unroll
    for (int g = 0; g < GN; ++g)
#pragma unroll
        for (int u = 0; u < BB; ++u) {
            const int b = b0 + u;
            if (b < batch) {
                u64* o = out + (size_t)g * 2 * ops + ((((size_t)b * ne) + t) << log_n) + p;
                double r0 = FHE_ACC(g, 0, u), r1 = FHE_ACC(g, 1, u);
                if (accumulate) { r0 = d_add(r0, u64_to_f(o[0])); r1 = d_add(r1, u64_to_f(o[ops])); }    // a second pass of baby steps
                o[0] = f_to_u64(reduce_canon(r0, q, qi));
                o[ops] = f_to_u64(reduce_canon(r1, q, qi));
            }
        }
#undef FHE_ACC
}

// ---------------------------------------------------------------- base conversion
// One table per source basis.  Output value for target t:
//     sum_k yc_k * f[k][t]   (mod m_t),     yc_k = centred representative of y_k mod q_k
// where y_k (already multiplied by (Q/q_k)^-1 mod q_k by the preceding iNTT) is the k-th
// source row.  Using the centred digit (-q_k/2, q_k/2] is the same as subtracting
// #{k : y_k > q_k/2} * Q, and makes the conversion error zero-mean (see oracle/refmod.cpp).
struct BConvTable {
    int ns, nt;
    int src_mod[FHE_MAX_SRC];
    int src_slot[FHE_MAX_SRC];      // row slot of source k inside the source block
    int dst_mod[FHE_MAX_DST];
    int dst_slot[FHE_MAX_DST];      // row slot of target t inside the destination group
    ConstF f[FHE_MAX_DST][FHE_MAX_SRC];   // (Q/q_k) mod m_t prepared for mulmod_const
};

// grid: (N/256, groups).  group g uses table g % n_tables and the source block g / n_tables
// (ModUp: groups = batch x digits, all digits of one ciphertext read the same source block;
//  ModDown: one table, groups = polys).  NS_MAX bounds every table's ns.  The constants of the
// table are staged in shared memory once per CTA.
// Every thread converts TWO coefficients (idx and idx + N/2): each constant fetched from shared
// memory feeds two modular products, which halves the non-FP64 instructions per FP64 instruction
// (the kernel is FP64-pipe bound and was co-limited by issue slots).   grid: (N/512, groups)
// (forcing 4 CTAs/SM with 64 registers changes nothing: 245.8 against 248.1 us for 12-limb digits)
// Every table of one launch has EXACTLY NS_MAX sources (the host launches the digits of a ModUp one by one: group
// g = blockIdx.y * g_step + g_first), so the source loops carry no `k < ns` test: with the test the unrolled
// accumulation compiled into one basic block per term and the FP64 pipe idled between them.
template <int NS_MAX>
__global__ void __launch_bounds__(256) k_bconv(DevTables T, const BConvTable* tables, int n_tables,
                                               u64* dst, long long dst_group_stride,
                                               const u64* src, long long src_block_stride, int g_first, int g_step) {
    FHE_SHARED ConstF sf[FHE_MAX_DST * NS_MAX];
    FHE_SHARED double sq[FHE_MAX_DST], sqi[FHE_MAX_DST];
    FHE_SHARED int sslot[FHE_MAX_DST];
    const int g = blockIdx.y * g_step + g_first;
    const BConvTable& tb = tables[g % n_tables];
    constexpr int ns = NS_MAX;
    const int nt = tb.nt;
    for (int i = threadIdx.x; i < nt * NS_MAX; i += 256) sf[i] = tb.f[i / NS_MAX][i % NS_MAX];
    for (int t = threadIdx.x; t < nt; t += 256) {
        const Modulus M = T.mod[tb.dst_mod[t]];
        sq[t] = M.qd; sqi[t] = M.qinv; sslot[t] = tb.dst_slot[t];
    }
    const int log_n = T.log_n;
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const u32 half = 1u << (log_n - 1);
    const u64* s = src + (size_t)(g / n_tables) * src_block_stride + idx;
    u64* d = dst + (size_t)g * dst_group_stride + idx;
    double y0[NS_MAX], y1[NS_MAX];
    u64 v0[NS_MAX], v1[NS_MAX];
#pragma unroll
    for (int k = 0; k < ns; ++k) {                                   // all loads in flight before the first use
        const u64* sp = s + ((size_t)tb.src_slot[k] << log_n);
        v0[k] = sp[0]; v1[k] = sp[half];
    }
#pragma unroll
    for (int k = 0; k < ns; ++k) {
        const u64 qk = T.mod[tb.src_mod[k]].q;
        const double qkd = u64_to_f(qk);
        y0[k] = v0[k] > (qk >> 1) ? d_add(u64_to_f(v0[k]), -qkd) : u64_to_f(v0[k]);
        y1[k] = v1[k] > (qk >> 1) ? d_add(u64_to_f(v1[k]), -qkd) : u64_to_f(v1[k]);
    }
    __syncthreads();
    // two targets per iteration: four independent accumulation chains per thread
    int t = 0;
    for (; t + 1 < nt; t += 2) {
        const double qa = sq[t], qb = sq[t + 1];
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int k = 0; k < ns; ++k) {
            const ConstF ca = sf[t * NS_MAX + k], cb = sf[(t + 1) * NS_MAX + k];
            a0 = d_add(a0, mulmod_const(y0[k], ca, qa));
            a1 = d_add(a1, mulmod_const(y1[k], ca, qa));
            b0 = d_add(b0, mulmod_const(y0[k], cb, qb));
            b1 = d_add(b1, mulmod_const(y1[k], cb, qb));
        }
        u64* da = d + ((size_t)sslot[t] << log_n);
        u64* db = d + ((size_t)sslot[t + 1] << log_n);
        da[0] = f_to_u64(reduce_canon(a0, qa, sqi[t]));
        da[half] = f_to_u64(reduce_canon(a1, qa, sqi[t]));
        db[0] = f_to_u64(reduce_canon(b0, qb, sqi[t + 1]));
        db[half] = f_to_u64(reduce_canon(b1, qb, sqi[t + 1]));
    }
    if (t < nt) {
        const double q = sq[t];
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < ns; ++k) {
            const ConstF c = sf[t * NS_MAX + k];
            a0 = d_add(a0, mulmod_const(y0[k], c, q));
            a1 = d_add(a1, mulmod_const(y1[k], c, q));
        }
        u64* dp = d + ((size_t)sslot[t] << log_n);
        dp[0] = f_to_u64(reduce_canon(a0, q, sqi[t]));
        dp[half] = f_to_u64(reduce_canon(a1, q, sqi[t]));
    }
}

// Base conversion with the exact three-FMA dot product of modarith.cuh (dot3): the centred digits are split at 2^23
// once per source, the table constants once per table, and a term costs three FMAs (A += xh ch, B += xs cs, C += xl cl)
// instead of the seven operations of mulmod_const + add; one reduction (dot3_finish, ~18 operations) per target.  A first
// form kept the split table in shared memory (k_bconv_dot3, round 2): three LDS per three FMAs made it load/store bound
// and slower than k_bconv (7 155 against 8 084 blocks/s on the AES-128 step); it was replaced by:
// the table as a KERNEL PARAMETER: a launch converts with exactly one table (the digits of a ModUp are launched one by
// one), so the split constants sit in the constant bank and, with the loops over targets and sources fully unrolled,
// every FMA takes its constant through a uniform register loaded from the constant bank -- no shared-memory load per
// term.  40 FP64 operations per output
// with seven sources (21 FMA, ~18 for the reduction, the splits of the sources amortised over the targets) against 54
// in k_bconv.  C coefficients per thread (idx, idx + N/2): two for digits of up to seven limbs, one above (3 NS C
// doubles of split sources in registers).  Same canonical residues as k_bconv.     grid: (N / (256 C), groups)
template <int NS>
struct BConvParam {
    int nt;
    int src_slot[NS];
    u64 src_q[NS];
    int dst_slot[FHE_MAX_DST];
    double q[FHE_MAX_DST], qi[FHE_MAX_DST];
    Split3 f[FHE_MAX_DST][NS];              // (Q/q_k) mod m_t split at 2^23
};
// targets t and t + 1 in one straight-line block (6 C independent FMA chains); the second is stored only if it exists
// (the table is zero beyond nt: its arithmetic is harmless and never leaves the registers)
template <int NS, int C>
FHE_D void bconv_param_pair(const BConvParam<NS>& tb, int t, bool second, const Split3 (&y)[C][NS], u64* d, u32 half, int log_n) {
    double A[2][C], B[2][C], Cc[2][C];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c) { A[u][c] = 0.0; B[u][c] = 0.0; Cc[u][c] = 0.0; }
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const Split3 f = tb.f[t + u][k];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                A[u][c] = d_fma(y[c][k].h, f.h, A[u][c]); B[u][c] = d_fma(y[c][k].s, f.s, B[u][c]);
                Cc[u][c] = d_fma(y[c][k].l, f.l, Cc[u][c]);
            }
        }
    double r[2][C];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c) r[u][c] = dot3_finish(A[u][c], B[u][c], Cc[u][c], tb.q[t + u], tb.qi[t + u]);
    u64* d0 = d + ((size_t)tb.dst_slot[t] << log_n);
#pragma unroll
    for (int c = 0; c < C; ++c) d0[c * half] = f_to_u64(r[0][c]);
    if (second) {
        u64* d1 = d + ((size_t)tb.dst_slot[t + 1] << log_n);
#pragma unroll
        for (int c = 0; c < C; ++c) d1[c * half] = f_to_u64(r[1][c]);
    }
}
template <int NS, int C>
__global__ void __launch_bounds__(256, (C == 2 || NS > 9 ? 2 : 3)) k_bconv_param(int log_n, const FHE_GRID_CONSTANT BConvParam<NS> tb, int n_tables,
                                                                       u64* dst, long long dst_group_stride, const u64* src,
                                                                       long long src_block_stride, int g_first, int g_step) {
    const int g = blockIdx.y * g_step + g_first;
    const u32 idx = blockIdx.x * 256 + threadIdx.x;
    const u32 half = 1u << (log_n - 1);
    const u64* s = src + (size_t)(g / n_tables) * src_block_stride + idx;
    u64* d = dst + (size_t)g * dst_group_stride + idx;
    u64 v[C][NS];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) v[c][k] = s[((size_t)tb.src_slot[k] << log_n) + c * half];      // all loads before the first use
    Split3 y[C][NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const u64 qk = tb.src_q[k];
        const double qkd = u64_to_f(qk);
#pragma unroll
        for (int c = 0; c < C; ++c)
            y[c][k] = split23(v[c][k] > (qk >> 1) ? d_add(u64_to_f(v[c][k]), -qkd) : u64_to_f(v[c][k]));
    }
    const int nt = tb.nt;
    // two targets per (uniform) test: the index into the table stays a compile-time constant
#pragma unroll
    for (int t = 0; t < FHE_MAX_DST; t += 2) {
        if (t >= nt) break;
        bconv_param_pair<NS, C>(tb, t, t + 1 < nt, y, d, half, log_n);
    }
}


// ---------------------------------------------------------------- key inner product
// acc[c][b][t] = sum_j e_bj[t] * ksk[j][c][id(t)],  e_bj[t] = d[b][t] if t in digit j else ext[b][j][t]
// grid: (N/256, nq + n_p).  One thread owns (t, idx) for EVERY ciphertext of the batch, so each
// key word is read from HBM once per batch.      ksk: [dnum][2][n_q_total + n_p][N]
// ext: [batch][beta][ne][N], d: [batch][nq][N], acc: [2][batch][ne][N]
#define FHE_MAX_BETA 8
struct U2 { u64 x, y; };
FHE_D U2 ld2(const u64* p) {
#ifndef FHE_EMU
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    U2 r; r.x = v.x; r.y = v.y; return r;
#else
    U2 r; r.x = p[0]; r.y = p[1]; return r;
#endif
}
FHE_D void st2(u64* p, u64 a, u64 b) {
#ifndef FHE_EMU
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
// grid: (N/256, nq + n_p).  A thread owns one coefficient of one limb for the whole batch: the key
// words are read once (2 * beta doubles in registers), then the batch is walked UNR ciphertexts at
// a time with all UNR * (beta + 2) loads issued before the first multiply -- the kernel is
// HBM-bound and what matters is the number of bytes in flight per SM (Little's law: ~150 KB).
// AB: fused multiply + relinearise + rescale.  `d` and `lift` then are the two operand ciphertexts a, b
// ([2][batch][nq][N]) and the tensor product is formed here: the digit's own limb is a1 b1, the lifted part
// P * (a0 b0, a0 b1 + a1 b0) -- the 3-polynomial product never goes through HBM.
// BETA (digits touched at this level) is compile-time and the presence of the lifted term is decided ONCE per CTA
// (one uniform branch into two specialised bodies), so the per-ciphertext arithmetic is straight-line code: with the
// runtime `j < beta` / `has_lift` tests of the first version every modular product was its own basic block.
// BETA == 0: no key-switch part at all (acc (+)= P * lift on the q-limbs, (+)= 0 on the special limbs).
// PTR (with AB): the operands of the fused multiply are gathered -- batch element b of the product is
// (a0[b], a1[b]) x (b0[b], b1[b]), four device pointers to polynomials of >= nq contiguous limbs each.  The table is a
// kernel parameter; the AES services multiply slices, permutations and concatenations of their state tensors through
// it without materialising them (fhe_mul_relin_rescale_ptrs).
#define FHE_MAX_MULB 128
struct MulPtrs {
    const u64* a0[FHE_MAX_MULB]; const u64* a1[FHE_MAX_MULB];
    const u64* b0[FHE_MAX_MULB]; const u64* b1[FHE_MAX_MULB];
};
template <int BETA, int UNR, bool AB, bool LIFT, bool PTR = false>
FHE_D void ks_inner_body(const DevTables& T, int nq, int alpha, int batch, u64* acc, const u64* ext, const u64* d,
                         const u64* ksk, const u64* lift, const ConstF* lift_c, int d_nq, int lift_nq, int lift_polys,
                         int accum, const MulPtrs* mp = nullptr) {
    constexpr int BE = BETA > 0 ? BETA : 1;
    const int t = blockIdx.y;
    const int ne = nq + T.n_p;
    const int id = t < nq ? t : T.n_q + (t - nq);
    const Modulus M = T.mod[id];
    const double q = M.qd, qi = M.qinv;
    const int log_n = T.log_n;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const size_t tot = (size_t)(T.n_q + T.n_p);
    double k0[BE], k1[BE];
#pragma unroll
    for (int j = 0; j < BETA; ++j) {
        const u64* kp = ksk + ((((size_t)j * 2) * tot + id) << log_n) + idx;
        k0[j] = u64_to_f(ld_u64(kp)); k1[j] = u64_to_f(ld_u64(kp + (tot << log_n)));
    }
    const int own = t < nq ? t / alpha : -1;
    const size_t ps = (size_t)batch * lift_nq << log_n;          // polynomial strides of lift / d
    const size_t pd = (size_t)batch * d_nq << log_n;
    ConstF pc; pc.w = 0; pc.wq = 0;
    if (LIFT) pc = lift_c[t];
    const u64* dp = d + ((size_t)t << log_n) + idx;                 // + b * nq * N
    const u64* ep = ext + ((size_t)t << log_n) + idx;               // + (b * beta + j) * ne * N
    const u64* lp = LIFT ? lift + ((size_t)t << log_n) + idx : nullptr;
    const bool two = AB || lift_polys == 2;
    // software pipeline: the loads of the next UNR ciphertexts are in flight while the current ones are
    // multiplied, so the memory system never waits for the FP64 work (and vice versa)
    // (AB: l0/l1 carry b0/b1 and m0/m1 carry a0/a1 of the q-limbs; the own-digit slot of e is not loaded)
    u64 e[UNR][BE], l0[UNR], l1[UNR], m0[UNR], m1[UNR];
    u64 en[UNR][BE], l0n[UNR], l1n[UNR], m0n[UNR], m1n[UNR];
    const u64* bp = AB && LIFT ? d + ((size_t)t << log_n) + idx : nullptr;      // operand a (named d), q-limbs only
    auto fetch = [&](int b0, u64 (&E)[UNR][BE], u64 (&L0)[UNR], u64 (&L1)[UNR], u64 (&M0)[UNR], u64 (&M1)[UNR]) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int b = min(b0 + u, batch - 1);                 // a clamped tail element is recomputed, not stored
#pragma unroll
            for (int j = 0; j < BETA; ++j) {
                if (AB) { E[u][j] = j != own ? ld_u64(ep + ((((size_t)b * BETA + j) * ne) << log_n)) : 0; }
                else E[u][j] = ld_u64(j == own ? dp + (((size_t)b * d_nq) << log_n)
                                               : ep + ((((size_t)b * BETA + j) * ne) << log_n));
            }
            if (LIFT) {
                if (PTR) {
                    const size_t off = ((size_t)t << log_n) + idx;
                    L0[u] = ld_u64(mp->b0[b] + off); L1[u] = ld_u64(mp->b1[b] + off);
                    M0[u] = ld_u64(mp->a0[b] + off); M1[u] = ld_u64(mp->a1[b] + off);
                } else {
                    L0[u] = ld_u64(lp + (((size_t)b * lift_nq) << log_n));
                    L1[u] = two ? ld_u64(lp + ps + (((size_t)b * lift_nq) << log_n)) : 0;
                    if (AB) {
                        M0[u] = ld_u64(bp + (((size_t)b * d_nq) << log_n));
                        M1[u] = ld_u64(bp + pd + (((size_t)b * d_nq) << log_n));
                    }
                }
            }
        }
    };
    fetch(0, e, l0, l1, m0, m1);
    for (int b0 = 0; b0 < batch; b0 += UNR) {
        if (b0 + UNR < batch) fetch(b0 + UNR, en, l0n, l1n, m0n, m1n);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int b = b0 + u;
            double a0 = 0.0, a1 = 0.0;
            double own_v = 0.0, t0 = 0.0, t1 = 0.0;
            if (AB && LIFT) {
                const double x0 = u64_to_f(m0[u]), x1 = u64_to_f(m1[u]), y0 = u64_to_f(l0[u]), y1 = u64_to_f(l1[u]);
                own_v = mulmod_var(x1, y1, q, qi);                                   // d2 on this limb
                t0 = mulmod_var(x0, y0, q, qi);                                      // d0
                // d1 = a0 b1 + a1 b0 = (a0 + a1)(b0 + b1) - d0 - d2: three products instead of four
                t1 = d_add(d_add(mulmod_var(d_add(x0, x1), d_add(y0, y1), q, qi), -t0), -own_v);   // |.| <= 1.6 q
            }
#pragma unroll
            for (int j = 0; j < BETA; ++j) {
                const double ev = AB && j == own ? own_v : u64_to_f(e[u][j]);
                a0 = d_add(a0, mulmod_var(ev, k0[j], q, qi));
                a1 = d_add(a1, mulmod_var(ev, k1[j], q, qi));
            }
            if (LIFT) {
                // fused relinearise + rescale: add P * (d0, d1) so the sum can be divided by P q_last at once
                a0 = d_add(a0, mulmod_const(AB ? t0 : u64_to_f(l0[u]), pc, q));
                a1 = d_add(a1, mulmod_const(AB ? t1 : u64_to_f(l1[u]), pc, q));
            }
            if (b < batch) {
                u64* o0 = acc + (((size_t)b * ne + t) << log_n) + idx;
                u64* o1 = acc + ((((size_t)batch + b) * ne + t) << log_n) + idx;
                if (accum) { a0 = d_add(a0, u64_to_f(*o0)); a1 = d_add(a1, u64_to_f(*o1)); }
                *o0 = f_to_u64(reduce_canon(a0, q, qi));
                *o1 = f_to_u64(reduce_canon(a1, q, qi));
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
            for (int j = 0; j < BETA; ++j) e[u][j] = en[u][j];
            l0[u] = l0n[u]; l1[u] = l1n[u];
            if (AB) { m0[u] = m0n[u]; m1[u] = m1n[u]; }
        }
    }
}
template <int BETA, int UNR, bool AB>
__global__ void __launch_bounds__(256, (BETA <= 4 ? 3 : 2)) k_ks_inner(DevTables T, int nq, int alpha, int batch,
                                                  u64* acc, const u64* ext, const u64* d, const u64* ksk,
                                                  const u64* lift, const ConstF* lift_c, int d_nq, int lift_nq,
                                                  int lift_polys, int accum) {
    // d_nq / lift_nq: limbs per batch element of `d` / `lift` (>= nq; only the AB operands may carry more)
    // lift_polys: 2, or 1 when only polynomial 0 is lifted (a rotation: (sigma c0, 0));  accum: acc += instead of acc =
    if (lift != nullptr && (int)blockIdx.y < nq)
        ks_inner_body<BETA, UNR, AB, true>(T, nq, alpha, batch, acc, ext, d, ksk, lift, lift_c, d_nq, lift_nq, lift_polys, accum);
    else
        ks_inner_body<BETA, UNR, AB, false>(T, nq, alpha, batch, acc, ext, d, ksk, lift, lift_c, d_nq, lift_nq, lift_polys, accum);
}
// fused multiply with gathered operands: acc = <ext, ksk> (d2 = a1 b1 on the digit's own limb) + P (a0 b0, a0 b1 + a1 b0)
template <int BETA, int UNR>
__global__ void __launch_bounds__(256, (BETA <= 4 ? 3 : 2)) k_ks_inner_ptr(DevTables T, int nq, int alpha, int batch, u64* acc,
                                                                          const u64* ext, const u64* ksk, const ConstF* lift_c,
                                                                          const FHE_GRID_CONSTANT MulPtrs mp) {
    if ((int)blockIdx.y < nq)
        ks_inner_body<BETA, UNR, true, true, true>(T, nq, alpha, batch, acc, ext, nullptr, ksk, ext, lift_c, nq, nq, 2, 0, &mp);
    else
        ks_inner_body<BETA, UNR, true, false, true>(T, nq, alpha, batch, acc, ext, nullptr, ksk, ext, lift_c, nq, nq, 2, 0, &mp);
}

// ---------------------------------------------------------------- fused NTT functors
// rescale: load the centred remainder of the dropped limb.  |r| <= q_last/2 < 2^44 is a valid
// lazy NTT input for any row modulus, so no reduction is needed here.
struct LoadCentered {
    const u64* last;            // [npoly][N] coefficient-domain rows of the dropped limb
    u64 q_last;
    FHE_D double operator()(const RowMap& map, RowRef row, int idx, int, const Modulus&) const {
        const u64 v = last[((size_t)row.blk << map.log_n) + idx];
        return v > (q_last >> 1) ? d_add(u64_to_f(v), -u64_to_f(q_last)) : u64_to_f(v);
    }
    FHE_D void prefetch(const RowMap& map, RowRef row, int idx, int count) const {
        const u64* p = last + ((size_t)row.blk << map.log_n) + idx;
        for (int k = 0; k < count; k += 16) prefetch_l2(p + k);
    }
};
// fused multiply + relinearise: the key switch transforms d2 = a1 * b1 without d2 ever being written
struct LoadMul {
    const u64* a; const u64* b;                 // polynomial 1 of each operand, rows [batch][>= nq]
    long long a_stride, b_stride;               // words per batch element (an operand may carry more limbs than nq)
    FHE_D double operator()(const RowMap& map, RowRef row, int idx, int, const Modulus& M) const {
        return canon(mulmod_var(u64_to_f(a[row_off(map, row, a_stride) + idx]), u64_to_f(b[row_off(map, row, b_stride) + idx]),
                                M.qd, M.qinv), M.qd);
    }
    FHE_D void prefetch(const RowMap& map, RowRef row, int idx, int count) const {
        const u64* pa = a + row_off(map, row, a_stride) + idx;
        const u64* pb = b + row_off(map, row, b_stride) + idx;
        for (int k = 0; k < count; k += 16) { prefetch_l2(pa + k); prefetch_l2(pb + k); }
    }
};
struct LoadMulPtr {         // the same product with gathered operands: polynomial 1 of batch element row.blk at a1[blk] / b1[blk]
    const u64* a1[FHE_MAX_MULB]; const u64* b1[FHE_MAX_MULB];
    FHE_D double operator()(const RowMap& map, RowRef row, int idx, int, const Modulus& M) const {
        const size_t off = ((size_t)row.j << map.log_n) + idx;
        return canon(mulmod_var(u64_to_f(a1[row.blk][off]), u64_to_f(b1[row.blk][off]), M.qd, M.qinv), M.qd);
    }
    FHE_D void prefetch(const RowMap& map, RowRef row, int idx, int count) const {
        const size_t off = ((size_t)row.j << map.log_n) + idx;
        for (int k = 0; k < count; k += 16) { prefetch_l2(a1[row.blk] + off + k); prefetch_l2(b1[row.blk] + off + k); }
    }
};
// rescale / ModDown epilogue: out = (in - ntt_value) * c[j]
struct StoreSubMul {
    u64* out; long long out_poly_stride;
    const u64* in; long long in_poly_stride;
    const ConstF* c;            // per limb slot j
    FHE_D void operator()(const RowMap& map, RowRef row, int idx, double v, int, const Modulus& M) const {
        const int poly = row.blk, j = row.j;
        const size_t lo = ((size_t)j << map.log_n) + idx;
        const double x = u64_to_f(in[(size_t)poly * in_poly_stride + lo]);
        out[(size_t)poly * out_poly_stride + lo] = f_to_u64(canon(mulmod_const(d_add(x, -v), c[j], M.qd), M.qd));
    }
    // the epilogue reads `in`: bring the 128-byte line the 16 lanes of a row cover into L2 early
    FHE_D void warm(const RowMap& map, RowRef row, int idx) const {
        prefetch_l2(in + (size_t)row.blk * in_poly_stride + ((size_t)row.j << map.log_n) + idx);
    }
    FHE_D double fetch(const RowMap& map, RowRef row, int idx) const {
        return u64_to_f(ld_u64(in + (size_t)row.blk * in_poly_stride + ((size_t)row.j << map.log_n) + idx));
    }
    FHE_D void put(const RowMap& map, RowRef row, int idx, double v, int, const Modulus& M, double x) const {
        out[(size_t)row.blk * out_poly_stride + ((size_t)row.j << map.log_n) + idx] =
            f_to_u64(canon(mulmod_const(d_add(x, -v), c[row.j], M.qd), M.qd));
    }
};
