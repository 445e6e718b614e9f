// modarith.cuh -- modular arithmetic for moduli q < 2^45, carried out on the FP64 pipe.
//
// Why FP64: measured on B200 (tools/ubench/pipes.cu) the SM retires 63 DFMA/DMUL/DADD per
// clock but only 21 IMAD.WIDE (32x32->64) and 6.8 mul.hi.u64, so a 64-bit Shoup/Barrett
// product costs about three times as many issue cycles as the same product done with
// error-free double arithmetic.  All residues are integers below 2^45, every intermediate
// value stays an exactly representable integer (|x| < 2^50), and the results are the same
// canonical residues an integer implementation produces -- bit-exact with oracle/refmod.cpp.
//
//   mulmod by a constant w (Shoup style, wq = RN(w/q) precomputed):
//       qe = rint(x * wq)                  (one FMA against the 1.5*2^52 magic constant)
//       p  = RN(x * w),  plo = fma(x, w, -p)           (exact product as p + plo)
//       r  = fma(-qe, q, p) + plo          (both steps exact)      |r| <= q (1/2 + |x| 2^-53)
//   six FP64 operations; a Cooley-Tukey butterfly is eight.  Values are kept signed and lazy:
//   a forward transform never needs a correction (|x| grows by < 0.63 q per stage).
//
// A few cold paths (reducing 63-bit encoder coefficients, CRT decode) still use integer Barrett.
#pragma once
#include "compat.h"

struct Modulus {
    u64 q;
    u64 mu_hi, mu_lo;     // floor(2^128 / q)      (integer Barrett, cold paths)
    double qd, qinv;      // (double) q, RN(1/q)
};

struct ConstF {           // a constant c in [0,q) prepared for mulmod_const
    double w, wq;         // (double) c, RN(c / q)
};

#define FHE_MAGIC 6755399441055744.0      /* 1.5 * 2^52 */
#define FHE_TWO52 4503599627370496.0      /* 2^52 */

#ifndef FHE_EMU
FHE_D double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
FHE_D double d_mul(double a, double b) { return __dmul_rn(a, b); }
FHE_D double d_add(double a, double b) { return __dadd_rn(a, b); }
FHE_D double bits_to_f(u64 b) { return __longlong_as_double((long long)b); }
FHE_D u64 f_to_bits(double d) { return (u64)__double_as_longlong(d); }
#else
#include <cmath>
#include <cstring>
inline double d_fma(double a, double b, double c) { return std::fma(a, b, c); }
inline double d_mul(double a, double b) { volatile double r = a * b; return r; }
inline double d_add(double a, double b) { volatile double r = a + b; return r; }
inline double bits_to_f(u64 b) { double d; std::memcpy(&d, &b, 8); return d; }
inline u64 f_to_bits(double d) { u64 b; std::memcpy(&b, &d, 8); return b; }
#endif

// exact u64 (< 2^52) <-> double without conversion instructions
FHE_D double u64_to_f(u64 x) { return d_add(bits_to_f(0x4330000000000000ull | x), -FHE_TWO52); }
FHE_D u64 f_to_u64(double r) { return f_to_bits(d_add(r, FHE_TWO52)) & 0x000FFFFFFFFFFFFFull; }

// rint(x * c) for |x * c| < 2^51
#ifdef FHE_RINT_QUOT
// variant: DMUL + FRND.F64 -- the round-to-integer issues on the conversion pipe (16 lanes/clk/SM,
// tools/ubench/frnd.cu) instead of a second FP64-pipe operation.  The quotient may differ by one from
// the single-rounding form in rare near-tie cases, which only changes the lazy representative.
#ifndef FHE_EMU
FHE_D double round_quot(double x, double c) { return rint(d_mul(x, c)); }
#else
inline double round_quot(double x, double c) { return std::rint(d_mul(x, c)); }
#endif
#else
FHE_D double round_quot(double x, double c) { return d_add(d_fma(x, c, FHE_MAGIC), -FHE_MAGIC); }
#endif

// x * w mod q, signed result with |r| <= q (1/2 + |x| 2^-53); needs |x| < 2^50
FHE_D double mulmod_const(double x, const ConstF w, double q) {
    const double qe = round_quot(x, w.wq);
    const double p = d_mul(x, w.w);
    const double plo = d_fma(x, w.w, -p);
    return d_add(d_fma(-qe, q, p), plo);
}
// a * b mod q for |a|, |b| < 2^47 (so that |a b| / q < 2^51); |r| <= 0.51 q
FHE_D double mulmod_var(double a, double b, double q, double qinv) {
    const double p = d_mul(a, b);
    const double plo = d_fma(a, b, -p);
    const double qe = round_quot(p, qinv);
    return d_add(d_fma(-qe, q, p), plo);
}
// symmetric reduction of a lazy value, |x| < 2^51  ->  |r| <= q/2 (+1)
FHE_D double reduce_sym(double x, double q, double qinv) { return d_fma(-round_quot(x, qinv), q, x); }
// (-q, q) -> [0, q)
FHE_D double canon(double r, double q) { return r < 0.0 ? d_add(r, q) : r; }
FHE_D double reduce_canon(double x, double q, double qinv) { return canon(reduce_sym(x, q, qinv), q); }

// ---- exact dot products with constants:  sum_t x_t c_t mod q  in three FP64 operations per term
// Both factors are split at 2^23 (x = xh 2^23 + xl, c = ch 2^23 + cl, |xl|, |cl| <= 2^22, xh, ch <= 2^22), and
// the three Karatsuba sums  A = sum xh ch,  B = sum (xh+xl)(ch+cl),  C = sum xl cl  are plain FMA chains whose
// partial sums stay below 2^51 for up to 16 terms -- every FMA is exact.  The split of x is done once per input
// and the split of c once per table entry, so a term costs 3 FP64 operations instead of the 7 of
// mulmod_const + add; the modular reduction happens once per output:
//     sum = (A 2^23 + (B - A - C)) 2^23 + C       (Horner; a product by 2^23 is exact, so each step is a
//                                                  quotient estimate and one FMA)
// Needs |x_t| < 2^45, 0 <= c_t < 2^45, at most 16 terms, q > 2^24.
#define FHE_TWO23 8388608.0
#define FHE_INV_TWO23 1.1920928955078125e-07
struct Split3 { double h, l, s; };     // v = h 2^23 + l,  s = h + l
FHE_D Split3 split23(double v) {
    Split3 r;
    r.h = d_add(d_fma(v, FHE_INV_TWO23, FHE_MAGIC), -FHE_MAGIC);
    r.l = d_fma(r.h, -FHE_TWO23, v);
    r.s = d_add(r.h, r.l);
    return r;
}
// x 2^23 mod q, signed, |r| <= q (1/2 + |x| 2^-52): the product by a power of two is exact, so no error term
FHE_D double mul23_mod(double x, double q, double wq23) {
    const double qe = round_quot(x, wq23);
    return d_fma(-qe, q, d_mul(x, FHE_TWO23));
}
// canonical residue of A 2^46 + (B - A - C) 2^23 + C
FHE_D double dot3_finish(double A, double B, double C, double q, double qinv) {
    const double wq23 = d_mul(FHE_TWO23, qinv);
    const double mid = d_add(d_add(B, -A), -C);
    const double t1 = mul23_mod(A, q, wq23);
    const double t2 = mul23_mod(d_add(t1, mid), q, wq23);
    return reduce_canon(d_add(t2, C), q, qinv);
}

// ---- integer helpers (canonical add/sub are pure ALU work; Barrett only on cold paths)
FHE_HD u64 add_mod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
FHE_HD u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
FHE_HD u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

struct u128t { u64 lo, hi; };
FHE_D u128t mul_wide(u64 a, u64 b) { u128t r; r.lo = a * b; r.hi = umulhi64(a, b); return r; }
FHE_D u64 barrett_reduce(u128t z, const Modulus& m) {
    u64 t = umulhi64(z.hi, m.mu_lo) + umulhi64(z.lo, m.mu_hi) + z.hi * m.mu_hi;
    u64 r = z.lo - t * m.q;
    if (r >= 2 * m.q) r -= 2 * m.q;
    if (r >= 2 * m.q) r -= 2 * m.q;
    if (r >= m.q) r -= m.q;
    return r;
}
FHE_D u64 mul_mod(u64 a, u64 b, const Modulus& m) { return barrett_reduce(mul_wide(a, b), m); }
FHE_D u64 reduce_u64(u64 v, const Modulus& m) { u128t z; z.lo = v; z.hi = 0; return barrett_reduce(z, m); }
