// modarith.cuh -- 64-bit modular arithmetic for moduli q < 2^62.
//
// * Shoup multiplication by a constant w with companion w' = floor(w 2^64 / q): used for
//   every NTT twiddle and every per-limb constant (1 mul.hi + 2 mul.lo).
// * Harvey lazy butterflies keep values in [0,4q) (forward) / [0,2q) (inverse).
// * Barrett reduction of a 128-bit value with mu = floor(2^128/q) for variable x variable
//   products and for the 128-bit accumulators of the base-conversion / key inner products.
#pragma once
#include "compat.h"

struct Modulus {
    u64 q;
    u64 mu_hi, mu_lo;     // floor(2^128 / q)
};

struct ShoupConst {
    u64 w, ws;            // w in [0,q),  ws = floor(w * 2^64 / q)
};

FHE_HD u64 add_mod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
FHE_HD u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
FHE_HD u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// x*w mod q in [0,2q) for ANY 64-bit x.
FHE_D u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
    u64 hi = umulhi64(x, ws);
    return x * w - hi * q;
}
// canonical result
FHE_D u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
    u64 r = mul_shoup_lazy(x, w, ws, q);
    return r >= q ? r - q : r;
}

// 128-bit helpers -------------------------------------------------------------------
struct u128t { u64 lo, hi; };

FHE_D u128t mul_wide(u64 a, u64 b) {
    u128t r; r.lo = a * b; r.hi = umulhi64(a, b); return r;
}
FHE_D void acc_wide(u128t& acc, u64 a, u64 b) {
    u64 lo = a * b, hi = umulhi64(a, b);
    acc.lo += lo;
    acc.hi += hi + (acc.lo < lo ? 1ull : 0ull);
}
// z mod q for any 128-bit z with z < q * 2^64  (always true for our accumulators: <= 16
// products of a 62-bit by a 62-bit value is < 2^128 but we additionally keep z.hi < q by
// construction: see callers).  Result canonical.
FHE_D u64 barrett_reduce(u128t z, const Modulus& m) {
    // t ~= floor(z * mu / 2^128), under-estimated by at most 4
    u64 t = umulhi64(z.hi, m.mu_lo) + umulhi64(z.lo, m.mu_hi) + z.hi * m.mu_hi;
    u64 r = z.lo - t * m.q;
    // r < 5q < 2^64
    if (r >= 2 * m.q) r -= 2 * m.q;
    if (r >= 2 * m.q) r -= 2 * m.q;
    if (r >= m.q) r -= m.q;
    return r;
}
FHE_D u64 mul_mod(u64 a, u64 b, const Modulus& m) { return barrett_reduce(mul_wide(a, b), m); }

// value v < 2^64 reduced mod q (q may be much smaller than v)
FHE_D u64 reduce_u64(u64 v, const Modulus& m) {
    u128t z; z.lo = v; z.hi = 0;
    return barrett_reduce(z, m);
}
