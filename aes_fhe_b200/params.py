"""RNS-CKKS parameter sets for the B200 engine (pure Python, deterministic).

The reference never sees the parameter set: it only forwards ``max_level`` /
``log_coeff_count`` / ``special_prime_count`` to ``desilofhe.Engine``
(/root/reference/engine_context.py:32-56).  Everything below is therefore this
repo's own specification; it is *data* that both the CUDA backend and the CPU
oracle (oracle/refmod.cpp) consume, so that their residues can be compared
bit-exactly.

Chain layout (limb index -> modulus):
    0            q_0   "base" prime  (BASE_BITS bits)
    1 .. L       q_l   scale primes  (~2^SCALE_BITS), one dropped per rescale
    L+1 .. L+K   p_k   special primes (SPECIAL_BITS bits) for hybrid key-switching

Scale discipline: every ciphertext at level l carries exactly the scale
``delta[l]`` with ``delta[L] = 2^SCALE_BITS`` and ``delta[l-1] = delta[l]^2 / q_l``
(so ct x ct and ct x pt followed by one rescale land on the table again).  The
scale primes are picked greedily, top level first, as the unused NTT prime that
is closest to ``delta[l]^2 / 2^SCALE_BITS`` so the drift never compounds.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
from functools import lru_cache
from typing import List, Tuple

import numpy as np

# every modulus is < 2^45: the kernels do their modular arithmetic exactly on the FP64 pipe
# (csrc/modarith.cuh), which needs all lazy intermediates below 2^50
BASE_BITS = 45
SPECIAL_BITS = 45
DEFAULT_SCALE_BITS = 40
DEFAULT_DNUM = 4
# log2(P Q) allowed at N = 2^16 for 128-bit classical security, by secret distribution:
#  * uniform ternary secret: 1770 -- the HomomorphicEncryption.org standard table (881 bits at N = 2^15) doubled, the
#    value libraries quote for N = 2^16;
#  * sparse ternary secret of Hamming weight 192 (what bootstrapping's ModRaise needs): 1553 -- the largest of the
#    published bootstrappable N = 2^16, h = 192 sets (Bossuat, Mouchet, Troncoso-Pastoriza, Hubaux, EUROCRYPT 2021,
#    sets I-IV: log QP 1546..1553, estimated at 128 bits with the hybrid dual attack included).
# No lattice estimator is installed in this image; these are cited bounds, not a fresh estimate.
LOG_PQ_BUDGET_DENSE = 1770
LOG_PQ_BUDGET_SPARSE = 1553
LOG_PQ_BUDGET_2_16 = LOG_PQ_BUDGET_DENSE


# --------------------------------------------------------------------------- #
# number theory helpers
# --------------------------------------------------------------------------- #
_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n: int) -> bool:
    """Deterministic Miller-Rabin for n < 3.3e24 (covers every 64-bit modulus)."""
    if n < 2:
        return False
    for p in _MR_BASES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def primitive_2n_root(q: int, two_n: int) -> int:
    """Smallest-generator primitive ``two_n``-th root of unity mod q (psi^N = -1)."""
    assert (q - 1) % two_n == 0
    e = (q - 1) // two_n
    g = 2
    while True:
        psi = pow(g, e, q)
        if pow(psi, two_n // 2, q) == q - 1:
            return psi
        g += 1


def _ntt_primes_near(center: int, two_n: int):
    """Yield NTT-friendly primes (= 1 mod two_n) in order of distance from center."""
    base = center - (center % two_n) + 1
    lo, hi = base, base + two_n
    while True:
        # emit whichever of the two frontier candidates is closer to the centre
        if center - lo <= hi - center:
            if lo > two_n and is_prime(lo):
                yield lo
            lo -= two_n
        else:
            if is_prime(hi):
                yield hi
            hi += two_n


def _primes_below(bound: int, two_n: int, count: int, exclude=()) -> List[int]:
    out = []
    c = bound - (bound % two_n) + 1
    if c >= bound:
        c -= two_n
    while len(out) < count:
        if c not in exclude and is_prime(c):
            out.append(c)
        c -= two_n
    return out


# --------------------------------------------------------------------------- #
# parameter set
# --------------------------------------------------------------------------- #
@dataclass(frozen=True)
class CKKSParams:
    log_n: int
    max_level: int                    # L : a fresh ciphertext has limbs 0..L
    special_count: int                # K
    alpha: int                        # q-limbs per key-switch digit
    scale_bits: int
    moduli: Tuple[int, ...]           # q_0..q_L, p_0..p_{K-1}
    psi: Tuple[int, ...]              # primitive 2N-th root per modulus
    delta: Tuple[Fraction, ...] = field(repr=False, default=())   # exact scale per level

    @property
    def n(self) -> int:
        return 1 << self.log_n

    @property
    def slot_count(self) -> int:
        return 1 << (self.log_n - 1)

    @property
    def n_q(self) -> int:
        return self.max_level + 1

    @property
    def n_p(self) -> int:
        return self.special_count

    @property
    def q(self) -> Tuple[int, ...]:
        return self.moduli[: self.n_q]

    @property
    def p(self) -> Tuple[int, ...]:
        return self.moduli[self.n_q:]

    @property
    def dnum(self) -> int:
        return -(-self.n_q // self.alpha)

    def digits_at(self, n_active: int) -> int:
        """beta: number of key-switch digits touching limbs 0..n_active-1."""
        return -(-n_active // self.alpha)

    def scale(self, level: int) -> float:
        return float(self.delta[level])

    def galois_for_rotation(self, delta: int) -> int:
        """Galois element whose automorphism realises ``np.roll(slots, delta)``.

        Pinned by /root/reference/test/test_engine_rot.py:32-40 (rotate(ct,+5)
        == np.roll(base,+5)).  sigma_{5^r} moves slot j+r into slot j, i.e. a
        left rotation by r, so roll(+delta) needs r = -delta.
        """
        r = (-delta) % self.slot_count
        return pow(5, r, 2 * self.n)

    @property
    def galois_conj(self) -> int:
        return 2 * self.n - 1

    @property
    def log_pq(self) -> float:
        """log2 of the full modulus P Q (what the security bound is stated on)"""
        import math
        return sum(math.log2(m) for m in self.moduli)


@lru_cache(maxsize=None)
def auto_dnum(log_n: int, max_level: int, scale_bits: int, budget: int = LOG_PQ_BUDGET_DENSE) -> int:
    """Key-switch digit count when the caller does not give one.

    Fewer, wider digits mean fewer ModUp rows to transform (beta (n + K) - n per key switch), a smaller
    key stream (2 beta (n + K) limbs) and fewer inner-product terms, at the price of more special primes
    (P must exceed the widest digit).  On the B200 the trade is a clear win as long as P Q stays inside
    the security budget: SubBytes at max_level = 22 takes 55.1 ms per 16 ciphertexts with 4 digits,
    54.0 with 3 and 53.7 with 2 (profiles/r01_dnum_sweep.md).  So at N = 2^16: the smallest digit count
    in {2, 3, 4} whose log2(P Q) fits the budget of the secret distribution in use.  Smaller rings (tests) keep
    four digits."""
    if log_n != 16:
        return DEFAULT_DNUM
    n_q = max_level + 1
    for d in (2, 3):
        alpha = -(-n_q // d)
        digit_bits = BASE_BITS + (alpha - 1) * scale_bits
        k = -(-(digit_bits + 1) // (SPECIAL_BITS - 1))          # special primes are just below 2^45
        if BASE_BITS + max_level * scale_bits + k * SPECIAL_BITS <= budget:
            return d
    return DEFAULT_DNUM


@lru_cache(maxsize=None)
def make_params(log_n: int = 16, max_level: int = 30, special_count: int = 0,
                dnum: int = 0, scale_bits: int = DEFAULT_SCALE_BITS, log_pq_budget: int = 0) -> CKKSParams:
    """Build the deterministic parameter set.  ``dnum == 0`` means "choose" (auto_dnum) within
    ``log_pq_budget`` (0 = the bound for a uniform ternary secret; engines with a sparse secret pass
    LOG_PQ_BUDGET_SPARSE).

    ``special_count == 0`` means "derive K from the digit size" (enough special
    primes that P exceeds the largest digit product, the hybrid key-switching
    noise condition).
    """
    if not (10 <= log_n <= 16):
        raise ValueError("log_coeff_count must be in [10, 16]")
    if not (30 <= scale_bits <= 44):
        raise ValueError("scale_bits must be in [30, 44] (every modulus has to stay below 2^45)")
    if max_level < 1:
        raise ValueError("max_level must be >= 1")
    n = 1 << log_n
    two_n = 2 * n
    n_q = max_level + 1
    if dnum <= 0:
        dnum = auto_dnum(log_n, max_level, scale_bits, log_pq_budget or LOG_PQ_BUDGET_DENSE)

    q0 = _primes_below(1 << BASE_BITS, two_n, 1)[0]

    # scale primes with drift cancellation, exact rational bookkeeping
    target = 1 << scale_bits
    delta = [None] * n_q
    delta[max_level] = Fraction(target)
    q = [0] * n_q
    q[0] = q0
    used = set()
    for l in range(max_level, 0, -1):
        want = delta[l] * delta[l] / target          # q_l that would restore 2^scale_bits
        for cand in _ntt_primes_near(int(want), two_n):
            if cand not in used:
                q[l] = cand
                used.add(cand)
                break
        # keep ~600 bits of the exact value (squaring would double the size every level)
        delta[l - 1] = (delta[l] * delta[l] / q[l]).limit_denominator(1 << 600)

    if special_count <= 0:
        alpha = -(-n_q // max(1, dnum))
        # smallest K whose special modulus exceeds the largest digit (the one holding q_0)
        digit = 1
        for m in q[:alpha]:
            digit *= m
        cand = _primes_below(1 << SPECIAL_BITS, two_n, 48, exclude={q0} | used)
        special_count, prod = 0, 1
        while prod <= digit:
            prod *= cand[special_count]
            special_count += 1
    else:
        # K given (signature 3): largest alpha whose digit still fits under P
        p_bits = special_count * (SPECIAL_BITS - 1)
        alpha = max(1, min(n_q, 1 + (p_bits - BASE_BITS - 1) // scale_bits))
    p = _primes_below(1 << SPECIAL_BITS, two_n, special_count, exclude={q0} | used)
    moduli = tuple(q) + tuple(p)
    psi = tuple(primitive_2n_root(m, two_n) for m in moduli)
    # limit denominators: the exact value is kept, floats are derived on demand
    return CKKSParams(log_n=log_n, max_level=max_level, special_count=special_count,
                      alpha=alpha, scale_bits=scale_bits, moduli=moduli, psi=psi,
                      delta=tuple(delta))


# --------------------------------------------------------------------------- #
# host-side tables shared by keygen / encode (small, exact integer maths)
# --------------------------------------------------------------------------- #
def bit_reverse(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def sqrt_minus_one(params: CKKSParams, limb: int) -> int:
    """psi^(N/2): the NTT image of X^(N/2) is +I on the first half of the
    bit-reversed spectrum and -I on the second half."""
    return pow(params.psi[limb], params.n // 2, params.moduli[limb])


def limb_ids(params: CKKSParams, n_q_active: int, with_special: bool) -> List[int]:
    ids = list(range(n_q_active))
    if with_special:
        ids += list(range(params.n_q, params.n_q + params.n_p))
    return ids
