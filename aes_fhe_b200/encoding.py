"""CKKS canonical-embedding encode / decode (host side, complex128 NumPy).

Slot j is the evaluation of the plaintext polynomial at xi^(5^j), xi = exp(i*pi/N);
the conjugate half of the spectrum is filled so the coefficients are real.  This is
the convention that makes ``rotate(ct, key, d)`` equal ``np.roll(slots, d)`` as pinned
by /root/reference/test/test_engine_rot.py:32-40, and under which the reference's
"constant plaintexts" ``np.full(slot_count, c)`` (/root/reference/xor_service.py:192,
sbox/sbox_service.py:85-88) encode to the two-term polynomial Re(c) + Im(c) X^(N/2).

Only the float <-> integer-coefficient step lives here; residues are produced by the
backend (``from_i64``) so the CUDA path and the CPU oracle see identical integers.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np


@lru_cache(maxsize=8)
def _tables(log_n: int):
    n = 1 << log_n
    half = n >> 1
    two_n = 2 * n
    # k_j with 2*k_j + 1 = 5^j mod 2N ; conjugate slot at 2N - 5^j
    pw = np.empty(half, dtype=np.int64)
    g = 1
    for j in range(half):
        pw[j] = g
        g = (g * 5) % two_n
    k_pos = (pw - 1) // 2
    k_neg = (two_n - pw - 1) // 2
    twist = np.exp(1j * np.pi * np.arange(n) / n)          # xi^n
    return k_pos, k_neg, twist


def slots_to_coeffs(values: np.ndarray, log_n: int) -> np.ndarray:
    """Real coefficient vector m (float64, length N) with m(xi^(5^j)) = values[j]."""
    n = 1 << log_n
    k_pos, k_neg, twist = _tables(log_n)
    z = np.zeros(n >> 1, dtype=np.complex128)
    v = np.asarray(values, dtype=np.complex128).ravel()
    if v.size > z.size:
        raise ValueError(f"too many slots: {v.size} > {z.size}")
    z[: v.size] = v
    spec = np.empty(n, dtype=np.complex128)
    spec[k_pos] = z
    spec[k_neg] = np.conj(z)
    # m_n = xi^-n * (1/N) sum_k spec_k e^{-2 pi i k n / N}
    m = np.fft.fft(spec) * np.conj(twist) / n
    return m.real


def coeffs_to_slots(m: np.ndarray, log_n: int) -> np.ndarray:
    n = 1 << log_n
    k_pos, _, twist = _tables(log_n)
    spec = np.fft.ifft(np.asarray(m, dtype=np.float64) * twist) * n
    return spec[k_pos]


def encode_i64(values: np.ndarray, scale: float, log_n: int) -> np.ndarray:
    """Scaled, rounded integer coefficients (int64)."""
    m = slots_to_coeffs(values, log_n) * scale
    if np.max(np.abs(m)) >= 2.0 ** 62:
        raise OverflowError("plaintext magnitude too large for the scale")
    return np.rint(m).astype(np.int64)


def const_i64(value: complex, scale) -> tuple[int, int]:
    """A constant slot vector is  round(s*Re) + round(s*Im) X^(N/2): return both integers."""
    from fractions import Fraction
    s = Fraction(scale)
    v = complex(value)
    re = int(round(Fraction(v.real) * s))
    im = int(round(Fraction(v.imag) * s))
    return re, im


def as_constant(values: np.ndarray):
    """Return the scalar if every slot holds the same value, else None."""
    v = np.asarray(values)
    if v.ndim == 0:
        return complex(v)
    if v.size == 0:
        return None
    first = v.flat[0]
    if np.all(v == first):
        return complex(first)
    return None
