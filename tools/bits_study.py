#!/usr/bin/env python
"""Design study for the bit-sliced (+-1) AES-128 schedule (aes_fhe_b200/services/aes_bits.py): plain
NumPy, no encryption.  It checks the three facts the schedule rests on and prints the numbers DESIGN.md
quotes:

1. every S-box output bit is a multilinear polynomial sum_{A,B} w[A,B] m_A(hi) m_B(lo) of the 16 x 16
   monomials of the high / low four input bits (Walsh-Hadamard spectrum): exactness on all 256 bytes,
   sum |w|, and the first-order error gain at Boolean points (<= 8: every partial derivative is in
   {-1, 0, 1});
2. MixColumns + AddRoundKey as products of +-1 values (depth 3, 140 products per 32 state bits);
3. the error recursion of ten rounds when every round ends in the bit bootstrap
   s -> sin(pi/2 (s + e)) + d  (the refresh squares the incoming error; only the EvalMod error d stays).

    python tools/bits_study.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aes_fhe_b200.services import aes_bits as AB          # noqa: E402
from oracle import aes_plain as A                          # noqa: E402


def main():
    W = AB.sbox_walsh()                                   # [8 outputs, 16 hi-monomials, 16 lo-monomials]
    x = np.arange(256)
    s = AB.bits_pm(x)                                     # [8, 256]  +-1
    mh, ml = AB.monomials(s[4:]), AB.monomials(s[:4])     # [16, 256]
    out = np.einsum("kab,ax,bx->kx", W, mh, ml)
    want = AB.bits_pm(A.SBOX[x])
    print("S-box Walsh form: max |err| over 256 bytes =", np.abs(out - want).max())
    print("  sum |w| per output bit:", np.round(np.abs(W).sum(axis=(1, 2)), 2))
    rng = np.random.default_rng(0)
    n = 1 << 17
    xb = rng.integers(0, 256, n)
    for e_in in (1e-4, 1e-3, 1e-2, 5e-2):
        sn = AB.bits_pm(xb) + rng.normal(0, e_in, (8, n))
        o = np.einsum("kab,ax,bx->kx", W, AB.monomials(sn[4:]), AB.monomials(sn[:4]))
        err = np.abs(o - AB.bits_pm(A.SBOX[xb]))
        print(f"  input rms {e_in:.0e}: output rms {err.std():.2e}  max {err.max():.2e}")

    # ten rounds with the refresh model
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    rks = A.key_schedule(key)
    blocks = rng.integers(0, 256, (8192, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(bytes.fromhex("3243f6a8885a308d313198a2e0370734"), np.uint8)
    for d in (1e-4, 1e-3, 3e-3, 1e-2):
        st = AB.PlainBits.from_blocks(blocks)
        st = AB.PlainBits.xor(st, AB.PlainBits.from_key(rks[0], blocks.shape[0]))
        worst_in = worst_out = 0.0
        for r in range(1, 11):
            st = AB.PlainBits.shift_rows(st)
            ideal = np.sign(st)
            worst_in = max(worst_in, np.abs(st - ideal).max())
            st = np.sin(np.pi / 2 * st) + rng.normal(0, d, st.shape)            # bit bootstrap
            worst_out = max(worst_out, np.abs(st - ideal).max())
            st = AB.PlainBits.sub_bytes(st, W)
            st = AB.PlainBits.mix_ark(st, AB.PlainBits.from_key(rks[r], blocks.shape[0]), last=(r == 10))
        got = AB.PlainBits.to_blocks(st)
        ok = np.array_equal(got, A.encrypt_blocks(blocks, key))
        print(f"refresh noise rms {d:.0e}: max |err| before a refresh {worst_in:.2e}, after {worst_out:.2e}, "
              f"AES-128 bytes exact: {ok}, FIPS-197 App. B: {got[0].tobytes().hex() == '3925841d02dc09fbdc118597196a0b32'}")


if __name__ == "__main__":
    main()
