"""AESFHETransformer: the reference's zeta_256-domain "merged ShiftRows+MixColumns" and its inverse.  Mirror of
/root/reference/shiftrow_mixcolumns.py:8-131 (row a12) -- same constructor, same two methods, same order of engine
calls:

  merged_shift_mix   (:16-80)   4 x [state x row product (+ relinearize), rotate -1/-2/-3, three ct x ct products,
                                mask product], 4 x collapse [rotate -2, XOR LUT, rotate -1, XOR LUT, mask product
                                (+ relinearize)], rotate 0/-5/-10/-15 and three XOR LUTs;
  merged_inv_mixshift (:82-131) 4 x [product, rotate -1/-2/-3, three XOR LUTs], 4 x collapse [rotate -2, XOR LUT,
                                rotate -1, XOR LUT (+ relinearize)], rotate 0/-5/-10/-15 and three XOR LUTs.

SURVEY defects D4 / D8 apply: the sequence is not AES (a product of zeta values adds exponents, it does not multiply in
GF(2^8)), and the reference's ``ZetaEncoder.to_zeta(uint8, modulus=256)`` overflows on NumPy 2 -- this package's
ZetaEncoder casts first.  It is kept for operation-sequence parity against a plain-complex evaluation of the same
calls (tests/test_services_plain_and_oracle.py); it needs bootstrapping (xor_cipher refreshes operands below level 8).
The FIPS-correct ShiftRows / MixColumns of the throughput path are in services/aes_bits.py.
"""
from __future__ import annotations

import numpy as np

from .xor_service import EngineWrapper, XORService, ZetaEncoder

_X = ([2, 3, 1, 1], [1, 1, 2, 3], [1, 1, 2, 3], [3, 1, 1, 2])                       # as written at :28-33
_X_INV = ([14, 11, 13, 9], [9, 14, 11, 13], [13, 9, 14, 11], [11, 13, 9, 14])
_SHIFTS = (0, 5, 10, 15)


class AESFHETransformer:
    def __init__(self, xor_service: XORService, engine_wrapper: EngineWrapper):
        self.xor_svc = xor_service
        self.eng = engine_wrapper

    def _padded_ct(self, values):
        """zeta_256 encoding, padded with ones to the slot count (:38-41)"""
        z = ZetaEncoder.to_zeta(np.asarray(values), modulus=256)
        sc = self.eng.engine.slot_count
        if z.size < sc:
            z = np.pad(z, (0, sc - z.size), constant_values=1.0)
        return self.eng.encrypt(z)

    def _row_cts(self, table):
        return [self._padded_ct(np.repeat(np.array(row, dtype=np.uint8), 4)) for row in table]

    def _mask_ct(self):
        return self._padded_ct(np.array([1 if i % 4 == 0 else 0 for i in range(16)], dtype=np.uint8))

    def _combine(self, cts):
        out = None
        for shift, ct in zip(_SHIFTS, cts):
            p = self.eng.rotate(ct, -shift)
            out = p if out is None else self.xor_svc.xor_cipher(out, p)
        return out

    def merged_shift_mix(self, state_bytes: np.ndarray):
        return self._combine(self.collapsed_columns(state_bytes))

    def collapsed_columns(self, state_bytes: np.ndarray):
        """steps 1-4 of merged_shift_mix (:22-72): the four collapsed column ciphertexts, before the final
        rotate-and-XOR combination.  Their slots stay below 7 in magnitude, so they can be compared with a
        plain-complex evaluation on a real engine; the combination (:74-80) feeds values of magnitude 4..6 into the
        degree-15 x degree-15 XOR polynomial and leaves the range any CKKS scale can hold (4e6 after the first XOR)."""
        e, x = self.eng, self.xor_svc
        enc_state = e.encrypt(ZetaEncoder.to_zeta(state_bytes, modulus=256))
        rows = self._row_cts(_X)
        masked, ct_mask = [], None
        for ct_x in rows:
            tb = e.relinearize(e.multiply(enc_state, ct_x, e.relin_key))
            r1, r2, r3 = e.rotate(tb, -1), e.rotate(tb, -2), e.rotate(tb, -3)
            comp = e.multiply(e.multiply(e.multiply(tb, r1), r2), r3)
            ct_mask = self._mask_ct()
            masked.append(e.multiply(comp, ct_mask, e.relin_key))
        collapsed = []
        for ct in masked:
            u1 = x.xor_cipher(ct, e.rotate(ct, -2))
            u2 = x.xor_cipher(u1, e.rotate(u1, -1))
            collapsed.append(e.relinearize(e.multiply(u2, ct_mask, e.relin_key)))     # the LAST mask, as at :71
        return collapsed

    def merged_inv_mixshift(self, enc_state):
        e, x = self.eng, self.xor_svc
        mixed = []
        for ct_x in self._row_cts(_X_INV):
            tb = e.relinearize(e.multiply(enc_state, ct_x, e.relin_key))
            comp = x.xor_cipher(tb, e.rotate(tb, -1))
            comp = x.xor_cipher(comp, e.rotate(tb, -2))
            mixed.append(x.xor_cipher(comp, e.rotate(tb, -3)))
        collapsed = []
        for ct in mixed:
            u1 = x.xor_cipher(ct, e.rotate(ct, -2))
            collapsed.append(e.relinearize(x.xor_cipher(u1, e.rotate(u1, -1))))
        return self._combine(collapsed)
