"""MixRow: the reference's "merged ShiftRows+MixColumns" operation sequence on a zeta_16
ciphertext.  Mirror of /root/reference/shift_mix_zeta.py:8-122 -- same order of engine calls
(4 x [ct x ct product, rotate -1/-2/-3, three XOR LUTs, mask product], 4 x collapse [rotate -2,
XOR, rotate -1, XOR, mask product], rotate 0/-5/-10/-15, three XOR LUTs).  SURVEY defect D8:
the sequence is not AES (products of zeta values add exponents); it is kept because config 3
names it, for operation-sequence parity against a plain-complex evaluation.  It needs ~43
levels, i.e. bootstrapping (xor_cipher refreshes operands below level 8)."""
from __future__ import annotations

import numpy as np

from .xor_service import EngineWrapper, XORService


def _zeta16(v):
    return np.exp(-2j * np.pi / 16) ** (np.asarray(v, dtype=np.int64) % 16)


_FWD = ([2, 3, 1, 1], [1, 1, 2, 3], [3, 1, 1, 2], [1, 2, 3, 1])
_INV = ([14, 11, 13, 9], [9, 14, 11, 13], [13, 9, 14, 11], [11, 13, 9, 14])


class MixRow:
    def __init__(self, xor_service: XORService, engine_wrapper: EngineWrapper):
        self.xor_svc = xor_service
        self.eng = engine_wrapper

    def _mask_ct(self):
        return self.eng.encrypt(_zeta16([1 if i % 4 == 0 else 0 for i in range(16)]))

    def _rows(self, table):
        return [self.eng.encrypt(_zeta16(np.tile(row, 4))) for row in table]

    def _b(self, ct_s, ct_x):
        e, x = self.eng, self.xor_svc
        tb = e.relinearize(e.multiply(ct_s, ct_x))
        c = x.xor_cipher(tb, e.rotate(tb, -1))
        c = x.xor_cipher(c, e.rotate(tb, -2))
        c = x.xor_cipher(c, e.rotate(tb, -3))
        return e.relinearize(e.multiply(c, self._mask_ct()))

    def _collapse(self, ct_b):
        e, x = self.eng, self.xor_svc
        u1 = x.xor_cipher(ct_b, e.rotate(ct_b, -2))
        u2 = x.xor_cipher(u1, e.rotate(u1, -1))
        return e.relinearize(e.multiply(u2, self._mask_ct()))

    def _combine(self, cts):
        out = None
        for k, ct in enumerate(cts):
            p = self.eng.rotate(ct, -(0, 5, 10, 15)[k])
            out = p if out is None else self.xor_svc.xor_cipher(out, p)
        return out

    def _run(self, ct_state, table):
        rows = self._rows(table)
        return self._combine([self._collapse(b) for b in [self._b(ct_state, r) for r in rows]])

    def merged_shift_mix_fhe(self, state_matrix):
        vec = np.array(state_matrix, dtype=np.float64).reshape(16, order="C")
        return self._run(self.eng.encrypt(_zeta16(vec)), _FWD)

    def merged_inv_mixshift_fhe_from_ct(self, ct_state):
        raw = self.eng.decrypt(self._run(ct_state, _INV))
        k = np.mod(np.rint(-np.angle(raw) * 16 / (2 * np.pi)), 16)
        return k[:16].astype(np.int64).reshape((4, 4), order="C")
