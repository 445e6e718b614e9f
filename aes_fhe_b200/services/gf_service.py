"""GF(2^8) x2 / x3 on a zeta_256 byte ciphertext as 1-D LUT polynomials.  Mirror of
/root/reference/gf_service.py:22-78 (GFService._eval_1d_lut / mul1 / mul2 / mul3).  The
reference loads generator/coeffs/gf{2,3}_{hi,lo}_coeffs.json, which are not in its tree
(SURVEY defect D10); the coefficients are generated here from the definition with the same
hi/lo convention as the S-box files (hi(t) = zeta_16^(y>>4), lo(t) = zeta_256^(y&15), whose
product is zeta_256^y).  `_eval_1d_lut` keeps the reference's operation order; `mul2_bsgs` /
`mul3_bsgs` use the Paterson-Stockmeyer schedule."""
from __future__ import annotations

from typing import List

import numpy as np

from . import lut
from .xor_service import EngineWrapper, XORService


def gf_tables():
    x = np.arange(256)
    x2 = lut.xtime(x).astype(np.int64)
    return x2, x2 ^ x


class GFService:
    def __init__(self, eng_wrap: EngineWrapper, xor_svc: XORService):
        self.eng = eng_wrap
        self.xor_svc = xor_svc
        t2, t3 = gf_tables()
        self.coeffs2_hi = lut.lut_coeffs_1d(t2 >> 4, 256, 16)
        self.coeffs2_lo = lut.lut_coeffs_1d(t2 & 15, 256, 256)
        self.coeffs3_hi = lut.lut_coeffs_1d(t3 >> 4, 256, 16)
        self.coeffs3_lo = lut.lut_coeffs_1d(t3 & 15, 256, 256)
        self._pts = {}

    def _plain(self, name: str) -> List:
        if name not in self._pts:
            sc = self.eng.engine.slot_count
            self._pts[name] = [self.eng.encode(np.full(sc, c, dtype=np.complex128)) for c in getattr(self, name)]
        return self._pts[name]

    def _eval_1d_lut(self, ct, pt_list):
        powers = self.eng.make_power_basis(ct, len(pt_list) - 1)
        out = self.eng.add(self.eng.multiply(ct, 0.0), pt_list[0])
        for i, pt in enumerate(pt_list[1:], start=1):
            out = self.eng.add(out, self.eng.multiply(powers[i - 1], pt))
        return out

    def mul1(self, ct):
        return ct

    def mul2(self, ct):
        return self._eval_1d_lut(ct, self._plain("coeffs2_hi")), self._eval_1d_lut(ct, self._plain("coeffs2_lo"))

    def mul3(self, ct):
        return self._eval_1d_lut(ct, self._plain("coeffs3_hi")), self._eval_1d_lut(ct, self._plain("coeffs3_lo"))

    def _bsgs(self, ct, hi, lo, key):
        from ..fused import poly_eval_bsgs
        return tuple(poly_eval_bsgs(self.eng.engine, self.eng.relin_key, ct, [hi, lo], baby=16, cache_key=key))

    def mul2_bsgs(self, ct):
        return self._bsgs(ct, self.coeffs2_hi, self.coeffs2_lo, "gf2")

    def mul3_bsgs(self, ct):
        return self._bsgs(ct, self.coeffs3_hi, self.coeffs3_lo, "gf3")
