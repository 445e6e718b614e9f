"""Nibble-pair AddRoundKey end to end.  Mirror of /root/reference/new.py: split_nibbles (:38-48),
decrypt_and_recombine (:51-72), AESFHERound.encrypt_nibbles / add_round_key / full_round
(:99-109, :186-227).  The reference's own shift_rows / mix_columns in that file are broken
(SURVEY defect D7); the corrected round lives in aes_round.AESRoundService and is exposed here
under the same method names."""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np

from .xor_service import EngineWrapper, XORService, ZetaEncoder


def split_nibbles(flat: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    flat = np.asarray(flat).astype(np.uint8, copy=False)
    return np.right_shift(flat, 4), np.bitwise_and(flat, 0x0F)


def decrypt_and_recombine(ct_hi: Any, ct_lo: Any, eng: EngineWrapper, length: Optional[int] = None) -> np.ndarray:
    hi = ZetaEncoder.from_zeta(eng.decrypt(ct_hi), modulus=16)
    lo = ZetaEncoder.from_zeta(eng.decrypt(ct_lo), modulus=16)
    if length is not None:
        hi, lo = hi[..., :length], lo[..., :length]
    return (hi.astype(np.uint8) << 4) | lo.astype(np.uint8)


class AESFHERound:
    def __init__(self, eng_wrap: EngineWrapper, xor_svc: XORService):
        self.eng = eng_wrap
        self.xor = xor_svc
        self._round = None

    def _svc(self):
        if self._round is None:
            from .aes_round import AESRoundService
            self._round = AESRoundService(self.eng, self.xor)
        return self._round

    def encrypt_nibbles(self, hi: np.ndarray, lo: np.ndarray) -> Tuple[Any, Any]:
        return self.eng.encrypt(ZetaEncoder.to_zeta(hi, modulus=16)), self.eng.encrypt(ZetaEncoder.to_zeta(lo, modulus=16))

    def add_round_key(self, s_hi, s_lo, k_hi, k_lo) -> Tuple[Any, Any]:
        return self.xor.xor_cipher(s_hi, k_hi), self.xor.xor_cipher(s_lo, k_lo)

    def shift_rows(self, ct_hi, ct_lo) -> Tuple[Any, Any]:
        """2K-block packed ShiftRows (layout slot = byte_index * blocks + block)."""
        return self._svc().shift_rows((ct_hi, ct_lo))

    def mix_columns(self, planes) -> Tuple[Any, Any]:
        return self._svc().shift_rows_mix_columns(planes)

    def full_round(self, state: np.ndarray, key: np.ndarray, recombine: bool = True) -> Any:
        s_hi, s_lo = split_nibbles(state)
        k_hi, k_lo = split_nibbles(key)
        ct_s = self.encrypt_nibbles(s_hi, s_lo)
        ct_k = self.encrypt_nibbles(k_hi, k_lo)
        out_hi, out_lo = self.add_round_key(ct_s[0], ct_s[1], ct_k[0], ct_k[1])
        if not recombine:
            return out_hi, out_lo
        return decrypt_and_recombine(out_hi, out_lo, self.eng, length=np.asarray(state).shape[-1])
