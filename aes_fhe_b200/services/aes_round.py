"""A FIPS-197-correct homomorphic AES round assembled from the reference's sound primitives
(SURVEY.md section 7.3), batched over ciphertexts.

Layout: one ciphertext packs B = slot_count/16 blocks, byte i of block b in slot i*B + b (the
reference's "2K blocks per ciphertext", README.md:29, new.py:19-20).  A rotation by a multiple
of B is then a true cyclic shift of the 16 byte positions, which is what makes the reference's
rotate-mask-add ShiftRows (shiftrows_service.py:33-51) exact instead of losing the wrapped bytes
(SURVEY defect D6).  The state is a nibble pair (ct_hi, ct_lo) of zeta_16-valued ciphertexts as
in new.py:199-216.

  AddRoundKey   : two 4-bit XOR LUTs (xor_service.py:271-286, fused schedule)         5 levels
  SubBytes      : six bivariate 16x16 LUTs on shared power bases -> nibble planes of
                  S(x), 2*S(x), 3*S(x)  (generator/generate_nibble_coeff.py:33-44 style)  5 levels
  ShiftRows+Mix : out[r,c] = 2S[r,c+r] ^ 3S[r+1,c+r+1] ^ S[r+2,..] ^ S[r+3,..]; each operand
                  is sum_o mask_o (.) rot(plane, -o*B)  (rotate-mask-add), then three XORs   1 + 10 levels
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..fused import bivariate_lut, power_basis_16
from . import lut
from .xor_service import XORService, ZetaEncoder


def _nibble_tables() -> Dict[str, np.ndarray]:
    s = lut.AES_SBOX.astype(np.int64)
    s2 = lut.xtime(s).astype(np.int64)
    s3 = s2 ^ s
    out = {}
    for name, t in (("S", s), ("2S", s2), ("3S", s3)):
        out[name + "_hi"] = lut.lut_coeffs_2d((t >> 4).reshape(16, 16), 16)
        out[name + "_lo"] = lut.lut_coeffs_2d((t & 15).reshape(16, 16), 16)
    return out


class AESRoundService:
    PLANES = ("S_hi", "S_lo", "2S_hi", "2S_lo", "3S_hi", "3S_lo")

    def __init__(self, eng_wrap, xor_svc: XORService):
        self.eng = eng_wrap
        self.xor = xor_svc
        self.engine = eng_wrap.engine
        self.sc = self.engine.slot_count
        self.B = self.sc // 16
        self.tables = _nibble_tables()
        self._rot_keys: Dict[int, object] = {}
        self._masks: Dict[Tuple[int, ...], object] = {}
        xc = np.zeros((16, 16), dtype=np.complex128)
        for (i, j), c in xor_svc.coeff_cache.load_coeffs().items():
            xc[i, j] = c
        self._xor_coeffs = xc

    # ------------------------------------------------------------------ packing
    def pack(self, blocks: np.ndarray) -> np.ndarray:
        """[nb <= B, 16] bytes -> flat [slot_count] with slot = i*B + b (unused blocks are 0)."""
        blocks = np.asarray(blocks, dtype=np.uint8).reshape(-1, 16)
        if blocks.shape[0] > self.B:
            raise ValueError(f"at most {self.B} blocks per ciphertext")
        flat = np.zeros((16, self.B), dtype=np.uint8)
        flat[:, :blocks.shape[0]] = blocks.T
        return flat.reshape(-1)

    def unpack(self, flat: np.ndarray, nb: Optional[int] = None) -> np.ndarray:
        out = np.asarray(flat, dtype=np.uint8).reshape(16, self.B).T
        return out if nb is None else out[:nb]

    def encrypt_bytes(self, flat: np.ndarray):
        """flat [slot_count] or [batch, slot_count] bytes -> (ct_hi, ct_lo)."""
        flat = np.asarray(flat, dtype=np.uint8)
        return (self.eng.encrypt(ZetaEncoder.to_zeta(flat >> 4)), self.eng.encrypt(ZetaEncoder.to_zeta(flat & 15)))

    def encrypt_state(self, blocks):
        """blocks: [nb,16] (one ciphertext) or a list of such arrays (a batch)."""
        if isinstance(blocks, (list, tuple)):
            return self.encrypt_bytes(np.stack([self.pack(b) for b in blocks]))
        return self.encrypt_bytes(self.pack(blocks))

    def encrypt_round_key(self, rk16: np.ndarray):
        """one 16-byte round key replicated over every block slot"""
        rk = np.asarray(rk16, dtype=np.uint8).reshape(16)
        return self.encrypt_bytes(np.repeat(rk, self.B))

    def decrypt_bytes(self, state) -> np.ndarray:
        hi = ZetaEncoder.from_zeta(self.eng.decrypt(state[0]))
        lo = ZetaEncoder.from_zeta(self.eng.decrypt(state[1]))
        return (hi.astype(np.uint8) << 4) | lo.astype(np.uint8)

    def decrypt_state(self, state, nb: Optional[int] = None) -> np.ndarray:
        flat = self.decrypt_bytes(state)
        if flat.ndim == 2:
            return np.stack([self.unpack(f, nb) for f in flat])
        return self.unpack(flat, nb)

    # ------------------------------------------------------------------ primitives
    def _xor(self, a, b):
        return bivariate_lut(self.eng, a, b, [self._xor_coeffs], cache_key="xor4")[0]

    def add_round_key(self, state, key):
        return self._xor(state[0], key[0]), self._xor(state[1], key[1])

    def sbox_planes(self, state, names=PLANES) -> Dict[str, object]:
        outs = bivariate_lut(self.eng, state[0], state[1], [self.tables[n] for n in names],
                             cache_key=("sbox16",) + tuple(names))
        return dict(zip(names, outs))

    def sub_bytes(self, state):
        p = self.sbox_planes(state, ("S_hi", "S_lo"))
        return p["S_hi"], p["S_lo"]

    def _rot(self, ct, offset: int):
        """bring byte position i+offset to position i (cyclic over the 16 positions)"""
        offset %= 16
        if offset == 0:
            return ct
        key = self._rot_keys.get(offset)
        if key is None:
            key = self.engine.create_fixed_rotation_key(self.eng.secret_key, -offset * self.B)
            self._rot_keys[offset] = key
        return self.engine.rotate(ct, key)

    def _mask(self, rows: Tuple[int, ...]):
        pt = self._masks.get(rows)
        if pt is None:
            m = np.zeros((16, self.B), dtype=np.float64)
            for i in range(16):
                if i % 4 in rows:
                    m[i] = 1.0
            pt = self.engine.encode(m.reshape(-1))
            self._masks[rows] = pt
        return pt

    def _rot_key(self, offset: int):
        key = self._rot_keys.get(offset)
        if key is None:
            key = self.engine.create_fixed_rotation_key(self.eng.secret_key, -offset * self.B)
            self._rot_keys[offset] = key
        return key

    @staticmethod
    def _gather_plan(k: int) -> Dict[int, List[int]]:
        """T_k[r, c] = plane[(r+k)%4, (c + (r+k)%4) % 4]: source byte position of target i = 4c + r is
        i + (5 r' - r), r' = (r+k) % 4.  Returns {rotation offset: rows it serves}."""
        by_off: Dict[int, List[int]] = {}
        for r in range(4):
            rp = (r + k) % 4
            by_off.setdefault((5 * rp - r) % 16, []).append(r)
        return by_off

    def _gather_many(self, plane, ks: Sequence[int]):
        """rotate-mask-add gathers T_k of one plane for several k: all rotations of the plane share
        one ModUp (Engine.rotate_hoisted), each gather is one fused multiply-accumulate."""
        plans = {k: self._gather_plan(k) for k in ks}
        offs = sorted({o for p in plans.values() for o in p if o})
        rots = dict(zip(offs, self.engine.rotate_hoisted(plane, [self._rot_key(o) for o in offs])))
        rots[0] = plane
        out = []
        for k in ks:
            items = sorted(plans[k].items())
            out.append(self.engine.multiply_plain_sum([rots[o] for o, _ in items], [self._mask(tuple(r)) for _, r in items]))
        return out

    def _gather(self, plane, k: int):
        return self._gather_many(plane, [k])[0]

    def shift_rows(self, state):
        """ShiftRows alone (k = 0 gather), the corrected form of shiftrows_service.shift_rows."""
        return self._gather(state[0], 0), self._gather(state[1], 0)

    def shift_rows_mix_columns(self, planes: Dict[str, object]):
        out = []
        for nib in ("hi", "lo"):
            t0 = self._gather(planes["2S_" + nib], 0)
            t1 = self._gather(planes["3S_" + nib], 1)
            t2, t3 = self._gather_many(planes["S_" + nib], [2, 3])
            out.append(self._xor(self._xor(t0, t1), self._xor(t2, t3)))
        return out[0], out[1]

    # ------------------------------------------------------------------ rounds
    def round(self, state, round_key, last: bool = False):
        """SubBytes, ShiftRows, (MixColumns), AddRoundKey on a nibble-pair state."""
        if last:
            s = self.sub_bytes(state)
            s = self.shift_rows(s)
        else:
            s = self.shift_rows_mix_columns(self.sbox_planes(state))
        return self.add_round_key(s, round_key)
