"""ShiftRows as rotate-mask-add.  Mirror of /root/reference/shiftrows_service.py:6-69
(AESFHEShiftRows.shift_rows / inverse_shift_rows): Sum_r mask_r (.) rot(ct, -/+4r) on a
16-slot column-major state.  As SURVEY defect D6 documents, on a 2^k-slot ring this loses the
bytes that wrap inside the block; `packed=True` applies the same pattern to the 2048-block
layout (slot = byte * blocks + block, rotations scaled by `blocks`), where it is exact."""
from __future__ import annotations

import numpy as np

from .xor_service import EngineWrapper, XORService


class AESFHEShiftRows:
    def __init__(self, engine_wrapper: EngineWrapper, xor_svc: XORService, packed: bool = False):
        self.eng = engine_wrapper
        self.xor_svc = xor_svc
        sc = self.eng.engine.slot_count
        self.stride = sc // 16 if packed else 1
        self.row_rot = [0, -4, -8, -12]
        self.masks = []
        for r in range(4):
            m = np.zeros(16, dtype=float)
            m[r::4] = 1.0
            full = np.repeat(m, self.stride) if packed else np.pad(m, (0, sc - 16))
            self.masks.append(self.eng.encode(full))

    def _apply(self, ct, rots):
        out = None
        for r in range(4):
            part = self.eng.multiply(ct, self.masks[r])
            if rots[r] != 0:
                part = self.eng.rotate(part, rots[r] * self.stride)
            out = part if out is None else self.eng.add(out, part)
        return out

    def shift_rows(self, ct):
        return self._apply(ct, self.row_rot)

    def inverse_shift_rows(self, ct):
        return self._apply(ct, [0, 4, 8, 12])
