"""AES-128 key schedule (FIPS-197 section 5.2), cleartext.  The reference's
/root/reference/key_expansion.py is an empty file; round keys are produced by the key owner,
zeta-encoded and encrypted like any state (new.py:205-212)."""
from __future__ import annotations

import numpy as np

from .lut import AES_SBOX

_RCON = [0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36]


def expand_key(key) -> np.ndarray:
    """16-byte key -> [11, 16] uint8 round keys (byte order = FIPS-197 state order)."""
    k = np.frombuffer(bytes(key), dtype=np.uint8) if not isinstance(key, np.ndarray) else key.astype(np.uint8)
    if k.size != 16:
        raise ValueError("AES-128 key must be 16 bytes")
    w = [k[4 * i:4 * i + 4].copy() for i in range(4)]
    for i in range(4, 44):
        t = w[i - 1].copy()
        if i % 4 == 0:
            t = np.roll(t, -1)
            t = AES_SBOX[t]
            t[0] ^= _RCON[i // 4 - 1]
        w.append(w[i - 4] ^ t)
    return np.stack([np.concatenate(w[4 * r:4 * r + 4]) for r in range(11)]).astype(np.uint8)
