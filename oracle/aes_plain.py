"""Plain AES-128 (FIPS-197) used as ground truth by the tests.  TEST INFRASTRUCTURE ONLY.
Written from the standard; the S-box is derived (GF(2^8) inverse + affine map), not copied."""
from __future__ import annotations

import numpy as np


def _gmul(a: int, b: int) -> int:
    r = 0
    while b:
        if b & 1:
            r ^= a
        a = ((a << 1) ^ 0x11B) if a & 0x80 else (a << 1)
        b >>= 1
    return r


def _make_sbox():
    sb = []
    for x in range(256):
        inv = 0
        if x:
            for y in range(1, 256):
                if _gmul(x, y) == 1:
                    inv = y
                    break
        s = inv
        for sh in range(1, 5):
            s ^= ((inv << sh) | (inv >> (8 - sh))) & 0xFF
        sb.append(s ^ 0x63)
    return np.array(sb, dtype=np.uint8)


SBOX = _make_sbox()
_RCON = [1, 2, 4, 8, 16, 32, 64, 128, 27, 54]


def key_schedule(key: bytes) -> np.ndarray:
    w = [list(key[4 * i:4 * i + 4]) for i in range(4)]
    for i in range(4, 44):
        t = list(w[i - 1])
        if i % 4 == 0:
            t = t[1:] + t[:1]
            t = [int(SBOX[b]) for b in t]
            t[0] ^= _RCON[i // 4 - 1]
        w.append([a ^ b for a, b in zip(w[i - 4], t)])
    return np.array([sum(w[4 * r:4 * r + 4], []) for r in range(11)], dtype=np.uint8)


def sub_bytes(s: np.ndarray) -> np.ndarray:
    return SBOX[s]


def shift_rows(s: np.ndarray) -> np.ndarray:
    """s: [..., 16] in FIPS byte order (i = 4*col + row)"""
    out = np.empty_like(s)
    for c in range(4):
        for r in range(4):
            out[..., 4 * c + r] = s[..., 4 * ((c + r) % 4) + r]
    return out


def mix_columns(s: np.ndarray) -> np.ndarray:
    out = np.empty_like(s)
    x2 = np.array([_gmul(v, 2) for v in range(256)], dtype=np.uint8)
    x3 = np.array([_gmul(v, 3) for v in range(256)], dtype=np.uint8)
    for c in range(4):
        a = [s[..., 4 * c + r] for r in range(4)]
        for r in range(4):
            out[..., 4 * c + r] = x2[a[r]] ^ x3[a[(r + 1) % 4]] ^ a[(r + 2) % 4] ^ a[(r + 3) % 4]
    return out


def round_fn(s: np.ndarray, rk: np.ndarray, last: bool = False) -> np.ndarray:
    s = shift_rows(sub_bytes(s))
    if not last:
        s = mix_columns(s)
    return s ^ rk


def encrypt_blocks(blocks: np.ndarray, key: bytes) -> np.ndarray:
    rks = key_schedule(key)
    s = np.asarray(blocks, dtype=np.uint8) ^ rks[0]
    for r in range(1, 10):
        s = round_fn(s, rks[r])
    return round_fn(s, rks[10], last=True)
