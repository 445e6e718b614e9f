"""Block I/O helpers with the interface of /root/reference/utils.py (bytes_to_state / state_to_bytes :11-37,
zeta_encode / zeta_decode :40-59, chunk_bytes :62-67, pkcs7_pad / pkcs7_unpad :70-91), written from the
definitions (FIPS-197 section 3.4 column-major state, RFC 5652 section 6.3 padding)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def bytes_to_state(block: bytes) -> np.ndarray:
    """16 bytes -> 4 x 4 state, column-major: block[4 c + r] = state[r, c]"""
    if len(block) != 16:
        raise ValueError("Block length must be 16 bytes")
    return np.frombuffer(bytes(block), dtype=np.uint8).reshape(4, 4).T.copy()


def state_to_bytes(state: np.ndarray) -> bytes:
    state = np.asarray(state)
    if state.shape != (4, 4):
        raise ValueError("State must be a 4x4 array")
    return state.T.astype(np.uint8).tobytes()


def zeta_encode(arr: Sequence[int], modulus: int = 16) -> np.ndarray:
    a = np.asarray(arr).astype(np.int64) % modulus
    return np.exp(-2j * np.pi * a / modulus)


def zeta_decode(z: np.ndarray, modulus: int = 16) -> np.ndarray:
    k = np.rint(-np.angle(np.asarray(z)) * modulus / (2 * np.pi))
    return np.mod(k, modulus).astype(np.uint8)


def chunk_bytes(data: bytes, block_size: int = 16) -> List[bytes]:
    """consecutive blocks; the last one may be short (pad first)"""
    return [data[i:i + block_size] for i in range(0, len(data), block_size)]


def pkcs7_pad(block: bytes, block_size: int = 16) -> bytes:
    """always pads: a full extra block when the length already is a multiple of block_size"""
    pad_len = block_size - (len(block) % block_size)
    return bytes(block) + bytes([pad_len] * pad_len)


def pkcs7_unpad(data: bytes) -> bytes:
    if not data:
        return data
    pad_len = data[-1]
    if pad_len < 1 or pad_len > len(data):
        raise ValueError("Invalid padding")
    if data[-pad_len:] != bytes([pad_len] * pad_len):
        raise ValueError("Invalid PKCS#7 padding bytes")
    return data[:-pad_len]
