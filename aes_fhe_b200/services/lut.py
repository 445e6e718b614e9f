"""LUT -> polynomial coefficient generation in the zeta domain, and the JSON formats the
reference stores them in.

A table f on Z_n is evaluated on t = zeta_n^x (zeta_n = exp(-2*pi*i/n)) by the polynomial
sum_k c_k t^k with c = ifft(zeta_m^f(x)); two-input tables use ifft2.  Same mathematics as
/root/reference/sbox/generate_sbox_coeffs.py:34-73 and
/root/reference/generator/generate_nibble_coeff.py:21-44, written from the definition;
tests/test_golden_coeffs.py pins the output against the reference's committed JSONs
(xor_mono_coeffs.json, sbox_hi/lo_coeffs.json).

JSON layouts (kept so the reference's loaders read our files and vice versa):
  1-D: {"n": n, "tol": tol, "entries": [[k, re, im], ...]}      (sbox_service.py:52-63)
  2-D: {"shape": [n, n], "tol": tol, "entries": [[i, j, re, im], ...]}   (xor_service.py:166-182)
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Dict, Tuple

import numpy as np

COEFF_DIR = Path(__file__).resolve().parent / "coeffs"

AES_SBOX = None  # filled below


def _aes_sbox() -> np.ndarray:
    """FIPS-197 section 5.1.1: multiplicative inverse in GF(2^8) then the affine map."""
    def gmul(a, b):
        r = 0
        while b:
            if b & 1:
                r ^= a
            a = ((a << 1) ^ 0x11B) if a & 0x80 else (a << 1)
            b >>= 1
        return r
    inv = [0] * 256
    for a in range(1, 256):
        for b in range(1, 256):
            if gmul(a, b) == 1:
                inv[a] = b
                break
    out = np.zeros(256, dtype=np.uint8)
    for x in range(256):
        v = inv[x]
        s = v
        for sh in range(1, 5):
            s ^= ((v << sh) | (v >> (8 - sh))) & 0xFF
        out[x] = s ^ 0x63
    return out


AES_SBOX = _aes_sbox()


def xtime(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint16)
    return (((x << 1) ^ np.where(x & 0x80, 0x11B, 0)) & 0xFF).astype(np.uint8)


def zeta(n: int) -> complex:
    return np.exp(-2j * np.pi / n)


def lut_coeffs_1d(table: np.ndarray, n_in: int, n_out: int) -> np.ndarray:
    """c with sum_k c_k zeta_in^(k x) = zeta_out^table[x] for x in Z_{n_in}."""
    vals = zeta(n_out) ** (np.asarray(table, dtype=np.int64) % n_out)
    return np.fft.ifft(vals)


def lut_coeffs_2d(table: np.ndarray, n_out: int) -> np.ndarray:
    """c with sum_ij c_ij zeta^(i a) zeta^(j b) = zeta_out^table[a, b]."""
    vals = zeta(n_out) ** (np.asarray(table, dtype=np.int64) % n_out)
    return np.fft.ifft2(vals)


def xor4_coeffs() -> np.ndarray:
    a = np.arange(16)
    return lut_coeffs_2d(a[:, None] ^ a[None, :], 16)


def sbox_hi_lo_coeffs() -> Tuple[np.ndarray, np.ndarray]:
    """Two degree-255 polynomials in t = zeta_256^x whose PRODUCT is zeta_256^S(x):
    hi(t) = zeta_256^(16*(S>>4)) = zeta_16^(S>>4),  lo(t) = zeta_256^(S & 15)
    (sbox/generate_sbox_coeffs.py:34-73; SBoxService multiplies them, sbox_service.py:114,138)."""
    s = AES_SBOX.astype(np.int64)
    return lut_coeffs_1d(s >> 4, 256, 16), lut_coeffs_1d(s & 15, 256, 256)


def save_1d(path: Path, c: np.ndarray, tol: float = 1e-12):
    entries = [[int(k), float(v.real), float(v.imag)] for k, v in enumerate(c) if abs(v) >= tol]
    Path(path).write_text(json.dumps({"n": int(len(c)), "tol": tol, "entries": entries}, indent=1))


def save_2d(path: Path, c: np.ndarray, tol: float = 1e-12):
    n0, n1 = c.shape
    entries = [[i, j, float(c[i, j].real), float(c[i, j].imag)]
               for i in range(n0) for j in range(n1) if abs(c[i, j]) >= tol]
    Path(path).write_text(json.dumps({"shape": [n0, n1], "tol": tol, "entries": entries}, indent=1))


def load_json_coeffs(path: Path) -> np.ndarray:
    """Dense 1-D coefficient array (same contract as sbox_service.py:52-63)."""
    data = json.loads(Path(path).read_text(encoding="utf-8"))
    n = data.get("n") or len(data["entries"])
    out = np.zeros(n, dtype=np.complex128)
    for k, re, im in data["entries"]:
        out[int(k)] = re + 1j * im
    return out


def load_entries(path: Path) -> Dict:
    """{k: c} for 3-field entries, {(i, j): c} for 4-field entries (xor_service.py:166-182)."""
    data = json.loads(Path(path).read_text(encoding="utf-8"))
    out: Dict = {}
    for e in data["entries"]:
        if len(e) == 3:
            out[int(e[0])] = e[1] + 1j * e[2]
        elif len(e) == 4:
            out[(int(e[0]), int(e[1]))] = e[2] + 1j * e[3]
        else:
            raise ValueError(f"Unrecognized entry format: {e}")
    return out


def ensure_default_files() -> Dict[str, Path]:
    """Write the coefficient JSONs this package ships defaults for (idempotent)."""
    COEFF_DIR.mkdir(exist_ok=True)
    files = {
        "xor": COEFF_DIR / "xor_mono_coeffs.json",
        "sbox_hi": COEFF_DIR / "sbox_hi_coeffs.json",
        "sbox_lo": COEFF_DIR / "sbox_lo_coeffs.json",
    }
    if not files["xor"].exists():
        save_2d(files["xor"], xor4_coeffs())
    if not files["sbox_hi"].exists() or not files["sbox_lo"].exists():
        hi, lo = sbox_hi_lo_coeffs()
        save_1d(files["sbox_hi"], hi)
        save_1d(files["sbox_lo"], lo)
    return files
