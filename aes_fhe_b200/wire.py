"""Wire format for ciphertexts and evaluation keys (row f-4 of SURVEY.md section 8).

The reference never serialises anything (every object lives in one Python process:
engine_context.py:62-85), so this format is new.  It exists for the multi-process deployment of
section 8(e): the key owner ships the public / relinearisation / Galois keys once, clients ship
ciphertext batches, results come back.

Layout (little endian): 8-byte magic ``AESFHE01`` | u32 kind | u32 header length | UTF-8 JSON header
| raw residues (uint64, C order, the backend's [npoly | dnum.., batch, limbs, N] layout).
The header carries the shape, the level and a digest of the parameter set (ring degree and every
modulus), so that loading into an engine with another RNS chain fails loudly instead of
decrypting garbage.  Secret keys are deliberately not serialisable through this module.
"""
from __future__ import annotations

import hashlib
import json
import struct
from typing import Union

import numpy as np

from .engine import (Ciphertext, ConjugationKey, Engine, FixedRotationKey, PublicKey, RelinearizationKey,
                     SwitchKey)

MAGIC = b"AESFHE01"
KINDS = {"ciphertext": 1, "public_key": 2, "relinearization_key": 3, "conjugation_key": 4, "fixed_rotation_key": 5}
_NAMES = {v: k for k, v in KINDS.items()}


class WireError(ValueError):
    pass


def params_digest(params) -> str:
    h = hashlib.sha256()
    h.update(struct.pack("<III", params.log_n, params.n_q, params.n_p))
    for q in params.moduli:
        h.update(struct.pack("<Q", int(q)))
    return h.hexdigest()


def _pack(kind: str, header: dict, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    header = dict(header, shape=list(arr.shape))
    hj = json.dumps(header, sort_keys=True).encode()
    return MAGIC + struct.pack("<II", KINDS[kind], len(hj)) + hj + arr.tobytes()


def _unpack(blob: Union[bytes, memoryview]):
    blob = memoryview(blob)
    if len(blob) < 16 or bytes(blob[:8]) != MAGIC:
        raise WireError("not an aes-fhe wire object (bad magic)")
    kind, hl = struct.unpack("<II", blob[8:16])
    if kind not in _NAMES or 16 + hl > len(blob):
        raise WireError("corrupt wire header")
    header = json.loads(bytes(blob[16:16 + hl]).decode())
    n = int(np.prod(header["shape"]))
    body = blob[16 + hl:]
    if len(body) != 8 * n:
        raise WireError(f"payload size {len(body)} does not match shape {header['shape']}")
    arr = np.frombuffer(body, dtype=np.uint64).reshape(header["shape"])
    return _NAMES[kind], header, arr


def _check(engine: Engine, header: dict, arr: np.ndarray, limbs_axis: int):
    if header.get("params") != params_digest(engine.params):
        raise WireError("object was produced under a different parameter set (ring degree / RNS chain)")
    if arr.shape[-1] != engine.params.n:
        raise WireError("ring degree mismatch")
    mods = engine.params.moduli
    # residues must be canonical for their limb: cheap sanity check against truncated / foreign data
    limbs = arr.shape[limbs_axis]
    ids = list(range(limbs)) if limbs <= engine.params.n_q else list(range(engine.params.n_q)) + \
        list(range(engine.params.n_q, engine.params.n_q + engine.params.n_p))
    if limbs not in (len(ids),) or limbs > len(mods):
        raise WireError("limb count does not fit the parameter set")
    top = arr.reshape(-1, limbs, arr.shape[-1]).max(axis=(0, 2))
    for j, mid in enumerate(ids):
        if int(top[j]) >= int(mods[mid]):
            raise WireError(f"residue out of range for limb {j}")


def dumps(engine: Engine, obj) -> bytes:
    be = engine.backend
    base = {"params": params_digest(engine.params)}
    if isinstance(obj, Ciphertext):
        if obj.zero:
            raise WireError("symbolic zero ciphertexts are not serialisable; add them to a real ciphertext first")
        return _pack("ciphertext", dict(base, level=obj.level), be.to_numpy(obj.polys))
    if isinstance(obj, PublicKey):
        return _pack("public_key", base, be.to_numpy(obj.polys))
    if isinstance(obj, FixedRotationKey):
        return _pack("fixed_rotation_key", dict(base, galois=int(obj.galois), delta=int(obj.delta)), be.to_numpy(obj.data))
    if isinstance(obj, ConjugationKey):
        return _pack("conjugation_key", dict(base, galois=int(obj.galois)), be.to_numpy(obj.data))
    if isinstance(obj, RelinearizationKey):
        return _pack("relinearization_key", base, be.to_numpy(obj.data))
    raise WireError(f"{type(obj).__name__} is not serialisable (secret keys never are)")


def _expect(cond: bool, what: str):
    if not cond:
        raise WireError(what)


def loads(engine: Engine, blob):
    """Untrusted input: every dimension is checked against the engine's parameter set before a raw pointer can
    reach a kernel (the kernels index [dnum][2][n_q + n_p][N] keys and [npoly][batch][level + 1][N] ciphertexts
    without size information)."""
    kind, header, arr = _unpack(blob)
    be, P = engine.backend, engine.params
    if header.get("params") != params_digest(P):
        raise WireError("object was produced under a different parameter set (ring degree / RNS chain)")
    if kind == "ciphertext":
        level = header.get("level")
        _expect(isinstance(level, int) and 0 <= level <= P.max_level, "ciphertext level outside the parameter set")
        _expect(arr.ndim == 4, "ciphertext must be [npoly, batch, level + 1, N]")
        _expect(arr.shape[0] in (2, 3), "ciphertext must have 2 or 3 polynomials")
        _expect(arr.shape[1] >= 1, "ciphertext batch must be at least 1")
        _expect(arr.shape[2] == level + 1, "ciphertext shape does not match its level")
        _check(engine, header, arr, 2)
        return Ciphertext(engine, be.from_numpy(arr), level)
    if kind == "public_key":
        _expect(arr.ndim == 4 and tuple(arr.shape[:3]) == (2, 1, P.n_q), f"public key must be [2, 1, {P.n_q}, N]")
        _check(engine, header, arr, 2)
        return PublicKey(be.from_numpy(arr))
    want = (P.dnum, 2, 1, P.n_q + P.n_p)
    _expect(arr.ndim == 5 and tuple(arr.shape[:4]) == want,
            f"switching key must be [dnum = {P.dnum}, 2, 1, n_q + n_p = {P.n_q + P.n_p}, N]")
    _check(engine, header, arr, 3)
    two_n = 2 * P.n
    if kind != "relinearization_key":
        g = header.get("galois")
        _expect(isinstance(g, int) and 0 < g < two_n and g % 2 == 1, "Galois element must be odd and below 2N")
        if kind == "conjugation_key":
            _expect(g == P.galois_conj, "conjugation key carries the wrong Galois element")
        else:
            d = header.get("delta")
            _expect(isinstance(d, int) and g == P.galois_for_rotation(d), "Galois element does not match the rotation amount")
    data = be.from_numpy(arr)
    if kind == "relinearization_key":
        return RelinearizationKey(data)
    if kind == "conjugation_key":
        return ConjugationKey(data, int(header["galois"]))
    return FixedRotationKey(data, int(header["galois"]), int(header["delta"]))
