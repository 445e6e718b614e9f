"""Multi-GPU plumbing: the path shards by ciphertext batch -- no collective on the data
path.  torch.distributed (NCCL on B200s over NVLink/NVSwitch, gloo in CPU tests) does exactly two things:

* ``distribute_keys``: ONE rank owns the secret key.  It generates every key; the public key, the relinearisation
  key, the conjugation key and every Galois key the services will use (ShiftRows rotations, all baby / giant steps of
  the bootstrap's CoeffToSlot / SlotToCoeff) are broadcast once.  The other ranks hold a ``ReceivedKeys`` object in
  the place where the owner holds the secret key: it serves Galois keys by rotation amount and cannot decrypt.  The
  bootstrap's plaintext matrices are public constants (the special FFT factors); every rank derives them locally.
* ``gather_ciphertexts``: result ciphertexts (cut to the two limbs decryption reads) come back to the owner.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import (Ciphertext, ConjugationKey, Engine, FixedRotationKey, PublicKey, RelinearizationKey)


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of the ciphertext-batch index owned by `rank` (remainder spread
    over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _as_tensor(handle) -> torch.Tensor:
    if isinstance(handle, torch.Tensor):
        return handle
    return torch.from_numpy(np.ascontiguousarray(handle).view(np.int64))


def _from_tensor(t: torch.Tensor, like_numpy: bool):
    return t.numpy().view(np.uint64) if like_numpy else t


def broadcast_handle(backend, handle, src: int = 0):
    """Broadcast one backend tensor (key material) from `src`; returns the received handle."""
    t = _as_tensor(handle).contiguous()
    dist.broadcast(t, src=src)
    return _from_tensor(t, not isinstance(handle, torch.Tensor))


class ReceivedKeys:
    """What a rank that does not own the secret key holds in its place: the Galois keys that arrived, by rotation
    amount.  ``Engine.create_fixed_rotation_key(received, delta)`` returns the shipped key; anything that needs the
    secret itself (decrypt, new keys) fails."""

    def __init__(self, rotation: Dict[int, FixedRotationKey]):
        self.rotation = rotation

    def fetch_rotation_key(self, delta: int) -> FixedRotationKey:
        try:
            return self.rotation[int(delta)]
        except KeyError:
            raise RuntimeError(f"no Galois key for rotation {delta} was distributed by the key owner") from None


class EvalContext:
    """The members of EngineWrapper (xor_service.py:36-129 in the reference) the services use, on a rank that
    evaluates but cannot decrypt."""

    def __init__(self, engine: Engine, public_key, relin_key, conj_key, received: ReceivedKeys):
        self.engine = engine
        self.public_key, self.relin_key, self.conj_key = public_key, relin_key, conj_key
        self.secret_key = received
        self.owner = False

    def encrypt(self, data):
        return self.engine.encrypt(data, self.public_key)

    def decrypt(self, ct):
        raise RuntimeError("this rank does not hold the secret key: gather the ciphertext to the key owner")

    def conjugate(self, ct):
        return self.engine.conjugate(ct, self.conj_key)


def distribute_keys(engine: Engine, owner_ctx=None, src: int = 0):
    """Call on every rank with that rank's Engine (same parameter set); `owner_ctx` is the EngineWrapper /
    EngineContext holding the keys on rank `src` and None elsewhere.  Every Galois key the owner's engine has issued
    so far is shipped (run the services' ``prepare_keys`` first).  Returns ``(ctx, stats)``: the context to build
    services on (the owner's own on `src`, an EvalContext elsewhere) and
    ``{"keys", "bytes", "seconds", "gbs"}`` of the broadcast as seen by this rank."""
    rank = dist.get_rank()
    be = engine.backend
    numpy_handles = not hasattr(be, "device")
    if rank == src:
        if owner_ctx is None:
            raise ValueError("the key owner must pass its context")
        pk = owner_ctx.public_key
        rlk = getattr(owner_ctx, "relin_key", None) or owner_ctx.relinearization_key
        cj = getattr(owner_ctx, "conj_key", None) or owner_ctx.conjugation_key
        rot = dict(engine.issued_rotation_keys)
        manifest = {"pk": list(_as_tensor(pk.polys).shape), "ksk": list(_as_tensor(rlk.data).shape),
                    "rot": [(d, int(rot[d].galois)) for d in sorted(rot)], "digest": _param_digest(engine)}
        box = [manifest]
    else:
        box = [None]
    dist.broadcast_object_list(box, src=src)
    manifest = box[0]
    if manifest["digest"] != _param_digest(engine):
        raise RuntimeError("key owner and this rank use different parameter sets")

    def recv_buffer(shape):
        return be.alloc(shape)

    cuda = isinstance(getattr(be, "device", None), torch.device) and be.device.type == "cuda"
    # one small broadcast first: the collective's channels are set up on its first use, which is not key traffic
    broadcast_handle(be, be.alloc([1, 1, 1, 1, engine.params.n]) if rank != src else rlk.data[:1, :1, :, :1], src)
    if cuda:
        torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    nbytes = 0

    def ship(handle, shape):
        nonlocal nbytes
        h = handle if rank == src else recv_buffer(shape)
        out = broadcast_handle(be, h, src)
        nbytes += int(np.prod(shape)) * 8
        return out

    pk_h = ship(pk.polys if rank == src else None, manifest["pk"])
    rlk_h = ship(rlk.data if rank == src else None, manifest["ksk"])
    cj_h = ship(cj.data if rank == src else None, manifest["ksk"])
    rot_h = {}
    for d, g in manifest["rot"]:
        rot_h[d] = (ship(rot[d].data if rank == src else None, manifest["ksk"]), g)
    if cuda:
        torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    stats = {"keys": 3 + len(rot_h), "bytes": nbytes, "seconds": dt, "gbs": nbytes / dt / 1e9 if dt > 0 else None}
    if rank == src:
        owner_ctx.owner = True
        return owner_ctx, stats
    received = ReceivedKeys({d: FixedRotationKey(h, g, d) for d, (h, g) in rot_h.items()})
    return EvalContext(engine, PublicKey(pk_h), RelinearizationKey(rlk_h),
                       ConjugationKey(cj_h, engine.params.galois_conj), received), stats


def _param_digest(engine: Engine) -> str:
    import hashlib
    P = engine.params
    return hashlib.sha256(repr((P.log_n, P.max_level, P.alpha, P.moduli)).encode()).hexdigest()[:16]


def gather_ciphertexts(engine: Engine, ct: Ciphertext, dst: int = 0, limbs: Optional[int] = 2) -> Optional[List[Ciphertext]]:
    """Result ciphertexts of every rank on `dst` (None elsewhere).  Only the limbs decryption reads travel
    (`limbs`, default 2; None = all): the declared level is kept, so the owner decodes at the right scale."""
    be = engine.backend
    polys = ct.polys
    if limbs is not None and ct.level + 1 > limbs:
        polys = be.take_limbs(polys, limbs, False)
    t = _as_tensor(polys).contiguous()
    rank, world = dist.get_rank(), dist.get_world_size()
    out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst)
    if rank != dst:
        return None
    numpy_handles = not isinstance(ct.polys, torch.Tensor)
    return [Ciphertext(engine, _from_tensor(o, numpy_handles), ct.level) for o in out]
