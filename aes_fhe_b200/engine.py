"""`desilofhe`-compatible CKKS engine facade (host logic only).

This is the drop-in boundary of the hot path: the reference's Python services call
exactly these members of ``desilofhe.Engine`` (call sites
/root/reference/engine_context.py:28-85, xor_service.py:59-129, sbox/sbox_service.py:85-138,
gf_service.py:44-64, new.py:94-146, shiftrows_service.py:30-68).  The facade keeps the
level / scale bookkeeping, key objects and operand normalisation in Python and hands
every polynomial operation to a *backend*:

* product: :class:`aes_fhe_b200.backend_cuda.CudaBackend` -- hand-written sm_100a
  kernels behind the C-ABI in ``include/aesfhe_b200.h``.  It is the only backend this
  package ships and the default; constructing an Engine without the built CUDA library
  or without a GPU raises.
* tests / CPU baseline only: ``oracle.refmod.RefBackend`` is injected explicitly by
  ``tests/`` and ``bench.py`` to obtain the bit-exact residues the kernels must match.

Ciphertext layout: ``polys`` is ``[npoly, batch, level+1, N]`` 64-bit residues in the NTT
(evaluation) domain, bit-reversed spectrum order, canonical [0, q).  ``batch`` independent
ciphertexts (each packing slot_count/16 AES blocks) move through every operation together;
the reference's one-ciphertext calls are the batch = 1 case.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import encoding
from .params import CKKSParams, LOG_PQ_BUDGET_DENSE, LOG_PQ_BUDGET_SPARSE, make_params, sqrt_minus_one

_SIGMA = 3.2


import os as _os
_NO_FUSE = _os.environ.get("FHE_NO_HOIST") == "1"      # A/B switch: separate rotations, mul + add per mask


class SystemRNG:
    """The subset of numpy.random.Generator the engine uses, drawn from ``os.urandom`` (the kernel CSPRNG):
    uniform integers by rejection on 64-bit words (no modulo bias), Gaussians by Box-Muller on 53-bit uniforms,
    choice without replacement by sorting random keys."""

    @staticmethod
    def _words(count: int) -> np.ndarray:
        return np.frombuffer(_os.urandom(8 * count), dtype=np.uint64)

    def integers(self, low, high=None, size=None, dtype=np.int64):
        if high is None:
            low, high = 0, low
        low, span = int(low), int(high) - int(low)
        if span <= 0:
            raise ValueError("empty range")
        shape = () if size is None else (tuple(size) if np.iterable(size) else (int(size),))
        count = int(np.prod(shape)) if shape else 1
        limit = (1 << 64) - ((1 << 64) % span)                    # accept words below the largest multiple of span
        out = np.empty(count, dtype=np.uint64)
        filled = 0
        while filled < count:
            w = self._words(count - filled + 16)
            if limit < (1 << 64):
                w = w[w < np.uint64(limit)]
            take = min(w.size, count - filled)
            out[filled:filled + take] = w[:take] % np.uint64(span)
            filled += take
        res = (out.astype(np.int64) + low) if np.dtype(dtype) != np.uint64 else (out + np.uint64(low))
        res = res.astype(dtype, copy=False).reshape(shape)
        return res if shape else res[()]

    def _uniform(self, count: int) -> np.ndarray:
        return ((self._words(count) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / (1 << 53))      # (0, 1)

    def normal(self, loc=0.0, scale=1.0, size=None):
        shape = () if size is None else (tuple(size) if np.iterable(size) else (int(size),))
        count = int(np.prod(shape)) if shape else 1
        half = (count + 1) // 2
        r = np.sqrt(-2.0 * np.log(self._uniform(half)))
        t = 2.0 * np.pi * self._uniform(half)
        z = np.concatenate([r * np.cos(t), r * np.sin(t)])[:count]
        return (loc + scale * z).reshape(shape) if shape else float(loc + scale * z[0])

    def choice(self, n: int, size: int, replace: bool = False):
        if replace:
            return self.integers(0, n, size=size)
        return np.argsort(self._words(int(n)), kind="stable")[:int(size)]


class Plaintext:
    """Level-agnostic plaintext: slot values plus a per-level cache of encodings.

    The reference encodes a plaintext once (``engine.encode``) and then multiplies it
    into ciphertexts at many different levels (/root/reference/xor_service.py:184-196,
    :283-285), so the residues are produced lazily for the level they are used at.
    """

    def __init__(self, engine: "Engine", values: np.ndarray, scale_fn=None):
        self.engine = engine
        self.scale_fn = scale_fn            # level -> scale override (bootstrapping's first matrix)
        v = np.asarray(values)
        if v.ndim == 0:
            v = v.reshape(1)
        self.values = v.astype(np.complex128 if np.iscomplexobj(v) else np.float64).ravel()
        self.const = encoding.as_constant(self.values) if self.values.size == engine.slot_count else None
        self._cache: Dict[int, Any] = {}

    def at_level(self, level: int, ext: bool = False):
        """residues at `level`; ``ext``: in the extended basis Q u P (limbs 0..level followed by the special primes),
        for products with key-switch accumulators that have not been divided by P yet (double hoisting)"""
        h = self._cache.get((level, ext))
        if h is None:
            eng = self.engine
            scale = eng.params.scale(level) if self.scale_fn is None else float(self.scale_fn(level))
            if scale < 2.0 ** 58:
                coeffs = encoding.encode_i64(self.values, scale, eng.params.log_n)
                h = eng.backend.from_i64(coeffs, level + 1, ext)
            elif ext:
                raise RuntimeError("extended-basis encoding is not available at this scale")
            else:
                # very high scale (first bootstrapping matrix): coefficient = hi * 2^40 + lo, both
                # exact in int64, recombined on the residues
                m = encoding.slots_to_coeffs(self.values, eng.params.log_n) * (scale / 2.0 ** 40)
                hi = np.floor(m)
                lo = np.rint((m - hi) * 2.0 ** 40)
                be, nq = eng.backend, level + 1
                fac = [(1 << 40) % eng.params.moduli[l] for l in range(nq)]
                h = be.add(be.mul_scalar(be.from_i64(hi.astype(np.int64), nq, False), fac, nq, 0),
                           be.from_i64(lo.astype(np.int64), nq, False), nq, 0)
            self._cache[(level, ext)] = h
        return h


class Ciphertext:
    def __init__(self, engine: "Engine", polys, level: int, zero: bool = False):
        self.engine = engine
        self.polys = polys
        self._level = level
        self.zero = zero

    @property
    def level(self) -> int:
        """Remaining multiplicative levels (the reference compares it with 8:
        /root/reference/xor_service.py:274-277)."""
        return self._level

    @property
    def npoly(self) -> int:
        return self.engine.backend.npoly(self.polys)

    @property
    def batch(self) -> int:
        return self.engine.backend.batch(self.polys)

    def __repr__(self):
        return f"Ciphertext(level={self._level}, npoly={self.npoly}, batch={self.batch})"


class SecretKey:
    def __init__(self, coeffs: np.ndarray, ntt_full):
        self.coeffs = coeffs          # int64 ternary, length N
        self.ntt = ntt_full           # [1, 1, n_q+n_p, N]


class PublicKey:
    def __init__(self, polys):
        self.polys = polys            # [2, 1, n_q, N]


class SwitchKey:
    """Hybrid key-switching key  [dnum, 2, 1, n_q+n_p, N]."""

    def __init__(self, data, galois: Optional[int] = None):
        self.data = data
        self.galois = galois


class RelinearizationKey(SwitchKey):
    pass


class ConjugationKey(SwitchKey):
    pass


class FixedRotationKey(SwitchKey):
    def __init__(self, data, galois, delta):
        super().__init__(data, galois)
        self.delta = delta


class RotationKey:
    """Power-of-two rotation keys (+-2^k); arbitrary deltas are composed in NAF form."""

    def __init__(self, keys: Dict[int, FixedRotationKey]):
        self.keys = keys


class BootstrapKey:
    """Bootstrapping key: Galois keys for the rotations of CoeffToSlot / SlotToCoeff plus the
    encoded matrices (filled by aes_fhe_b200.bootstrap.make_bootstrap_key).  The "small"
    variant of the reference API (engine_context.py:72) is an empty token."""

    def __init__(self, small: bool):
        self.small = small
        self.plan = None


def _naf_steps(delta: int, slot_count: int) -> List[int]:
    """Signed power-of-two decomposition of a rotation amount (shortest of +-)."""
    d = delta % slot_count
    if d > slot_count // 2:
        d -= slot_count
    steps, k, x = [], 0, abs(d)
    sign = 1 if d >= 0 else -1
    while x:
        if x & 1:
            digit = 2 - (x & 3)          # +1 or -1
            steps.append(sign * digit * (1 << k))
            x -= digit
        x >>= 1
        k += 1
    return steps


class Engine:
    """Same three constructor signatures as ``desilofhe.Engine``
    (/root/reference/engine_context.py:28-56)."""

    def __init__(self, *args, mode: str = 'cpu', use_bootstrap: bool = False,
                 use_multiparty: bool = False, thread_count: int = 0, device_id: int = 0,
                 max_level: Optional[int] = None, log_coeff_count: Optional[int] = None,
                 special_prime_count: Optional[int] = None, seed: Optional[int] = None, device_codec: bool = False,
                 secret_hamming_weight: Optional[int] = None, scale_bits: Optional[int] = None,
                 _backend=None, _params: Optional[CKKSParams] = None):
        args = list(args)
        # positional forms: Engine(mode) | Engine(max_level, mode) | Engine(log_n, K, mode)
        if len(args) >= 1 and isinstance(args[0], str):
            mode = args.pop(0)
        ints = [a for a in args if isinstance(a, (int, np.integer))]
        strs = [a for a in args if isinstance(a, str)]
        if strs:
            mode = strs[0]
        if len(ints) == 1 and max_level is None:
            max_level = int(ints[0])
        elif len(ints) >= 2:
            log_coeff_count, special_prime_count = int(ints[0]), int(ints[1])
        if use_multiparty:
            raise NotImplementedError("multiparty CKKS is outside the hot path (SURVEY.md section 8)")
        if mode not in ('cpu', 'parallel', 'gpu'):
            raise ValueError(f"unknown mode {mode!r}")
        self.mode = mode
        # device_codec: encode / sample / decode on the GPU (throughput path).  Off by default so
        # that ciphertexts are reproducible bit for bit from the host RNG (parity tests).
        self.device_codec = bool(device_codec)
        self.use_bootstrap = bool(use_bootstrap)
        self.thread_count = thread_count
        self.device_id = device_id

        if _params is not None:
            params = _params
        elif log_coeff_count is not None:
            log_n = int(log_coeff_count)
            k = int(special_prime_count or 0)
            # depth the modulus budget allows at this ring size (about 27*2^(log_n-10) bits)
            budget = 27 * (1 << (log_n - 10)) + 16
            lvl = max(1, min(30, (budget - 60 - 60 * max(k, 1)) // 40))
            params = make_params(log_n, lvl, special_count=k)
        else:
            # a bootstrappable engine works at the largest scale the < 2^45 moduli allow (every operation is
            # 16x more precise) and has a sparse secret, so its chain is sized against the sparse-secret bound:
            # 26 levels: a bit bootstrap (13) leaves 13 = one AES round (7) + the entry of the next bootstrap (4)
            # + 2 to spare, which lets three rounds run on a fresh input and the last round ride on the ninth's
            # refresh (six refreshes per AES-128; 24 levels: eight); log2(P Q) = 1504 with four key-switch digits of
            # seven.  Other engines: 30 levels, uniform ternary secret.
            sb = scale_bits if scale_bits is not None else (44 if use_bootstrap else 40)
            lvl = int(max_level) if max_level is not None else (26 if use_bootstrap else 30)
            params = make_params(16, lvl, scale_bits=sb,
                                 log_pq_budget=LOG_PQ_BUDGET_SPARSE if use_bootstrap else LOG_PQ_BUDGET_DENSE)
        self.params = params
        # Secret distribution.  Bootstrappable engines: sparse ternary, Hamming weight 192 (bounds the overflow
        # polynomial of ModRaise: |I| <= 32).  All others: uniform ternary -- the distribution the 1770-bit bound is
        # stated for.  ``secret_hamming_weight`` overrides (0 = uniform ternary).
        if secret_hamming_weight is None:
            secret_hamming_weight = min(192, params.n // 8) if use_bootstrap else 0
        self.secret_hamming_weight = int(secret_hamming_weight)
        self.slot_count = params.slot_count
        self.max_level = params.max_level

        if _backend is None:
            from .backend_cuda import CudaBackend      # raises if the CUDA library / GPU is missing
            _backend = CudaBackend(params, device_id=device_id)
        elif isinstance(_backend, type):
            _backend = _backend(params)
        self.backend = _backend
        # Randomness.  seed=None (the default, and what every reference call site gets: desilofhe.Engine takes no
        # seed) draws every key, mask and error from the operating system's CSPRNG.  An integer seed makes keys and
        # ciphertexts reproducible for the parity tests and is NOT secure: anyone can regenerate the secret key.
        self.seeded = seed is not None
        self._rng = np.random.Generator(np.random.PCG64(seed)) if self.seeded else SystemRNG()
        # per-limb sqrt(-1) (NTT image of X^(N/2))
        self._imag = [sqrt_minus_one(params, l) for l in range(len(params.moduli))]
        self.op_counts: Dict[str, int] = {}
        self.issued_rotation_keys: Dict[int, FixedRotationKey] = {}

    # ------------------------------------------------------------------ utils
    def _count(self, name: str, k: int = 1):
        self.op_counts[name] = self.op_counts.get(name, 0) + k

    def _sample_ternary(self, batch: int = 1) -> np.ndarray:
        return self._rng.integers(-1, 2, size=(batch, self.params.n), dtype=np.int64)

    def _sample_error(self, batch: int = 1) -> np.ndarray:
        return np.rint(self._rng.normal(0.0, _SIGMA, size=(batch, self.params.n))).astype(np.int64)

    def _sample_uniform(self, limb_ids: Sequence[int]) -> np.ndarray:
        """[1, 1, len(limb_ids), N] uniform residues"""
        out = np.empty((1, 1, len(limb_ids), self.params.n), dtype=np.uint64)
        for r, l in enumerate(limb_ids):
            out[0, 0, r] = self._rng.integers(0, self.params.moduli[l], size=self.params.n, dtype=np.uint64)
        return out

    def _const_residues(self, re: int, im: int, n_q_active: int):
        """(c_plus, c_minus) per active limb for the polynomial re + im*X^(N/2)."""
        cp, cm = [], []
        for l in range(n_q_active):
            q = self.params.moduli[l]
            t = (im % q) * self._imag[l] % q
            cp.append((re + t) % q)
            cm.append((re - t) % q)
        return cp, cm

    # ------------------------------------------------------------------ keys
    def create_secret_key(self) -> SecretKey:
        n, h = self.params.n, self.secret_hamming_weight
        if h <= 0:
            s = self._rng.integers(-1, 2, size=(1, n), dtype=np.int64)          # uniform ternary
        else:
            s = np.zeros((1, n), dtype=np.int64)
            pos = self._rng.choice(n, size=h, replace=False)
            s[0, pos] = self._rng.integers(0, 2, size=h, dtype=np.int64) * 2 - 1
        return SecretKey(s, self.backend.from_i64(s, self.params.n_q, True))

    @property
    def security(self) -> Dict[str, Any]:
        """log2(P Q) against the cited 128-bit bound for this engine's secret distribution (params.py); only
        meaningful at N = 2^16 -- smaller rings are test rings."""
        sparse = self.secret_hamming_weight > 0
        budget = LOG_PQ_BUDGET_SPARSE if sparse else LOG_PQ_BUDGET_DENSE
        lp = self.params.log_pq
        return {"log_n": self.params.log_n, "log_pq": round(lp, 1), "budget": budget,
                "secret": f"sparse ternary h={self.secret_hamming_weight}" if sparse else "uniform ternary",
                "within_128_bit_budget": bool(self.params.log_n == 16 and lp <= budget),
                "randomness": "seeded PCG64 (test only, insecure)" if self.seeded else "os.urandom"}

    def create_public_key(self, sk: SecretKey) -> PublicKey:
        be, P = self.backend, self.params
        nq = P.n_q
        a = be.from_numpy(self._sample_uniform(range(nq)))
        e = be.from_i64(self._sample_error(), nq, False)
        s = be.take_limbs(sk.ntt, nq, False)
        b = be.sub(e, be.mul(a, s, nq, 0), nq, 0)
        return PublicKey(be.concat([b, a]))

    def _make_switch_key(self, sk: SecretKey, s_from) -> Any:
        """KSK for s_from -> s.  s_from: [1, 1, n_q+n_p, N] NTT."""
        be, P = self.backend, self.params
        nq, npp = P.n_q, P.n_p
        tot = nq + npp
        ids = list(range(tot))
        p_prod = 1
        for p in P.p:
            p_prod *= p
        parts = []
        for j in range(P.dnum):
            lo, hi = j * P.alpha, min((j + 1) * P.alpha, nq)
            a = be.from_numpy(self._sample_uniform(ids))
            e = be.from_i64(self._sample_error(), nq, True)
            b = be.sub(e, be.mul(a, sk.ntt, nq, npp), nq, npp)
            # + (P mod q_i) * s_from on the digit's own limbs
            fac = [(p_prod % P.moduli[l]) if lo <= l < hi else 0 for l in ids]
            b = be.add(b, be.mul_scalar(s_from, fac, nq, npp), nq, npp)
            parts.append(be.concat([b, a]))
        return be.stack(parts)

    def create_relinearization_key(self, sk: SecretKey) -> RelinearizationKey:
        P = self.params
        s2 = self.backend.mul(sk.ntt, sk.ntt, P.n_q, P.n_p)
        return RelinearizationKey(self._make_switch_key(sk, s2))

    def _galois_key(self, sk: SecretKey, g: int):
        P = self.params
        sg = self.backend.automorphism(sk.ntt, g, P.n_q, P.n_p)
        return self._make_switch_key(sk, sg)

    def create_conjugation_key(self, sk: SecretKey) -> ConjugationKey:
        g = self.params.galois_conj
        return ConjugationKey(self._galois_key(sk, g), g)

    def create_fixed_rotation_key(self, sk: SecretKey, delta: int) -> FixedRotationKey:
        if hasattr(sk, "fetch_rotation_key"):
            # a rank that evaluates without the secret key: the key owner shipped this key (sharding.ReceivedKeys)
            return sk.fetch_rotation_key(int(delta))
        g = self.params.galois_for_rotation(int(delta))
        key = FixedRotationKey(self._galois_key(sk, g), g, int(delta))
        self.issued_rotation_keys[int(delta)] = key          # what distribute_keys ships to the other ranks
        return key

    def create_rotation_key(self, sk: SecretKey, steps: Optional[Sequence[int]] = None) -> RotationKey:
        """All +-2^k steps by default (the reference passes only ``sk``:
        /root/reference/engine_context.py:66); ``steps`` restricts the set."""
        if steps is None:
            steps = []
            k = 1
            while k < self.slot_count:
                steps += [k, -k]
                k <<= 1
        return RotationKey({int(d): self.create_fixed_rotation_key(sk, int(d)) for d in steps})

    # ------------------------------------------------------------------ encode / encrypt / decrypt
    def encode(self, values) -> Plaintext:
        v = np.asarray(values)
        if v.size > self.slot_count:
            raise ValueError(f"encode: {v.size} values exceed slot_count={self.slot_count}")
        return Plaintext(self, v)

    def encrypt(self, data, public_key: PublicKey, level: Optional[int] = None) -> Ciphertext:
        """1-D data -> one ciphertext; 2-D data [B, <= slot_count] -> a batch of B ciphertexts."""
        be, P = self.backend, self.params
        if isinstance(data, Plaintext):
            data = data.values
        lvl = P.max_level if level is None else int(level)
        nq = lvl + 1
        if self.device_codec and hasattr(be, "encode_device"):
            return self._encrypt_device(data, public_key, lvl)
        data = np.asarray(data)
        rows = data.reshape(1, -1) if data.ndim <= 1 else data
        bt = rows.shape[0]
        coeffs = np.stack([encoding.encode_i64(r, P.scale(lvl), P.log_n) for r in rows])
        m = be.from_i64(coeffs, nq, False)
        v = be.from_i64(self._sample_ternary(bt), nq, False)
        e0 = be.from_i64(self._sample_error(bt), nq, False)
        e1 = be.from_i64(self._sample_error(bt), nq, False)
        pk = public_key.polys if nq == P.n_q else be.take_limbs(public_key.polys, nq, False)
        vb = be.mul(pk, v, nq, 0)                      # [2,1,nq,N] * [1,B,nq,N]
        c = be.add(vb, be.concat([be.add(e0, m, nq, 0), e1]), nq, 0)
        self._count('encrypt', bt)
        return Ciphertext(self, c, lvl)

    def _encrypt_device(self, data, public_key: PublicKey, lvl: int) -> Ciphertext:
        """Throughput path: canonical embedding (cuFFT through torch), rounding and the
        ternary / Gaussian sampling all on the device; one H2D copy of the slot values."""
        import torch
        be, P = self.backend, self.params
        nq = lvl + 1
        x = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.complex128)))
        x = x.reshape(1, -1) if x.ndim <= 1 else x
        x = x.to(be.device, non_blocking=True)
        bt = x.shape[0]
        if x.shape[1] < self.slot_count:
            x = torch.nn.functional.pad(x, (0, self.slot_count - x.shape[1]))
        m = be.from_i64_device(be.encode_device(x, P.scale(lvl)), nq, False)
        g = self._torch_gen()
        v = be.from_i64_device(torch.randint(-1, 2, (bt, P.n), generator=g, device=be.device, dtype=torch.int64), nq, False)
        e = torch.round(torch.randn((2, bt, P.n), generator=g, device=be.device, dtype=torch.float64) * _SIGMA).to(torch.int64)
        e0 = be.from_i64_device(e[0], nq, False)
        e1 = be.from_i64_device(e[1], nq, False)
        pk = public_key.polys if nq == P.n_q else be.take_limbs(public_key.polys, nq, False)
        vb = be.mul(pk, v, nq, 0)
        c = be.add(vb, be.concat([be.add(e0, m, nq, 0), e1]), nq, 0)
        self._count('encrypt', bt)
        return Ciphertext(self, c, lvl)

    def _torch_gen(self):
        g = getattr(self, "_tgen", None)
        if g is None:
            import torch
            g = torch.Generator(device=self.backend.device)
            g.manual_seed(int(self._rng.integers(0, 2 ** 62)))            # OS entropy unless the engine is seeded
            self._tgen = g
        return g

    def decrypt_device(self, ct: Ciphertext, sk: SecretKey):
        """Slots as a complex128 torch tensor [batch, slot_count] on the device."""
        be = self.backend
        use = min(2, ct.level + 1)
        c = be.take_limbs(ct.polys, use, False)
        s = be.take_limbs(sk.ntt, use, False)
        acc = be.add(be.select_poly(c, 0), be.mul(be.select_poly(c, 1), s, use, 0), use, 0)
        if ct.npoly == 3:
            acc = be.add(acc, be.mul(be.select_poly(c, 2), be.mul(s, s, use, 0), use, 0), use, 0)
        acc = be.intt(acc, use, 0)
        self._count('decrypt', ct.batch)
        return be.decode_device(acc, use, self.params.scale(ct.level))

    # ---- byte-level codec on the device (the reference's ZetaEncoder.to_zeta / from_zeta around
    #      encrypt / decrypt: xor_service.py:132-145, 318-328), so a step moves bytes, not complex128
    def encrypt_zeta(self, values, public_key: PublicKey, modulus: int = 16, level: Optional[int] = None,
                     amplitude: float = 1.0) -> Ciphertext:
        """uint8 values [<= slot_count] or [B, <= slot_count] -> ciphertext(s) of zeta_m^x = exp(-2 pi i x / m)
        (times `amplitude`); one H2D copy of the bytes, everything else on the device.  Same result as
        ``encrypt(ZetaEncoder.to_zeta(values, modulus), pk)`` up to the encryption randomness."""
        import torch
        be = self.backend
        if not hasattr(be, "encode_device"):
            ang = -2.0 * np.pi * (np.asarray(values).astype(np.int64) % modulus) / modulus
            return self.encrypt(amplitude * np.exp(1j * ang), public_key, level)
        x = values if isinstance(values, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(values, dtype=np.uint8)))
        x = x.reshape(1, -1) if x.ndim <= 1 else x
        x = x.to(be.device, non_blocking=True)
        n_in = x.shape[1]
        ang = (x.to(torch.float64) % modulus) * (-2.0 * np.pi / modulus)
        z = torch.complex(torch.cos(ang), torch.sin(ang))
        if amplitude != 1.0:
            z = z * float(amplitude)
        if n_in < self.slot_count:
            z = torch.nn.functional.pad(z, (0, self.slot_count - n_in))            # zero-padded, like encrypt()
        return self._encrypt_device(z, public_key, self.params.max_level if level is None else int(level))

    def decrypt_zeta(self, ct: Ciphertext, sk: SecretKey, modulus: int = 16) -> np.ndarray:
        """decrypt + ZetaEncoder.from_zeta on the device: uint8 [slot_count] or [B, slot_count]."""
        import torch
        if not hasattr(self.backend, "decode_device"):
            z = np.atleast_2d(self.decrypt(ct, sk))
            out = (np.rint(-np.angle(z) * modulus / (2 * np.pi)).astype(np.int64) % modulus).astype(np.uint8)
        else:
            z = self.decrypt_device(ct, sk)
            k = torch.round(torch.angle(z) * (-modulus / (2.0 * np.pi))).to(torch.int64) % modulus
            out = k.to(torch.uint8).cpu().numpy()
        return out[0] if out.shape[0] == 1 else out

    def decrypt_to_plaintext_coeffs(self, ct: Ciphertext, sk: SecretKey) -> np.ndarray:
        """Centred coefficient vectors (float64 [batch, N]) of c0 + c1 s (+ c2 s^2)."""
        be = self.backend
        use = min(2, ct.level + 1)
        c = be.take_limbs(ct.polys, use, False)
        s = be.take_limbs(sk.ntt, use, False)
        acc = be.add(be.select_poly(c, 0), be.mul(be.select_poly(c, 1), s, use, 0), use, 0)
        if ct.npoly == 3:
            s2 = be.mul(s, s, use, 0)
            acc = be.add(acc, be.mul(be.select_poly(c, 2), s2, use, 0), use, 0)
        acc = be.intt(acc, use, 0)
        return be.crt_centered(acc, use)

    def decrypt(self, ct: Ciphertext, sk: SecretKey) -> np.ndarray:
        if self.device_codec and hasattr(self.backend, "decode_device"):
            out = self.decrypt_device(ct, sk).cpu().numpy()
            return out[0] if out.shape[0] == 1 else out
        m = self.decrypt_to_plaintext_coeffs(ct, sk)
        self._count('decrypt', m.shape[0])
        scale = self.params.scale(ct.level)
        out = np.stack([encoding.coeffs_to_slots(r, self.params.log_n) for r in m]) / scale
        return out[0] if out.shape[0] == 1 else out

    # ------------------------------------------------------------------ level management
    def _rescale(self, ct: Ciphertext) -> Ciphertext:
        if ct.level == 0:
            raise RuntimeError("cannot rescale: ciphertext is at level 0 (no multiplicative depth left)")
        self._count('rescale')
        return Ciphertext(self, self.backend.rescale(ct.polys, ct.level + 1), ct.level - 1)

    def _mul_int_const(self, ct: Ciphertext, re: int, im: int) -> Ciphertext:
        nq = ct.level + 1
        cp, cm = self._const_residues(re, im, nq)
        return Ciphertext(self, self.backend.mul_const(ct.polys, cp, cm, nq), ct.level)

    def level_down(self, ct: Ciphertext, target: int, factor=None) -> Ciphertext:
        """Bring ``ct`` to ``target`` (< ct.level) *and* onto that level's scale.  ``factor`` (a real number)
        multiplies the message on the way: the constant multiply that changes the scale carries it, so
        ``factor * ct`` at a lower level costs one constant multiply + one rescale on target + 2 limbs."""
        if target == ct.level and factor is None:
            return ct
        if target >= ct.level:
            raise RuntimeError("cannot raise the level of a ciphertext without bootstrapping" if factor is None
                               else "level_down with a factor needs a lower target level")
        P = self.params
        if ct.zero:
            return self._zero(target, ct.npoly, ct.batch)
        polys = ct.polys
        if ct.level > target + 1:
            polys = self.backend.take_limbs(polys, target + 2, False)
        c = P.delta[target] * P.moduli[target + 1] / P.delta[ct.level]
        if factor is not None:
            c = c * Fraction(float(factor))
        tmp = self._mul_int_const(Ciphertext(self, polys, target + 1), int(round(c)), 0)
        self._count('level_adjust')
        return self._rescale(tmp)

    def _align(self, a: Ciphertext, b: Ciphertext):
        if a.level > b.level:
            a = self.level_down(a, b.level)
        elif b.level > a.level:
            b = self.level_down(b, a.level)
        return a, b

    def _zero(self, level: int, npoly: int = 2, batch: int = 1) -> Ciphertext:
        return Ciphertext(self, self.backend.zeros(npoly, batch, level + 1, False), level, zero=True)

    # ------------------------------------------------------------------ arithmetic
    def add(self, a, b):
        if isinstance(a, Ciphertext) and isinstance(b, Ciphertext):
            a, b = self._align(a, b)
            if a.zero and a.npoly <= b.npoly:
                return b
            if b.zero and b.npoly <= a.npoly:
                return a
            if a.npoly != b.npoly:
                a, b = self._pad_poly(a, b)
            nq = a.level + 1
            self._count('add_ct')
            return Ciphertext(self, self.backend.add(a.polys, b.polys, nq, 0), a.level)
        if isinstance(b, Ciphertext):
            a, b = b, a
        if not isinstance(a, Ciphertext):
            raise TypeError("add: at least one operand must be a Ciphertext")
        if isinstance(b, Plaintext):
            if b.const is not None:
                return self.add_plain(a, b.const)
            self._count('add_pt')
            return Ciphertext(self, self.backend.add_poly0(a.polys, b.at_level(a.level), a.level + 1), a.level)
        return self.add_plain(a, b)

    def _pad_poly(self, a: Ciphertext, b: Ciphertext):
        be = self.backend
        if a.npoly < b.npoly:
            a = Ciphertext(self, be.concat([a.polys, be.zeros(b.npoly - a.npoly, a.batch, a.level + 1, False)]), a.level)
        else:
            b = Ciphertext(self, be.concat([b.polys, be.zeros(a.npoly - b.npoly, b.batch, b.level + 1, False)]), b.level)
        return a, b

    def add_plain(self, ct: Ciphertext, value, inplace: bool = False) -> Ciphertext:
        """ct + constant.  inplace: ``ct`` is a temporary of the caller (a product it has just formed): the constant
        is added into its storage instead of a copy of it."""
        re, im = encoding.const_i64(complex(value), self.params.delta[ct.level])
        nq = ct.level + 1
        cp, cm = self._const_residues(re, im, nq)
        self._count('add_const')
        if inplace:
            return Ciphertext(self, self.backend.add_const(ct.polys, cp, cm, nq, True), ct.level)
        return Ciphertext(self, self.backend.add_const(ct.polys, cp, cm, nq), ct.level)

    def subtract(self, a: Ciphertext, b: Ciphertext) -> Ciphertext:
        a, b = self._align(a, b)
        if a.npoly != b.npoly:
            a, b = self._pad_poly(a, b)
        return Ciphertext(self, self.backend.sub(a.polys, b.polys, a.level + 1, 0), a.level)

    def negate(self, a: Ciphertext) -> Ciphertext:
        return Ciphertext(self, self.backend.neg(a.polys, a.level + 1, 0), a.level, a.zero)

    def multiply(self, a, b, relin_key: Optional[RelinearizationKey] = None):
        if isinstance(a, Ciphertext) and isinstance(b, Ciphertext):
            return self._mul_ct(a, b, relin_key)
        if isinstance(b, Ciphertext):
            a, b = b, a
        if not isinstance(a, Ciphertext):
            raise TypeError("multiply: at least one operand must be a Ciphertext")
        if isinstance(b, Plaintext):
            if b.const is not None:
                return self._mul_scalar(a, b.const)
            if a.level == 0:
                raise RuntimeError("multiply: no multiplicative depth left")
            if a.zero:
                return self._zero(a.level - 1, a.npoly, a.batch)
            self._count('mul_pt')
            prod = self.backend.mul(a.polys, b.at_level(a.level), a.level + 1, 0)
            return self._rescale(Ciphertext(self, prod, a.level))
        return self._mul_scalar(a, b)

    def multiply_plain_sum(self, cts: Sequence[Ciphertext], pts: Sequence[Plaintext], rescale: bool = True) -> Ciphertext:
        """sum_i ct_i (.) pt_i with ONE rescale (the rotate-mask-add pattern of
        /root/reference/shiftrows_service.py:41-50 without a rescale per mask).
        ``rescale=False`` returns the sum before that rescale (same level, scale delta * q_level): the
        caller rotates / adds such sums and rescales once (baby-step/giant-step linear transforms)."""
        lvl = min(c.level for c in cts)
        if lvl == 0:
            raise RuntimeError("multiply: no multiplicative depth left")
        be = self.backend
        self._count('mul_pt', len(cts))
        if hasattr(be, "mul_plain_sum") and not _NO_FUSE and all(c.batch == cts[0].batch and c.npoly == 2 for c in cts):
            # one fused pass (fhe_mul_plain_sum): every operand read once, the sum written once
            acc = be.mul_plain_sum([self.level_down(c, lvl).polys for c in cts], [p.at_level(lvl) for p in pts], lvl + 1)
            return self._rescale(Ciphertext(self, acc, lvl)) if rescale else Ciphertext(self, acc, lvl)
        acc = None
        for ct, pt in zip(cts, pts):
            ct = self.level_down(ct, lvl)
            prod = be.mul(ct.polys, pt.at_level(lvl), lvl + 1, 0)
            acc = prod if acc is None else be.add(acc, prod, lvl + 1, 0)
        return self._rescale(Ciphertext(self, acc, lvl)) if rescale else Ciphertext(self, acc, lvl)

    def multiply_plain_sums(self, cts: Sequence[Ciphertext], pt_rows: Sequence[Sequence[Optional[Plaintext]]]) -> List[Ciphertext]:
        """[sum_t ct_t (.) pt_rows[g][t] for g] BEFORE the rescale (same level, scale delta * q_level), every
        ciphertext read once for all rows (fhe_mul_plain_multi): the diagonal sums of all giant steps of a
        baby-step/giant-step linear transform.  ``None`` entries are absent terms."""
        lvl = min(c.level for c in cts)
        if lvl == 0:
            raise RuntimeError("multiply: no multiplicative depth left")
        be = self.backend
        G, T = len(pt_rows), len(cts)
        if (hasattr(be, "mul_plain_multi") and not _NO_FUSE and T <= 16 and G <= 8
                and all(c.batch == cts[0].batch and c.npoly == 2 and not c.zero for c in cts)):
            self._count('mul_pt', sum(1 for row in pt_rows for p in row if p is not None))
            polys = [self.level_down(c, lvl).polys for c in cts]
            out = be.mul_plain_multi(polys, [[None if p is None else p.at_level(lvl) for p in row] for row in pt_rows], lvl + 1)
            return [Ciphertext(self, out[g], lvl) for g in range(G)]
        outs = []
        for row in pt_rows:
            sel = [(c, p) for c, p in zip(cts, row) if p is not None]
            outs.append(self.multiply_plain_sum([c for c, _ in sel], [p for _, p in sel], rescale=False))
        return outs

    def rotate_hoisted(self, ct: Ciphertext, keys: Sequence[FixedRotationKey]) -> List[Ciphertext]:
        """Several rotations of ONE ciphertext sharing a single ModUp (the base extension commutes
        with the Galois automorphism bit for bit: centred digits are odd functions and the
        automorphism only permutes and negates coefficients), so every rotation after the first
        costs an automorphism, the key inner product and the ModDown only."""
        be = self.backend
        if ct.zero or not keys:
            return [ct for _ in keys]
        if not hasattr(be, "modup_raw") or not hasattr(be, "automorphism_rows") or len(keys) < 2 or _NO_FUSE:
            return [self._apply_galois(ct, k) for k in keys]
        nq = ct.level + 1
        c1 = be.select_poly(ct.polys, 1)
        ext = be.modup_raw(c1, nq)
        outs = []
        two_n = 2 * self.params.n
        for key in keys:
            # sigma(<e, sigma^-1(key)>) = <sigma(e), key>: with the key permuted once (cached on the key object) the
            # shared ModUp output is used as it is, and the automorphism runs once on the 2 n rows of the result
            # instead of on the beta (n + K) extended rows of every rotation -- bit-identical either way
            pre = getattr(key, "_pre_permuted", None)
            if pre is None:
                pre = key._pre_permuted = be.automorphism_rows(key.data, pow(int(key.galois), -1, two_n))
            acc = be.ks_inner(ext, c1, pre, nq)
            ks = be.moddown_inplace(acc, nq) if hasattr(be, "moddown_inplace") else be.moddown(acc, nq)
            self._count('keyswitch_galois')
            add0 = be.add_poly0_inplace if hasattr(be, "add_poly0_inplace") else be.add_poly0
            outs.append(Ciphertext(self, be.automorphism(add0(ks, be.select_poly(ct.polys, 0), nq), key.galois, nq, 0),
                                   ct.level))
        return outs

    def _mul_scalar(self, ct: Ciphertext, value) -> Ciphertext:
        """ct x complex constant, rescaled (the reference uses ``multiply(ct, 0.0)`` as
        its "zero ciphertext": /root/reference/xor_service.py:282)."""
        if ct.level == 0:
            raise RuntimeError("multiply: no multiplicative depth left")
        v = complex(value)
        if v == 0 or ct.zero:
            return self._zero(ct.level - 1, ct.npoly, ct.batch)
        re, im = encoding.const_i64(v, self.params.delta[ct.level])
        self._count('mul_const')
        return self._rescale(self._mul_int_const(ct, re, im))

    def _mul_ct(self, a: Ciphertext, b: Ciphertext, rlk: Optional[RelinearizationKey]) -> Ciphertext:
        a, b = self._align(a, b)
        if a.npoly != 2 or b.npoly != 2:
            raise RuntimeError("multiply: operands must have 2 polynomials (relinearize first)")
        if a.level == 0:
            raise RuntimeError("multiply: no multiplicative depth left")
        if a.zero or b.zero:
            return self._zero(a.level - 1, 2 if rlk is not None else 3, max(a.batch, b.batch))
        nq = a.level + 1
        self._count('mul_ct')
        if rlk is not None:
            # relinearise and rescale together: one ModDown by P * q_l instead of ModDown + rescale
            self._count('keyswitch_relin')
            self._count('rescale')
            if hasattr(self.backend, "mul_relin_rescale"):
                # ... and the tensor product inside the key switch (never stored)
                return Ciphertext(self, self.backend.mul_relin_rescale(a.polys, b.polys, rlk.data, nq), a.level - 1)
            d = self.backend.tensor(a.polys, b.polys, nq)
            return Ciphertext(self, self.backend.relin_rescale(d, rlk.data, nq), a.level - 1)
        d = self.backend.tensor(a.polys, b.polys, nq)
        return self._rescale(Ciphertext(self, d, a.level))

    def _mul_ct_dropped(self, a: Ciphertext, b: Ciphertext, rlk: RelinearizationKey):
        """a x b + relinearise + rescale where an operand of a higher level is used at the lower level by
        ignoring its upper limbs -- an exact modulus switch (same message, same scale, no noise, no key
        switch, no transform) -- instead of ``level_down`` (constant multiply + rescale: 2 n limb transforms).
        The product then is off the scale table: returns ``(ct, dev)`` with the true scale of ``ct`` equal
        to ``delta[ct.level] * dev``.  Only the LUT schedules (aes_fhe_b200/fused.py) use this: they fold
        ``dev`` into their constants, so no ciphertext with an off-table scale leaves them."""
        if a.npoly != 2 or b.npoly != 2:
            raise RuntimeError("multiply: operands must have 2 polynomials (relinearize first)")
        P, be = self.params, self.backend
        lvl = min(a.level, b.level)
        if lvl == 0:
            raise RuntimeError("multiply: no multiplicative depth left")
        if a.zero or b.zero:
            return self._zero(lvl - 1, 2, max(a.batch, b.batch)), Fraction(1)
        dev = (P.delta[a.level] / P.delta[lvl]) * (P.delta[b.level] / P.delta[lvl])
        nq = lvl + 1
        self._count('mul_ct')
        self._count('keyswitch_relin')
        self._count('rescale')
        if hasattr(be, "mul_relin_rescale"):
            out = be.mul_relin_rescale(a.polys, b.polys, rlk.data, nq)
        else:
            ap = a.polys if a.level == lvl else be.take_limbs(a.polys, nq, False)
            bp = b.polys if b.level == lvl else be.take_limbs(b.polys, nq, False)
            out = be.relin_rescale(be.tensor(ap, bp, nq), rlk.data, nq)
        return Ciphertext(self, out, lvl - 1), dev

    def multiply_gather(self, parts_a, parts_b, rlk: RelinearizationKey):
        """Batched ct x ct + relinearise + rescale whose operands are GATHERED: parts_x = [(Ciphertext, batch indices),
        ...] (all ciphertexts of one side at one level); element i of the result is the product of the i-th listed
        element of side a and the i-th of side b.  On the B200 nothing is copied (fhe_mul_relin_rescale_ptrs); other
        backends materialise the two sides.  An operand above the product's level is used in place, as in
        ``_mul_ct_dropped``: returns ``(ct, dev)`` with the true scale ``delta[ct.level] * dev``."""
        P, be = self.params, self.backend
        la, lb = {c.level for c, _ in parts_a}, {c.level for c, _ in parts_b}
        if len(la) != 1 or len(lb) != 1:
            raise RuntimeError("multiply_gather: the ciphertexts of one side must share a level")
        la, lb = la.pop(), lb.pop()
        if any(c.npoly != 2 for c, _ in list(parts_a) + list(parts_b)):
            raise RuntimeError("multiply: operands must have 2 polynomials (relinearize first)")
        lvl = min(la, lb)
        if lvl == 0:
            raise RuntimeError("multiply: no multiplicative depth left")
        nq = lvl + 1
        if hasattr(be, "mul_relin_rescale_gather") and P.digits_at(nq) <= 4:
            dev = (P.delta[la] / P.delta[lvl]) * (P.delta[lb] / P.delta[lvl])
            self._count('mul_ct')
            self._count('keyswitch_relin')
            self._count('rescale')
            out = be.mul_relin_rescale_gather([(c.polys, list(i)) for c, i in parts_a], [(c.polys, list(i)) for c, i in parts_b],
                                              rlk.data, nq)
            return Ciphertext(self, out, lvl - 1), dev

        def side(parts, level):
            hs = [be.permute_batch(c.polys, list(i)) for c, i in parts]
            return Ciphertext(self, hs[0] if len(hs) == 1 else be.concat_batch(hs), level)
        return self._mul_ct_dropped(side(parts_a, la), side(parts_b, lb), rlk)

    def square(self, a: Ciphertext, relin_key=None) -> Ciphertext:
        return self._mul_ct(a, a, relin_key)

    def _relin(self, ct3: Ciphertext, rlk: SwitchKey) -> Ciphertext:
        be = self.backend
        nq = ct3.level + 1
        self._count('keyswitch_relin')
        ks = be.keyswitch(be.select_poly(ct3.polys, 2), rlk.data, nq)
        return Ciphertext(self, be.add(be.take_polys(ct3.polys, 2), ks, nq, 0), ct3.level)

    def relinearize(self, ct: Ciphertext, relin_key: RelinearizationKey) -> Ciphertext:
        if ct.npoly != 3:
            # message pinned by /root/reference/xor_service.py:112-118
            raise RuntimeError("relinearize: ciphertext should have 3 polynomials, "
                               f"but it has {ct.npoly}")
        if ct.zero:
            return self._zero(ct.level, 2, ct.batch)
        return self._relin(ct, relin_key)

    def rescale(self, ct: Ciphertext) -> Ciphertext:
        """Public alias; every multiply in this engine already rescales."""
        return self._rescale(ct)

    # ------------------------------------------------------------------ powers
    def make_power_basis(self, ct: Ciphertext, degree: int, relin_key) -> List[Ciphertext]:
        """[ct^1, ..., ct^degree]; ct^k sits ceil(log2 k) levels below ct
        (/root/reference/xor_service.py:86, sbox/sbox_service.py:93)."""
        degree = int(degree)
        if degree < 1:
            return []
        pw: List[Optional[Ciphertext]] = [None] * (degree + 1)
        pw[1] = ct
        for k in range(2, degree + 1):
            hi = (k + 1) // 2
            lo = k // 2
            pw[k] = self._mul_ct(pw[hi], pw[lo], relin_key)
        return pw[1:]

    # ------------------------------------------------------------------ automorphisms
    def _apply_galois(self, ct: Ciphertext, key: SwitchKey) -> Ciphertext:
        be = self.backend
        if ct.npoly != 2:
            raise RuntimeError("rotate/conjugate: ciphertext must have 2 polynomials")
        nq = ct.level + 1
        rot = be.automorphism(ct.polys, key.galois, nq, 0)
        ks = be.keyswitch(be.select_poly(rot, 1), key.data, nq)
        self._count('keyswitch_galois')
        add0 = be.add_poly0_inplace if hasattr(be, "add_poly0_inplace") else be.add_poly0      # ks is a temporary
        return Ciphertext(self, add0(ks, be.select_poly(rot, 0), nq), ct.level)

    def conjugate(self, ct: Ciphertext, conj_key: ConjugationKey) -> Ciphertext:
        if ct.zero:
            return ct
        return self._apply_galois(ct, conj_key)

    def rotate(self, ct: Ciphertext, rot_key, delta: Optional[int] = None) -> Ciphertext:
        """``out = np.roll(in, delta)`` (/root/reference/test/test_engine_rot.py:32-40)."""
        if isinstance(rot_key, FixedRotationKey):
            if delta is not None and int(delta) % self.slot_count != rot_key.delta % self.slot_count:
                raise ValueError("fixed rotation key was generated for a different delta")
            return ct if ct.zero else self._apply_galois(ct, rot_key)
        if delta is None:
            raise TypeError("rotate(ct, rotation_key, delta): delta is required")
        d = int(delta) % self.slot_count
        if d == 0 or ct.zero:
            return ct
        if d in rot_key.keys:
            return self._apply_galois(ct, rot_key.keys[d])
        if d - self.slot_count in rot_key.keys:
            return self._apply_galois(ct, rot_key.keys[d - self.slot_count])
        out = ct
        for step in _naf_steps(d, self.slot_count):
            key = rot_key.keys.get(step) or rot_key.keys.get(step % self.slot_count)
            if key is None:
                raise RuntimeError(f"rotation key lacks step {step} needed for delta {delta}")
            out = self._apply_galois(out, key)
        return out

    def multiply_by_i(self, ct: Ciphertext, sign: int = 1) -> Ciphertext:
        """ct * (+-i): multiplication by the monomial +-X^(N/2) -- exact, no level, no rescale."""
        nq = ct.level + 1
        cp, cm = self._const_residues(0, 1 if sign >= 0 else -1, nq)
        return Ciphertext(self, self.backend.mul_const(ct.polys, cp, cm, nq), ct.level, ct.zero)

    # ------------------------------------------------------------------ bootstrap
    def create_bootstrap_key(self, sk: SecretKey, groups: int = 3, groups_stc: Optional[int] = None):
        """engine_context.py:73.  `groups` / `groups_stc`: number of BSGS linear transforms of
        CoeffToSlot / SlotToCoeff (depth = 13 + groups + groups_stc; fewer = more levels left, more
        rotations)."""
        from .bootstrap import make_bootstrap_key
        return make_bootstrap_key(self, sk, groups, groups_stc)

    def create_small_bootstrap_key(self, sk: SecretKey):
        return BootstrapKey(small=True)

    def bootstrap(self, ct: Ciphertext, relin_key, conj_key, boot_key) -> Ciphertext:
        """Refresh the level of a ciphertext (xor_service.py:120-129, :274-277)."""
        from .bootstrap import bootstrap
        return bootstrap(self, ct, relin_key, conj_key, boot_key)

    def bootstrap_bits(self, ct: Ciphertext, relin_key, conj_key, boot_key, top_level: Optional[int] = None) -> Ciphertext:
        """Refresh of a ciphertext whose slots are u + i v with u, v = +-1 (two bit planes): returns the batch
        [all u, all v] of real ciphertexts, cleaned (aes_fhe_b200/bootstrap.py::bootstrap_bits).  top_level: raise to
        this level instead of max_level (the result sits at top_level - depth; every step runs on fewer limbs)."""
        from .bootstrap import bootstrap_bits
        return bootstrap_bits(self, ct, relin_key, conj_key, boot_key, top_level)
