"""PlainEngine: the `desilofhe.Engine` call surface on bare complex128 slot vectors (no
encryption, no noise).  TEST INFRASTRUCTURE ONLY.  Running a reference function on it gives the
slot values the same function must reproduce on the real engine within the CKKS error bound
("op-sequence parity"; precedent: the DummyEngine of
/root/reference/test/test_recombine_mixcol.py:9-13).  Semantics follow SURVEY Appendix A.3."""
from __future__ import annotations

import numpy as np


class Plaintext:
    def __init__(self, v):
        self.values = np.asarray(v, dtype=np.complex128)


class Ciphertext:
    def __init__(self, v, level, npoly=2):
        self.v = np.asarray(v, dtype=np.complex128)
        self.level = level
        self.npoly = npoly


class Engine:
    def __init__(self, *args, max_level=None, log_coeff_count=None, slot_count=None, bootstrap_level=10, **kw):
        self.max_level = 30 if max_level is None else int(max_level)
        if slot_count is None:
            slot_count = 1 << ((log_coeff_count or 16) - 1)
        self.slot_count = int(slot_count)
        self.bootstrap_level = bootstrap_level
        self.op_counts = {}

    def _c(self, k):
        self.op_counts[k] = self.op_counts.get(k, 0) + 1

    # keys are opaque tokens
    def create_secret_key(self): return "sk"
    def create_public_key(self, sk): return "pk"
    def create_relinearization_key(self, sk): return "rlk"
    def create_conjugation_key(self, sk): return "cjk"
    def create_rotation_key(self, sk, steps=None): return "rot"
    def create_fixed_rotation_key(self, sk, delta): return ("fixed", int(delta))
    def create_small_bootstrap_key(self, sk): return "sbk"
    def create_bootstrap_key(self, sk): return "bk"

    def _pad(self, data):
        v = np.zeros(self.slot_count, dtype=np.complex128)
        d = np.asarray(data).ravel()
        v[:d.size] = d
        return v

    def encode(self, v): return Plaintext(self._pad(v))
    def encrypt(self, data, pk, level=None):
        if isinstance(data, Plaintext):
            data = data.values
        return Ciphertext(self._pad(data), self.max_level if level is None else level)
    def decrypt(self, ct, sk): return ct.v.copy()

    def add(self, a, b):
        if isinstance(b, Ciphertext) and not isinstance(a, Ciphertext):
            a, b = b, a
        if isinstance(b, Ciphertext):
            return Ciphertext(a.v + b.v, min(a.level, b.level), max(a.npoly, b.npoly))
        if isinstance(b, Plaintext):
            return Ciphertext(a.v + b.values, a.level, a.npoly)
        return Ciphertext(a.v + b, a.level, a.npoly)

    def add_plain(self, ct, val): return Ciphertext(ct.v + val, ct.level, ct.npoly)

    def multiply(self, a, b, relin_key=None):
        if isinstance(b, Ciphertext) and not isinstance(a, Ciphertext):
            a, b = b, a
        if isinstance(b, Ciphertext):
            self._c("mul_ct")
            if relin_key is not None:
                self._c("keyswitch_relin")
            lvl = min(a.level, b.level) - 1
            if lvl < 0:
                raise RuntimeError("multiply: no multiplicative depth left")
            return Ciphertext(a.v * b.v, lvl, 2 if relin_key is not None else 3)
        if a.level < 1:
            raise RuntimeError("multiply: no multiplicative depth left")
        self._c("mul_pt")
        w = b.values if isinstance(b, Plaintext) else b
        return Ciphertext(a.v * w, a.level - 1, a.npoly)

    def relinearize(self, ct, rlk):
        if ct.npoly != 3:
            raise RuntimeError(f"relinearize: ciphertext should have 3 polynomials, but it has {ct.npoly}")
        self._c("keyswitch_relin")
        return Ciphertext(ct.v, ct.level, 2)

    def make_power_basis(self, ct, degree, rlk):
        out = []
        for k in range(1, int(degree) + 1):
            depth = int(np.ceil(np.log2(k))) if k > 1 else 0
            if ct.level - depth < 0:
                raise RuntimeError("multiply: no multiplicative depth left")
            out.append(Ciphertext(ct.v ** k, ct.level - depth))
            if k > 1:
                self._c("keyswitch_relin")
        return out

    def conjugate(self, ct, key):
        self._c("keyswitch_galois")
        return Ciphertext(np.conj(ct.v), ct.level)

    def rotate(self, ct, key, delta=None):
        if isinstance(key, tuple):
            delta = key[1]
        if int(delta) % self.slot_count:
            self._c("keyswitch_galois")
        return Ciphertext(np.roll(ct.v, int(delta)), ct.level)

    def bootstrap(self, ct, rlk, cjk, bk):
        self._c("bootstrap")
        return Ciphertext(ct.v, self.bootstrap_level)
