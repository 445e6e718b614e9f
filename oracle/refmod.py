"""ctypes binding of oracle/refmod.cpp + the `RefBackend` adapter.

TEST INFRASTRUCTURE ONLY (see the header of refmod.cpp).  `RefBackend` implements the
backend interface of ``aes_fhe_b200.engine.Engine`` on NumPy ``uint64`` arrays so the same
facade (level/scale bookkeeping) can be driven by the CPU restatement; tests compare its
residues bit-for-bit with the CUDA backend's, and bench.py times it as the CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import List, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "librefmod.so"
    src = _HERE / "refmod.cpp"
    if force or not so.exists() or (src.exists() and so.stat().st_mtime < src.stat().st_mtime):
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.ref_ctx_create.restype = C.c_void_p
        _LIB.ref_ctx_threads.restype = C.c_int
        _LIB.ref_ctx_threads.argtypes = [C.c_void_p]
        _LIB.ref_ntt_rows.restype = C.c_ulonglong
        _LIB.ref_ntt_rows.argtypes = []
    return _LIB


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class RefBackend:
    name = "refmod-cpu"

    def __init__(self, params, threads: int = 0):
        self.params = params
        self.n = params.n
        L = lib()
        mod = np.array(params.moduli, dtype=np.uint64)
        psi = np.array(params.psi, dtype=np.uint64)
        self._ctx = C.c_void_p(L.ref_ctx_create(params.log_n, params.n_q, params.n_p, params.alpha,
                                                _p(mod), _p(psi), int(threads)))
        self.threads = L.ref_ctx_threads(self._ctx)
        self._L = L

    def ntt_row_count(self) -> int:
        """length-N transforms done by the library since load"""
        return int(self._L.ref_ntt_rows())

    def __del__(self):
        try:
            self._L.ref_ctx_destroy(self._ctx)
        except Exception:
            pass

    # ---- layout helpers.  Handles are uint64 arrays [npoly, batch, limbs, N].
    def _ids(self, nq: int, np_: int, reps: int = 1) -> np.ndarray:
        ids = list(range(nq)) + [self.params.n_q + k for k in range(np_)]
        return np.array(ids * reps, dtype=np.int32)

    def npoly(self, h) -> int:
        return h.shape[0]

    def batch(self, h) -> int:
        return h.shape[1]

    def from_numpy(self, a: np.ndarray):
        return np.ascontiguousarray(a, dtype=np.uint64)

    def to_numpy(self, h) -> np.ndarray:
        return np.array(h, dtype=np.uint64, copy=True)

    def zeros(self, npoly: int, batch: int, nq: int, with_p: bool):
        return np.zeros((npoly, batch, nq + (self.params.n_p if with_p else 0), self.n), dtype=np.uint64)

    def take_limbs(self, h, nq: int, with_p: bool):
        if with_p:
            K = self.params.n_p
            return np.ascontiguousarray(np.concatenate([h[:, :, :nq], h[:, :, h.shape[2] - K:]], axis=2))
        return np.ascontiguousarray(h[:, :, :nq])

    def select_poly(self, h, i: int):
        return h[i:i + 1]

    def take_polys(self, h, k: int):
        return h[:k]

    def concat(self, hs: List):
        return np.ascontiguousarray(np.concatenate(hs, axis=0))

    def stack(self, hs: List):
        return np.ascontiguousarray(np.stack(hs, axis=0))

    def concat_batch(self, hs: List):
        return np.ascontiguousarray(np.concatenate(hs, axis=1))

    def split_batch(self, h, sizes: Sequence[int]):
        out, o = [], 0
        for s in sizes:
            out.append(np.ascontiguousarray(h[:, o:o + s]))
            o += s
        return out

    def alloc(self, shape):
        return np.empty(tuple(shape), dtype=np.uint64)

    def slice_batch(self, h, lo: int, hi: int):
        return np.ascontiguousarray(h[:, lo:hi])

    def permute_batch(self, h, idx: Sequence[int]):
        return np.ascontiguousarray(h[:, np.asarray(idx, dtype=np.int64)])

    def mod_raise(self, h, nq_out: int):
        coef = self.intt(h, 1, 0)                                   # [p, B, 1, N] coefficients mod q0
        q0 = self.params.moduli[0]
        c = coef[:, :, 0, :].astype(np.int64)
        c = np.where(c > (q0 >> 1), c - q0, c)                      # centred lift
        out = np.empty(h.shape[:2] + (nq_out, self.n), dtype=np.uint64)
        for p in range(h.shape[0]):
            out[p:p + 1] = self.from_i64(c[p], nq_out, False)
        return out

    def expand_batch(self, h, batch: int):
        return h if h.shape[1] == batch else np.ascontiguousarray(np.broadcast_to(h, (h.shape[0], batch) + h.shape[2:]))

    # ---- transforms
    def ntt(self, h, nq: int, np_: int):
        out = np.array(h, copy=True)
        ids = self._ids(nq, np_, out.shape[0] * out.shape[1])
        self._L.ref_ntt(self._ctx, _p(out), _p(ids), C.c_int(len(ids)))
        return out

    def intt(self, h, nq: int, np_: int):
        out = np.array(h, copy=True)
        ids = self._ids(nq, np_, out.shape[0] * out.shape[1])
        self._L.ref_intt(self._ctx, _p(out), _p(ids), C.c_int(len(ids)))
        return out

    def from_i64(self, coeffs: np.ndarray, nq: int, with_p: bool):
        np_ = self.params.n_p if with_p else 0
        ids = self._ids(nq, np_)
        c = np.ascontiguousarray(coeffs, dtype=np.int64).reshape(-1, self.n)
        out = np.empty((1, c.shape[0], len(ids), self.n), dtype=np.uint64)
        for b in range(c.shape[0]):
            o = np.empty((len(ids), self.n), dtype=np.uint64)
            cb = np.ascontiguousarray(c[b])
            self._L.ref_from_i64(self._ctx, _p(o), _p(cb), _p(ids), C.c_int(len(ids)))
            self._L.ref_ntt(self._ctx, _p(o), _p(ids), C.c_int(len(ids)))
            out[0, b] = o
        return out

    # ---- elementwise
    def _bin(self, fn, a, b, nq, np_):
        shape = (max(a.shape[0], b.shape[0]), max(a.shape[1], b.shape[1])) + a.shape[2:]
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
        b = np.ascontiguousarray(np.broadcast_to(b, shape))
        out = np.empty_like(a)
        ids = self._ids(nq, np_, shape[0] * shape[1])
        fn(self._ctx, _p(out), _p(a), _p(b), _p(ids), C.c_int(len(ids)))
        return out

    def add(self, a, b, nq, np_):
        return self._bin(self._L.ref_add, a, b, nq, np_)

    def sub(self, a, b, nq, np_):
        return self._bin(self._L.ref_sub, a, b, nq, np_)

    def mul(self, a, b, nq, np_):
        return self._bin(self._L.ref_mul, a, b, nq, np_)

    def neg(self, a, nq, np_):
        a = np.ascontiguousarray(a)
        out = np.empty_like(a)
        ids = self._ids(nq, np_, a.shape[0] * a.shape[1])
        self._L.ref_neg(self._ctx, _p(out), _p(a), _p(ids), C.c_int(len(ids)))
        return out

    def _constop(self, fn, a, cp, cm, nq, np_):
        a = np.ascontiguousarray(a)
        reps = a.shape[0] * a.shape[1]
        fp = np.array([int(v) for v in cp] * reps, dtype=np.uint64)
        fm = np.array([int(v) for v in cm] * reps, dtype=np.uint64)
        out = np.empty_like(a)
        ids = self._ids(nq, np_, reps)
        fn(self._ctx, _p(out), _p(a), _p(fp), _p(fm), _p(ids), C.c_int(len(ids)))
        return out

    def mul_scalar(self, a, fac: Sequence[int], nq, np_):
        return self._constop(self._L.ref_mul_const, a, fac, fac, nq, np_)

    def mul_const(self, a, cp, cm, nq):
        return self._constop(self._L.ref_mul_const, a, cp, cm, nq, 0)

    def add_const(self, a, cp, cm, nq, inplace: bool = False):
        out = np.array(a, copy=True)
        out[0:1] = self._constop(self._L.ref_add_const, out[0:1], cp, cm, nq, 0)
        return out

    def add_poly0(self, a, p, nq):
        out = np.array(a, copy=True)
        out[0:1] = self.add(out[0:1], p, nq, 0)
        return out

    def tensor(self, a, b, nq):
        a0, a1, b0, b1 = a[0:1], a[1:2], b[0:1], b[1:2]
        d0 = self.mul(a0, b0, nq, 0)
        d1 = self.add(self.mul(a0, b1, nq, 0), self.mul(a1, b0, nq, 0), nq, 0)
        d2 = self.mul(a1, b1, nq, 0)
        return self.concat([d0, d1, d2])

    # ---- structural
    def rescale(self, h, nq):
        h = np.ascontiguousarray(h)
        out = np.empty(h.shape[:2] + (nq - 1, self.n), dtype=np.uint64)
        for i in range(h.shape[0]):
            for b in range(h.shape[1]):
                o = np.empty((nq - 1, self.n), dtype=np.uint64)
                self._L.ref_rescale(self._ctx, _p(o), _p(np.ascontiguousarray(h[i, b])), C.c_int(nq))
                out[i, b] = o
        return out

    def automorphism(self, h, g: int, nq, np_):
        h = np.ascontiguousarray(h)
        out = np.empty_like(h)
        self._L.ref_automorphism(self._ctx, _p(out), _p(h), C.c_uint64(int(g)),
                                 C.c_int(h.shape[0] * h.shape[1] * h.shape[2]))
        return out

    def modup(self, d, nq):
        """d [1, B, nq, N] -> ext [B, beta, nq+K, N]"""
        P = self.params
        beta = P.digits_at(nq)
        ext = np.empty((d.shape[1], beta, nq + P.n_p, self.n), dtype=np.uint64)
        for b in range(d.shape[1]):
            e = np.empty((beta, nq + P.n_p, self.n), dtype=np.uint64)
            self._L.ref_modup(self._ctx, _p(e), _p(np.ascontiguousarray(d[0, b])), C.c_int(nq))
            ext[b] = e
        return ext

    def ks_inner(self, ext, d, ksk, nq):
        """-> acc [2, B, nq+K, N]"""
        ksk = np.ascontiguousarray(ksk)
        acc = np.empty((2, ext.shape[0], nq + self.params.n_p, self.n), dtype=np.uint64)
        for b in range(ext.shape[0]):
            a = np.empty((2, nq + self.params.n_p, self.n), dtype=np.uint64)
            self._L.ref_ks_inner(self._ctx, _p(a), _p(np.ascontiguousarray(ext[b])), _p(ksk), C.c_int(nq))
            acc[:, b] = a
        return acc

    def moddown(self, acc, nq):
        acc = np.ascontiguousarray(acc)
        flat = acc.reshape(-1, acc.shape[2], self.n)
        out = np.empty((flat.shape[0], nq, self.n), dtype=np.uint64)
        self._L.ref_moddown(self._ctx, _p(out), _p(flat), C.c_int(nq), C.c_int(flat.shape[0]))
        return out.reshape(acc.shape[0], acc.shape[1], nq, self.n)

    def relin_rescale(self, d3, ksk, nq):
        """(d0, d1) + KS(d2), divided by P * q_{nq-1} in one step: [3,B,nq,N] -> [2,B,nq-1,N]."""
        P = self.params
        acc = self.ks_inner(self.modup(d3[2:3], nq), d3[2:3], ksk, nq)           # [2, B, nq+K, N]
        pprod = 1
        for p in P.p:
            pprod *= p
        lift = self.mul_scalar(np.ascontiguousarray(d3[0:2]), [pprod % P.moduli[l] for l in range(nq)], nq, 0)
        acc = np.array(acc, copy=True)
        acc[:, :, :nq] = self.add(np.ascontiguousarray(acc[:, :, :nq]), lift, nq, 0)
        flat = np.ascontiguousarray(acc).reshape(-1, acc.shape[2], self.n)
        out = np.empty((flat.shape[0], nq - 1, self.n), dtype=np.uint64)
        self._L.ref_moddown_ex(self._ctx, _p(out), _p(flat), C.c_int(nq), C.c_int(flat.shape[0]), C.c_int(1))
        return out.reshape(2, acc.shape[1], nq - 1, self.n)

    def keyswitch(self, d, ksk, nq):
        return self.moddown(self.ks_inner(self.modup(d, nq), d, ksk, nq), nq)

    def ks_accum(self, acc, d, ksk, lift, nq):
        """acc (+)= <ModUp(d), ksk> + P * lift in the extended basis (restated with the plain primitives)."""
        P = self.params
        src = d if d is not None else lift
        bt = src.shape[1]
        if d is not None:
            part = np.array(self.ks_inner(self.modup(d, nq), d, ksk, nq), copy=True)
        else:
            part = np.zeros((2, bt, nq + P.n_p, self.n), dtype=np.uint64)
        if lift is not None:
            pprod = 1
            for p in P.p:
                pprod *= p
            l = self.mul_scalar(np.ascontiguousarray(lift), [pprod % P.moduli[j] for j in range(nq)], nq, 0)
            k = l.shape[0]
            part[:k, :, :nq] = self.add(np.ascontiguousarray(part[:k, :, :nq]), l, nq, 0)
        return part if acc is None else self.add(acc, part, nq, P.n_p)

    def moddown_rescale(self, acc, nq):
        flat = np.ascontiguousarray(acc).reshape(-1, acc.shape[2], self.n)
        out = np.empty((flat.shape[0], nq - 1, self.n), dtype=np.uint64)
        self._L.ref_moddown_ex(self._ctx, _p(out), _p(flat), C.c_int(nq), C.c_int(flat.shape[0]), C.c_int(1))
        return out.reshape(acc.shape[0], acc.shape[1], nq - 1, self.n)

    # ---- fused LUT evaluation pieces, restated with the plain primitives
    def prepare_lincomb(self, const_res, c0_res, nq: int):
        return dict(const_res=const_res, c0_res=c0_res, nq=nq, M=len(const_res), T=len(const_res[0]))

    def lincomb(self, inputs: List, prep) -> List:
        nq = prep["nq"]
        outs = []
        for m in range(prep["M"]):
            acc = None
            for t, x in enumerate(inputs):
                cp, cm = prep["const_res"][m][t]
                term = self.mul_const(np.ascontiguousarray(x[:, :, :nq]), cp, cm, nq)
                acc = term if acc is None else self.add(acc, term, nq, 0)
            if prep["c0_res"] is not None:
                cp, cm = prep["c0_res"][m]
                acc = self.add_const(acc, cp, cm, nq)
            outs.append(acc)
        return outs

    def tensor_acc(self, acc, a_list: List, b_list: List, nq: int):
        for a, b in zip(a_list, b_list):
            t = self.tensor(np.ascontiguousarray(a[:, :, :nq]), b, nq)
            acc = t if acc is None else self.add(acc, t, nq, 0)
        return acc

    def crt_centered(self, h, use: int) -> np.ndarray:
        """h [1, B, use, N] -> float64 [B, N]"""
        h = np.ascontiguousarray(h)
        out = np.empty((h.shape[1], self.n), dtype=np.float64)
        for b in range(h.shape[1]):
            o = np.empty(self.n, dtype=np.float64)
            if use >= 2:
                self._L.ref_crt2_centered(self._ctx, _p(o), _p(np.ascontiguousarray(h[0, b, 0])),
                                          _p(np.ascontiguousarray(h[0, b, 1])))
            else:
                self._L.ref_crt1_centered(self._ctx, _p(o), _p(np.ascontiguousarray(h[0, b, 0])))
            out[b] = o
        return out

    def synchronize(self):
        pass
