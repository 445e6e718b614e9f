"""CUDA backend of the engine facade: torch owns device memory and the stream, every
operation is one call through the C ABI (include/aesfhe_b200.h) into hand-written sm_100a
kernels.  No CPU path: without the built library or without a GPU the constructor raises.

Handles are ``torch.int64`` tensors ``[npoly, batch, limbs, N]`` (bit pattern = unsigned
residue): ``batch`` independent ciphertexts travel through every operation together, which
is what fills the 148 SMs and lets one pass over a key-switching key serve the whole batch.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _capi


class CudaBackend:
    name = "cuda-sm100a"

    def __init__(self, params, device_id: int = 0, _lib_path=None, _device: str | None = None):
        self.params = params
        self.n = params.n
        self.lib = _capi.load(_lib_path)
        if _device is None:
            if not torch.cuda.is_available():
                raise _capi.FheError("CUDA device required: this engine has no CPU fallback")
            self.device = torch.device("cuda", device_id)
        else:
            # tests/emu only: the g++ -DFHE_EMU build of the same kernels on host memory
            self.device = torch.device(_device)
        self.is_cuda = self.device.type == "cuda"
        mod = np.array(params.moduli, dtype=np.uint64)
        psi = np.array(params.psi, dtype=np.uint64)
        h = C.c_void_p()
        rc = self.lib.fhe_ctx_create(C.byref(h), params.log_n, params.n_q, params.n_p, params.alpha,
                                     mod.ctypes.data, psi.ctypes.data, device_id if self.is_cuda else 0)
        _capi.check(self.lib, rc, "fhe_ctx_create")
        self.ctx = h
        self._K = params.n_p

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.fhe_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- plumbing
    def _stream(self):
        """the caller's current torch stream; ONE stream per context (include/aesfhe_b200.h): the scratch arena and
        the NTT hand-over counters are per context, so a second stream on the same engine would race on them"""
        if not self.is_cuda:
            return C.c_void_p(0)
        s = torch.cuda.current_stream(self.device).cuda_stream
        bound = self.__dict__.get("_bound_stream")
        if bound is None:
            self._bound_stream = s
        elif s != bound:
            raise _capi.FheError("this engine's context is bound to another CUDA stream: use one Engine per stream "
                                 "(scratch memory and NTT hand-over counters are per context)")
        return C.c_void_p(s)

    def check_status(self):
        """raise if a single-launch NTT ever gave up waiting for its hand-over (its output would be wrong).
        Called wherever results leave the device: decrypt, download, synchronize."""
        if self.is_cuda:
            _capi.check(self.lib, self.lib.fhe_ntt_fused_status(self.ctx), "fhe_ntt_fused_status")

    def _empty(self, *shape):
        return torch.empty(shape, dtype=torch.int64, device=self.device)

    def _call(self, name, *args):
        _capi.check(self.lib, getattr(self.lib, name)(self.ctx, self._stream(), *args), name)

    @staticmethod
    def _ptr(t: torch.Tensor):
        return C.c_void_p(t.data_ptr())

    def synchronize(self):
        if self.is_cuda:
            torch.cuda.synchronize(self.device)
            self.check_status()

    def launch_count(self) -> int:
        return int(self.lib.fhe_launch_count())

    def ntt_row_count(self) -> int:
        return int(self.lib.fhe_ntt_row_count())

    # ---- layout helpers
    def npoly(self, h) -> int:
        return h.shape[0]

    def batch(self, h) -> int:
        return h.shape[1]

    def from_numpy(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(self.device)

    def to_numpy(self, h) -> np.ndarray:
        out = h.detach().cpu().contiguous().numpy().view(np.uint64).copy()
        self.check_status()
        return out

    def zeros(self, npoly: int, batch: int, nq: int, with_p: bool):
        return torch.zeros((npoly, batch, nq + (self._K if with_p else 0), self.n), dtype=torch.int64,
                           device=self.device)

    def take_limbs(self, h, nq: int, with_p: bool):
        if with_p:
            return torch.cat([h[:, :, :nq], h[:, :, h.shape[2] - self._K:]], dim=2).contiguous()
        return h[:, :, :nq].contiguous()

    def select_poly(self, h, i: int):
        return h[i:i + 1]

    def take_polys(self, h, k: int):
        return h[:k]

    def concat(self, hs: List):
        return torch.cat(hs, dim=0)

    def stack(self, hs: List):
        return torch.stack(hs, dim=0)

    def concat_batch(self, hs: List):
        return torch.cat(hs, dim=1)

    def split_batch(self, h, sizes: Sequence[int]):
        return [x.contiguous() for x in torch.split(h, list(sizes), dim=1)]

    def alloc(self, shape):
        """uninitialised device tensor of 64-bit residues (receive buffer for distributed keys)"""
        return self._empty(*shape)

    tensor_acc_into = True      # tensor_acc(out=..., init=...): sums written into a batch slice of a larger accumulator

    def view_batch(self, h, lo: int, hi: int):
        """batch elements lo..hi-1 as a VIEW: for the kernels that take a polynomial stride (fhe_lincomb, fhe_tensor_acc)"""
        return h[:, lo:hi]

    def slice_batch(self, h, lo: int, hi: int):
        """batch elements lo..hi-1 as a tensor of their own: ONE copy here instead of one inside every kernel call that
        would otherwise receive the strided view (a slice of the LUT bases is used by up to nine calls)"""
        return h[:, lo:hi].contiguous()

    def permute_batch(self, h, idx: Sequence[int]):
        """batch elements re-ordered / replicated: out[:, i] = h[:, idx[i]] (one device copy)"""
        key = tuple(idx)
        cache = self.__dict__.setdefault("_perm_cache", {})
        t = cache.get(key)
        if t is None:
            t = cache[key] = torch.tensor(list(idx), dtype=torch.int64, device=self.device)
        return h.index_select(1, t)

    def mod_raise(self, h, nq_out: int):
        """[p, B, 1, N] at level 0 (NTT) -> [p, B, nq_out, N]: centred lift of the coefficients
        from q_0 to the first nq_out moduli (the overflow polynomial q_0 * I comes with it)."""
        h = h.contiguous()
        out = self._empty(h.shape[0], h.shape[1], nq_out, self.n)
        self._call("fhe_mod_raise", self._ptr(out), self._ptr(h), h.shape[0] * h.shape[1], nq_out)
        return out

    def expand_batch(self, h, batch: int):
        return h if h.shape[1] == batch else h.expand(-1, batch, -1, -1).contiguous()

    # ---- transforms
    def ntt(self, h, nq: int, np_: int):
        out = h.clone()
        self._call("fhe_ntt_fwd", self._ptr(out), out.shape[0] * out.shape[1], nq, np_)
        return out

    def intt(self, h, nq: int, np_: int):
        out = h.clone()
        self._call("fhe_ntt_inv", self._ptr(out), out.shape[0] * out.shape[1], nq, np_)
        return out

    def from_i64(self, coeffs: np.ndarray, nq: int, with_p: bool):
        """coeffs: int64 [N] or [batch, N] -> NTT-domain [1, batch, limbs, N]."""
        np_ = self._K if with_p else 0
        c = np.ascontiguousarray(coeffs, dtype=np.int64).reshape(-1, self.n)
        b = c.shape[0]
        cd = torch.from_numpy(c).to(self.device)
        out = self._empty(1, b, nq + np_, self.n)
        self._call("fhe_from_i64", self._ptr(out), self._ptr(cd), nq, np_, b)
        self._call("fhe_ntt_fwd", self._ptr(out), b, nq, np_)
        return out

    # ---- elementwise.  a: [p, B, L, N]; b: [p or 1, B or 1, L, N] -- broadcasts are strides
    # inside the kernel, nothing is replicated.  If `a` is the smaller operand the roles swap
    # (add / mul commute); sub materialises.
    def _bin(self, name, a, b, nq, np_):
        full = (max(a.shape[0], b.shape[0]), max(a.shape[1], b.shape[1]))
        if tuple(a.shape[:2]) != full:
            if name != "fhe_sub" and tuple(b.shape[:2]) == full:
                a, b = b, a
            else:
                a = a.expand(full[0], full[1], -1, -1)
        a = a.contiguous()
        b = b.contiguous()
        out = torch.empty_like(a)
        self._call(name, self._ptr(out), self._ptr(a), self._ptr(b), a.shape[0], a.shape[1], b.shape[0], b.shape[1],
                   nq, np_)
        return out

    def add(self, a, b, nq, np_):
        return self._bin("fhe_add", a, b, nq, np_)

    def sub(self, a, b, nq, np_):
        return self._bin("fhe_sub", a, b, nq, np_)

    def mul(self, a, b, nq, np_):
        return self._bin("fhe_mul", a, b, nq, np_)

    def neg(self, a, nq, np_):
        a = a.contiguous()
        out = torch.empty_like(a)
        self._call("fhe_neg", self._ptr(out), self._ptr(a), a.shape[0] * a.shape[1], nq, np_)
        return out

    def _consts(self, cp: Sequence[int], cm: Sequence[int]):
        a = np.array([int(v) for v in cp], dtype=np.uint64)
        b = np.array([int(v) for v in cm], dtype=np.uint64)
        return a, b

    def mul_scalar(self, a, fac: Sequence[int], nq, np_):
        a = a.contiguous()
        f, _ = self._consts(fac, fac)
        out = torch.empty_like(a)
        self._call("fhe_mul_const", self._ptr(out), self._ptr(a), f.ctypes.data, f.ctypes.data,
                   a.shape[0] * a.shape[1], nq, np_)
        return out

    def mul_const(self, a, cp, cm, nq):
        a = a.contiguous()
        fp, fm = self._consts(cp, cm)
        out = torch.empty_like(a)
        self._call("fhe_mul_const", self._ptr(out), self._ptr(a), fp.ctypes.data, fm.ctypes.data,
                   a.shape[0] * a.shape[1], nq, 0)
        return out

    def add_const(self, a, cp, cm, nq, inplace: bool = False):
        """constant added to polynomial 0 of every ciphertext of the batch (inplace: `a` is a temporary the caller owns)"""
        out = a if inplace and a.is_contiguous() else a.clone()
        fp, fm = self._consts(cp, cm)
        self._call("fhe_add_const", self._ptr(out), self._ptr(out), fp.ctypes.data, fm.ctypes.data, a.shape[1], nq, 0)
        return out

    def add_poly0(self, a, p, nq):
        out = a.clone()
        p = self.expand_batch(p, a.shape[1]).contiguous()
        self._call("fhe_add", self._ptr(out), self._ptr(out), self._ptr(p), 1, a.shape[1], 1, a.shape[1], nq, 0)
        return out

    def add_poly0_inplace(self, a, p, nq):
        """as add_poly0 for a temporary `a` the caller owns"""
        p = self.expand_batch(p, a.shape[1]).contiguous()
        self._call("fhe_add", self._ptr(a), self._ptr(a), self._ptr(p), 1, a.shape[1], 1, a.shape[1], nq, 0)
        return a

    def add_into_polys(self, acc, p, nq):
        """acc[:p.shape[0]] += p in place (p broadcast over the batch if it has batch 1)"""
        k, bt = p.shape[0], acc.shape[1]
        p = self.expand_batch(p, bt).contiguous()
        self._call("fhe_add", self._ptr(acc), self._ptr(acc), self._ptr(p), k, bt, k, bt, nq, 0)
        return acc

    def tensor(self, a, b, nq):
        bt = max(a.shape[1], b.shape[1])
        a = self.expand_batch(a, bt).contiguous()
        b = self.expand_batch(b, bt).contiguous()
        out = self._empty(3, bt, nq, self.n)
        self._call("fhe_tensor", self._ptr(out), self._ptr(a), self._ptr(b), nq, bt)
        return out

    # ---- structural
    def rescale(self, h, nq):
        h = h.contiguous()
        out = self._empty(h.shape[0], h.shape[1], nq - 1, self.n)
        self._call("fhe_rescale", self._ptr(out), self._ptr(h), h.shape[0] * h.shape[1], nq)
        return out

    def automorphism(self, h, g: int, nq, np_):
        h = h.contiguous()
        out = torch.empty_like(h)
        self._call("fhe_automorphism", self._ptr(out), self._ptr(h), C.c_uint64(int(g)),
                   h.shape[0] * h.shape[1] * h.shape[2])
        return out

    def modup(self, d, nq):
        """d [1, B, nq, N] -> ext [B, beta, nq+K, N] (own-digit rows mirrored from d so the
        tensor is comparable with the oracle's ModUp output)."""
        d = d.contiguous()
        bt = d.shape[1]
        beta = self.params.digits_at(nq)
        ext = self._empty(bt, beta, nq + self._K, self.n)
        self._call("fhe_modup", self._ptr(ext), self._ptr(d), nq, bt)
        a = self.params.alpha
        for j in range(beta):
            lo, hi = j * a, min((j + 1) * a, nq)
            ext[:, j, lo:hi] = d[0, :, lo:hi]
        return ext

    def ks_inner(self, ext, d, ksk, nq):
        bt = d.shape[1]
        acc = self._empty(2, bt, nq + self._K, self.n)
        self._call("fhe_ks_inner", self._ptr(acc), self._ptr(ext.contiguous()), self._ptr(d.contiguous()),
                   self._ptr(ksk), nq, bt)
        return acc

    def moddown(self, acc, nq):
        acc = acc.clone()
        out = self._empty(acc.shape[0], acc.shape[1], nq, self.n)
        self._call("fhe_moddown", self._ptr(out), self._ptr(acc), nq, acc.shape[0] * acc.shape[1])
        return out

    def moddown_inplace(self, acc, nq):
        """as moddown, but the caller's acc (a temporary) is used as scratch"""
        out = self._empty(acc.shape[0], acc.shape[1], nq, self.n)
        self._call("fhe_moddown", self._ptr(out), self._ptr(acc), nq, acc.shape[0] * acc.shape[1])
        return out

    def relin_rescale(self, d3, ksk, nq):
        """[3,B,nq,N] tensor product -> [2,B,nq-1,N]: relinearise + rescale with one ModDown."""
        d3 = d3.contiguous()
        bt = d3.shape[1]
        out = self._empty(2, bt, nq - 1, self.n)
        self._call("fhe_relin_rescale", self._ptr(out), self._ptr(d3), self._ptr(ksk), nq, bt)
        return out

    def mul_relin_rescale(self, a, b, ksk, nq):
        """ct x ct + relinearise + rescale in one call: [2,B,>=nq,N] x [2,B,>=nq,N] -> [2,B,nq-1,N]; same residues
        as tensor() followed by relin_rescale(), without the 3-polynomial product in memory.  An operand with
        more than nq limbs (a higher level) is read in place, limbs 0..nq-1."""
        bt = max(a.shape[1], b.shape[1])
        a = self.expand_batch(a, bt).contiguous()
        b = self.expand_batch(b, bt).contiguous()
        out = self._empty(2, bt, nq - 1, self.n)
        self._call("fhe_mul_relin_rescale", self._ptr(out), self._ptr(a), a.shape[2], self._ptr(b), b.shape[2],
                   self._ptr(ksk), nq, bt)
        return out

    def mul_relin_rescale_gather(self, parts_a, parts_b, ksk, nq):
        """the same product with gathered operands (fhe_mul_relin_rescale_ptrs): parts_x = [(tensor [2,Bx,>=nq,N],
        batch indices), ...]; element i of the product multiplies the i-th listed batch element of side a with the
        i-th of side b -- slices, permutations, replications and concatenations of tensors without a copy."""
        es = self.n * 8

        def ptrs(parts):
            p0, p1, keep = [], [], []
            for t, idx in parts:
                t = t.contiguous()
                keep.append(t)
                step, base = t.shape[2] * es, t.data_ptr()
                poly1 = t.shape[1] * step
                for i in idx:
                    if not 0 <= i < t.shape[1]:
                        raise IndexError("mul_relin_rescale_gather: batch index out of range")
                    p0.append(base + i * step)
                    p1.append(base + poly1 + i * step)
            return p0, p1, keep
        a0, a1, keep_a = ptrs(parts_a)
        b0, b1, keep_b = ptrs(parts_b)
        bt = len(a0)
        if bt != len(b0) or bt == 0:
            raise ValueError("mul_relin_rescale_gather: the two sides list different numbers of batch elements")
        outs = []
        for lo in range(0, bt, 128):
            hi = min(bt, lo + 128)
            out = self._empty(2, hi - lo, nq - 1, self.n)
            arr = [(C.c_void_p * (hi - lo))(*p[lo:hi]) for p in (a0, a1, b0, b1)]
            self._call("fhe_mul_relin_rescale_ptrs", self._ptr(out), arr[0], arr[1], arr[2], arr[3], self._ptr(ksk), nq, hi - lo)
            outs.append(out)
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)

    def ks_accum(self, acc, d, ksk, lift, nq):
        """acc [2,B,nq+K,N] (or None) += <ModUp(d), ksk> + P * lift in the extended basis (no ModDown).
        d [1,B,nq,N] or None (lift only); lift [1 or 2,B,nq,N] or None.  Returns acc (updated in place)."""
        src = d if d is not None else lift
        bt = src.shape[1]
        fresh = acc is None
        if fresh:
            acc = self._empty(2, bt, nq + self._K, self.n)
        d = d.contiguous() if d is not None else None
        lift = lift.contiguous() if lift is not None else None
        self._call("fhe_ks_accum", self._ptr(acc), self._ptr(d) if d is not None else None,
                   self._ptr(ksk) if d is not None else None, self._ptr(lift) if lift is not None else None,
                   lift.shape[0] if lift is not None else 0, nq, bt, 0 if fresh else 1)
        return acc

    def moddown_rescale(self, acc, nq):
        """extended accumulator [2,B,nq+K,N] (a temporary: used as scratch) -> [2,B,nq-1,N] = round(acc / (P q_{nq-1}))"""
        out = self._empty(acc.shape[0], acc.shape[1], nq - 1, self.n)
        self._call("fhe_moddown_rescale", self._ptr(out), self._ptr(acc), nq, acc.shape[0] * acc.shape[1])
        return out

    def keyswitch(self, d, ksk, nq):
        d = d.contiguous()
        bt = d.shape[1]
        out = self._empty(2, bt, nq, self.n)
        self._call("fhe_keyswitch", self._ptr(out), self._ptr(d), self._ptr(ksk), nq, bt)
        return out

    # ---- fused LUT evaluation pieces
    def prepare_lincomb(self, const_res, c0_res, nq: int):
        """const_res[m][t] = (c_first[nq], c_second[nq]) residues; c0_res[m] likewise or None.
        Returns device tables for fhe_lincomb: (c, RN(c/q)) pairs as doubles."""
        from fractions import Fraction
        M, T = len(const_res), len(const_res[0])
        q = self.params.moduli
        tab = np.zeros((M, T, nq, 2, 2), dtype=np.float64)
        for m in range(M):
            for t in range(T):
                for h in range(2):
                    for j in range(nq):
                        c = int(const_res[m][t][h][j])
                        tab[m, t, j, h, 0] = float(c)
                        tab[m, t, j, h, 1] = float(Fraction(c, q[j]))
        dev = torch.from_numpy(tab).to(self.device)
        c0 = None
        if c0_res is not None:
            a = np.zeros((M, nq, 2), dtype=np.uint64)
            for m in range(M):
                for h in range(2):
                    a[m, :, h] = np.array([int(v) for v in c0_res[m][h]], dtype=np.uint64)
            c0 = torch.from_numpy(a.view(np.int64)).to(self.device)
        return dict(consts=dev, c0=c0, M=M, T=T, nq=nq)

    def _batch_view(self, x):
        """a tensor the strided kernels can read in place: contiguous, or a batch slice of a contiguous [2,B,L,N] tensor
        (every polynomial's part contiguous, the two parts B * L * N words apart); anything else is materialised"""
        if x.is_contiguous():
            return x
        L = x.shape[2]
        if x.dim() == 4 and x.shape[0] == 2 and x.stride(3) == 1 and x.stride(2) == self.n and x.stride(1) == L * self.n:
            return x
        return x.contiguous()

    def lincomb(self, inputs: List, prep) -> List:
        M, T, nq = prep["M"], prep["T"], prep["nq"]
        assert len(inputs) == T
        inputs = [self._batch_view(x) for x in inputs]
        bt = inputs[0].shape[1]
        ptrs = (C.c_void_p * T)(*[x.data_ptr() for x in inputs])
        nqs = (C.c_int * T)(*[x.shape[2] for x in inputs])
        pst = (C.c_longlong * T)(*[x.stride(0) for x in inputs])
        out = self._empty(M, 2, bt, nq, self.n)
        self._call("fhe_lincomb", self._ptr(out), ptrs, nqs, self._ptr(prep["consts"]),
                   self._ptr(prep["c0"]) if prep["c0"] is not None else None, M, T, nq, bt, pst)
        return [out[m] for m in range(M)]

    def mul_plain_sum(self, a_list: List, p_list: List, nq: int):
        """sum_t a_t (.) p_t in one pass per 16 terms; a_t [2,B,>=nq,N], p_t [1,1,nq,N] -> [2,B,nq,N]."""
        a_list = [x.contiguous() for x in a_list]
        p_list = [x.contiguous() for x in p_list]
        bt = a_list[0].shape[1]
        out = self._empty(2, bt, nq, self.n)
        done, T = 0, len(a_list)
        while done < T:
            g = min(16, T - done)
            ap = (C.c_void_p * g)(*[x.data_ptr() for x in a_list[done:done + g]])
            an = (C.c_int * g)(*[x.shape[2] for x in a_list[done:done + g]])
            pp = (C.c_void_p * g)(*[x.data_ptr() for x in p_list[done:done + g]])
            self._call("fhe_mul_plain_sum", self._ptr(out), ap, an, pp, g, nq, bt, 1 if done else 0)
            done += g
        return out

    def mul_plain_multi(self, a_list: List, p_rows: List[List], nq: int):
        """out_g = sum_t a_t (.) p_rows[g][t] for all g in one pass (None = absent term); a_t [2,B,>=nq,N],
        plaintexts [1,1,nq,N] -> [G,2,B,nq,N]."""
        a_list = [x.contiguous() for x in a_list]
        T, G = len(a_list), len(p_rows)
        bt = a_list[0].shape[1]
        keep = [[None if x is None else x.contiguous() for x in row] for row in p_rows]
        out = self._empty(G, 2, bt, nq, self.n)
        ap = (C.c_void_p * T)(*[x.data_ptr() for x in a_list])
        an = (C.c_int * T)(*[x.shape[2] for x in a_list])
        pp = (C.c_void_p * (G * T))(*[None if x is None else x.data_ptr() for row in keep for x in row])
        self._call("fhe_mul_plain_multi", self._ptr(out), ap, an, pp, T, G, nq, bt)
        return out

    def bsgs_inner(self, ext, ct, keys: List, galois: List[int], pt_rows: List[List], nq: int, acc=None):
        """double-hoisted baby steps (fhe_bsgs_inner): ext [B,beta,nq+K,N] = modup_raw(c1 of ct); ct [2,B,>=nq,N];
        keys[b] = switching-key tensor or None (no rotation); pt_rows[g][b] = extended-basis plaintext [1,1,nq+K,N] or
        None -> [G, 2, B, nq+K, N]; acc: the output of a previous pass (other baby steps) to add the sums to"""
        ct = ct.contiguous()
        ext = ext.contiguous()
        nb, G, bt = len(keys), len(pt_rows), ct.shape[1]
        keep = [[None if x is None else x.contiguous() for x in row] for row in pt_rows]
        kp = (C.c_void_p * nb)(*[None if k is None else k.data_ptr() for k in keys])
        gp = (C.c_uint64 * nb)(*[int(g) for g in galois])
        pp = (C.c_void_p * (G * nb))(*[None if x is None else x.data_ptr() for row in keep for x in row])
        out = self._empty(G, 2, bt, nq + self._K, self.n) if acc is None else acc
        self._call("fhe_bsgs_inner", self._ptr(out), self._ptr(ext), self._ptr(ct), ct.shape[2], kp, gp, pp, nb, G, nq, bt,
                   0 if acc is None else 1)
        return out

    def automorphism_rows(self, h, g: int):
        """X -> X^g on every row of an arbitrary [.., N] tensor (hoisted rotations act on ModUp output)."""
        h = h.contiguous()
        out = torch.empty_like(h)
        self._call("fhe_automorphism", self._ptr(out), self._ptr(h), C.c_uint64(int(g)), h.numel() // self.n)
        return out

    def modup_raw(self, d, nq):
        """fhe_modup without mirroring the digits' own rows (they are read from d by ks_inner)."""
        d = d.contiguous()
        bt = d.shape[1]
        ext = self._empty(bt, self.params.digits_at(nq), nq + self._K, self.n)
        self._call("fhe_modup", self._ptr(ext), self._ptr(d), nq, bt)
        return ext

    def tensor_acc(self, acc, a_list: List, b_list: List, nq: int, out=None, init=None):
        """acc [3,B,nq,N] (or None) += sum_g a_g (x) b_g ; a_g may carry more limbs than nq and
        either side may be a single ciphertext broadcast over the batch.  out: a batch slice [3, lo:hi] of a larger
        accumulator that receives the sums (acc must be None); init [2,B,nq,N]: a term the sums start from."""
        G = len(a_list)
        a_list = [self._batch_view(x) for x in a_list]
        b_batch = b_list[0].shape[1]
        bt = max(b_batch, max(x.shape[1] for x in a_list))
        # the inner sums usually are consecutive slices of one fhe_lincomb output: use them in place
        step = b_list[0].numel() * 8
        if all(x.is_contiguous() and x.shape == b_list[0].shape and x.data_ptr() == b_list[0].data_ptr() + k * step
               for k, x in enumerate(b_list)):
            b_base = b_list[0].data_ptr()
            keep = b_list                                           # keeps the storage alive during the calls
        else:
            keep = torch.stack([x.contiguous() for x in b_list], dim=0).contiguous()
            b_base = keep.data_ptr()
        accumulate = 1
        if acc is None:
            acc = self._empty(3, bt, nq, self.n) if out is None else out
            accumulate = 0
        elif out is not None:
            raise ValueError("tensor_acc: out and acc exclude each other")
        if not acc.is_contiguous() and not (acc.stride(3) == 1 and acc.stride(2) == self.n and acc.stride(1) == nq * self.n
                                            and acc.shape[2] == nq):
            raise ValueError("tensor_acc: accumulator must be contiguous or a batch slice of a contiguous tensor")
        acc_ps = acc.stride(0)
        init_p = None
        if init is not None:
            init = self.expand_batch(init, bt).contiguous()
            if tuple(init.shape) != (2, bt, nq, self.n):
                raise ValueError("tensor_acc: init must be [2, batch, nq, N]")
            init_p = self._ptr(init)
        done = 0
        while done < G:
            g = min(16, G - done)
            ptrs = (C.c_void_p * g)(*[x.data_ptr() for x in a_list[done:done + g]])
            nqs = (C.c_int * g)(*[x.shape[2] for x in a_list[done:done + g]])
            abs_ = (C.c_int * g)(*[x.shape[1] for x in a_list[done:done + g]])
            pst = (C.c_longlong * g)(*[x.stride(0) for x in a_list[done:done + g]])
            self._call("fhe_tensor_acc", self._ptr(acc), ptrs, nqs, abs_, C.c_void_p(b_base + done * step), b_batch, g, nq, bt,
                       accumulate, pst, acc_ps, init_p if done == 0 else None)
            accumulate = 1
            done += g
        return acc

    # ---- device-side encode / decode / sampling (throughput path; the default host path keeps
    # GPU and oracle ciphertexts bit-identical for the parity tests)
    def _enc_tables(self):
        t = getattr(self, "_enc_tab", None)
        if t is None:
            from . import encoding
            k_pos, k_neg, twist = encoding._tables(self.params.log_n)
            t = (torch.from_numpy(k_pos).to(self.device), torch.from_numpy(k_neg).to(self.device),
                 torch.from_numpy(twist).to(self.device))
            self._enc_tab = t
        return t

    def encode_device(self, values: torch.Tensor, scale: float):
        """complex128 [B, slots] (device) -> NTT-domain plaintext [1, B, nq?]: returns int64 coefficients [B, N]"""
        k_pos, k_neg, twist = self._enc_tables()
        bt = values.shape[0]
        spec = torch.empty((bt, self.n), dtype=torch.complex128, device=self.device)
        spec[:, k_pos] = values
        spec[:, k_neg] = values.conj()
        m = (torch.fft.fft(spec, dim=1) * twist.conj() / self.n).real * scale
        return torch.round(m).to(torch.int64)

    def from_i64_device(self, coeffs: torch.Tensor, nq: int, with_p: bool):
        np_ = self._K if with_p else 0
        coeffs = coeffs.contiguous()
        out = self._empty(1, coeffs.shape[0], nq + np_, self.n)
        self._call("fhe_from_i64", self._ptr(out), self._ptr(coeffs), nq, np_, coeffs.shape[0])
        self._call("fhe_ntt_fwd", self._ptr(out), coeffs.shape[0], nq, np_)
        return out

    def decode_device(self, h, use: int, scale: float) -> torch.Tensor:
        """coefficient-domain [1, B, use, N] -> complex128 slots [B, slots] on the device"""
        k_pos, _, twist = self._enc_tables()
        h = h.contiguous()
        bt = h.shape[1]
        m = torch.empty((bt, self.n), dtype=torch.float64, device=self.device)
        self._call("fhe_crt_centered", self._ptr(m), self._ptr(h), use, bt)
        self.check_status()                       # decrypt boundary: nothing upstream may have failed silently
        spec = torch.fft.ifft(m * twist, dim=1) * self.n
        return spec[:, k_pos] / scale

    def crt_centered(self, h, use: int) -> np.ndarray:
        """h [1, B, use, N] coefficient domain -> float64 [B, N]"""
        h = h.contiguous()
        bt = h.shape[1]
        out = torch.empty((bt, self.n), dtype=torch.float64, device=self.device)
        self._call("fhe_crt_centered", self._ptr(out), self._ptr(h), use, bt)
        res = out.cpu().numpy()
        self.check_status()
        return res
