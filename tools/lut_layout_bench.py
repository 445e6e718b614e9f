"""Does the placement of the LUT inputs matter?  The S-box outer sums (fused._outer_sum: one fhe_lincomb with 16 inputs
and 128 outputs, eight fhe_tensor_acc with 16 operands each, eight relinearisations) on synthetic ciphertexts of the shape
SubBytes sees after a refresh (batch 8, level 9), with the 15 + 15 monomials laid out three ways:
  copies     : every monomial its own allocation (what the service does)
  views      : batch slices of one tensor per basis (strided inputs, FHE_LUT_VIEWS)
  staggered  : copies placed in one arena at offsets that are NOT multiples of 2 MiB (each shifted by a further 264 KiB)
CUDA events, median of 7, L2 flushed between runs."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from aes_fhe_b200.engine import Ciphertext
    from aes_fhe_b200.fused import _outer_sum
    from aes_fhe_b200.params import LOG_PQ_BUDGET_SPARSE, make_params
    from aes_fhe_b200.services.aes_bits import sbox_walsh
    from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
    P = make_params(16, 26, scale_bits=44, log_pq_budget=LOG_PQ_BUDGET_SPARSE)
    w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, seed=3, device_codec=True), rotation_steps=[])
    eng = w.engine
    W = sbox_walsh()
    lvl, bt, n = 9, 8, P.n
    L = lvl + 1
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def rnd(*shape):
        return torch.randint(0, 2 ** 40, shape, dtype=torch.int64, device="cuda", generator=g)

    big_hi, big_lo = rnd(2, 15 * bt, L, n), rnd(2, 15 * bt, L, n)

    def layout(kind):
        out = []
        for big in (big_hi, big_lo):
            d = {}
            if kind == "staggered":
                words = 2 * bt * L * n
                pad = 264 * 1024 // 8
                arena = torch.empty(15 * (words + pad) + pad, dtype=torch.int64, device="cuda")
            for m in range(1, 16):
                v = big[:, (m - 1) * bt:m * bt]
                if kind == "copies":
                    t = v.contiguous()
                elif kind == "views":
                    t = v
                else:
                    off = (m - 1) * (words + pad) + pad
                    t = arena[off:off + words].view(2, bt, L, n)
                    t.copy_(v)
                d[m] = Ciphertext(eng, t, lvl)
            out.append(d)
        return out

    res = {}
    for kind in ("copies", "views", "staggered", "copies"):
        hi, lo = layout(kind)
        ts = []
        for it in range(9):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            outs = _outer_sum(eng, w.relin_key, hi, lo, [W[k] for k in range(8)], ("sbox-bits",))
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        ts.sort()
        chk = int(sum(int(o.polys.sum().item()) for o in outs) & 0xFFFFFFFF)
        print(kind, "ms median", round(ts[len(ts) // 2], 3), "min", round(ts[0], 3), "checksum", chk, flush=True)


if __name__ == "__main__":
    main()
