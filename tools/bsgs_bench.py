#!/usr/bin/env python
"""Micro-benchmark of fhe_bsgs_inner (the fused baby-step kernel of the double-hoisted linear transforms) at the
shape of a CoeffToSlot factor of the AES-128 parameter set: N = 2^16, 25 + 9 limbs, 3 digits, 16 baby steps x 4 giant
steps, batch 16.  Prints one JSON line: time per launch, FP64 and traffic figures derived from the shape.

    python tools/bsgs_bench.py [--batch 16] [--nq 25] [--babies 16] [--giants 4]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--nq", type=int, default=25)
    ap.add_argument("--babies", type=int, default=16)
    ap.add_argument("--giants", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import bench
    from aes_fhe_b200.backend_cuda import CudaBackend
    P = bench.aes_params()
    be = CudaBackend(P)
    nq, K, n = args.nq, P.n_p, P.n
    beta, ne, tot = P.digits_at(nq), nq + P.n_p, P.n_q + P.n_p
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda *shape: torch.randint(0, 2 ** 39, shape, dtype=torch.int64, device="cuda", generator=g)   # noqa: E731
    ext = rnd(args.batch, beta, ne, n)
    ct = rnd(2, args.batch, nq, n)
    keys = [None] + [rnd(P.dnum, 2, 1, tot, n) for _ in range(args.babies - 1)]
    gal = [1] + [pow(5, 3 * b + 1, 2 * n) for b in range(1, args.babies)]
    pts = [[rnd(1, 1, ne, n) for _ in range(args.babies)] for _ in range(args.giants)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(args.reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = be.bsgs_inner(ext, ct, keys, gal, pts, nq)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    work = args.babies * args.batch * ne * n
    fp64 = work * (2 * beta * 7 + 7 + 2 * args.giants * 7 + beta + 1)
    hbm = (ext.numel() + 2 * args.batch * nq * n + (args.babies - 1) * 2 * beta * ne * n + args.giants * args.babies * ne * n
           + out.numel()) * 8
    print(json.dumps({"ms": ms, "batch": args.batch, "nq": nq, "beta": beta, "babies": args.babies, "giants": args.giants,
                      "fp64_gops": fp64 / 1e9, "fp64_frac_of_18.4T": fp64 / (ms * 1e-3) / 18.4e12,
                      "hbm_algorithmic_gb": hbm / 1e9, "hbm_gbs": hbm / (ms * 1e-3) / 1e9,
                      "variant": __import__("os").environ.get("FHE_BSGS_VARIANT", "0")}))


if __name__ == "__main__":
    main()
