"""Where do the torch copy kernels of one AES-128 pass come from?  Bytes moved by every copy-producing torch call
(contiguous of a view, clone, cat, stack, index_select), grouped by the call site inside this repo.
    python tools/copy_sites.py [--states 1]"""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=1)
    args = ap.parse_args()
    from aes_fhe_b200.params import LOG_PQ_BUDGET_SPARSE, make_params
    from aes_fhe_b200.services.aes_bits import AESBitService
    from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
    P = make_params(16, 26, scale_bits=44, log_pq_budget=LOG_PQ_BUDGET_SPARSE)
    w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, seed=3, device_codec=True), rotation_steps=[])
    svc = AESBitService(w)
    G = args.states
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    blocks = np.random.default_rng(9).integers(0, 256, (G * svc.Bs, 16), dtype=np.uint8)
    fresh = svc.best_fresh_level()
    plan = svc.plan_levels(fresh)
    st = svc.encrypt_state(blocks, level=fresh)
    rkeys = svc.encrypt_round_keys(key, G, plan)
    svc.encrypt_blocks(st, key, round_keys=rkeys)
    torch.cuda.synchronize()
    # count the bytes every copy-producing torch call moves, by call site inside this repo
    import traceback
    sites = defaultdict(lambda: [0, 0])

    def where():
        fr = [f for f in traceback.extract_stack()[:-2] if "/aes_fhe_b200/" in f.filename]
        return " <- ".join(f"{Path(f.filename).name}:{f.lineno}({f.name})" for f in reversed(fr[-4:]))

    def note(kind, nbytes):
        k = (kind, where())
        sites[k][0] += 1
        sites[k][1] += int(nbytes)

    T = torch.Tensor
    o_contig, o_clone, o_isel, o_cat, o_gather, o_stack = T.contiguous, T.clone, T.index_select, torch.cat, torch.gather, torch.stack

    def contiguous(self, *a, **k):
        if self.is_cuda and not self.is_contiguous():
            note("contiguous", 2 * self.numel() * self.element_size())
        return o_contig(self, *a, **k)

    def clone(self, *a, **k):
        if self.is_cuda:
            note("clone", 2 * self.numel() * self.element_size())
        return o_clone(self, *a, **k)

    def index_select(self, dim, idx):
        out = o_isel(self, dim, idx)
        if self.is_cuda:
            note("index_select", 2 * out.numel() * out.element_size())
        return out

    def cat(ts, *a, **k):
        out = o_cat(ts, *a, **k)
        if out.is_cuda:
            note("cat", 2 * out.numel() * out.element_size())
        return out

    def stack(ts, *a, **k):
        out = o_stack(ts, *a, **k)
        if out.is_cuda:
            note("stack", 2 * out.numel() * out.element_size())
        return out

    T.contiguous, T.clone, T.index_select, torch.cat, torch.stack = contiguous, clone, index_select, cat, stack
    try:
        svc.encrypt_blocks(st, key, round_keys=rkeys)
        torch.cuda.synchronize()
    finally:
        T.contiguous, T.clone, T.index_select, torch.cat, torch.stack = o_contig, o_clone, o_isel, o_cat, o_stack
    total = sum(v[1] for v in sites.values())
    print(f"# copy-producing torch calls of one AES-128 pass ({G} state(s)): {total / 1e9:.1f} GB moved (read + write)")
    print("| call | calls | GB | call site (innermost first) |\n|---|---|---|---|")
    for (name, wh), (n, b) in sorted(sites.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {name} | {n} | {b / 1e9:.2f} | {wh} |")


if __name__ == "__main__":
    main()
