#!/usr/bin/env python
"""Full-size probe of the bit-sliced AES-128 path on one B200: N = 2^16, G states of 8192 blocks, per-stage
device time, every decoded block checked against plain AES.

    python tools/aes_bits_probe.py [--states G] [--level L] [--groups 3,3] [--rounds 10]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


class StageTimer:
    """CUDA events at stage boundaries on the current stream; `ms()` after a synchronize"""

    def __init__(self):
        import torch
        self.torch = torch
        self.marks = []
        self.peaks = {}
        self.reset()

    def reset(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks = [("start", e)]

    def __call__(self, name):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks.append((name, e))
        self.peaks[name] = max(self.peaks.get(name, 0.0), self.torch.cuda.max_memory_allocated() / 2 ** 30)
        self.torch.cuda.reset_peak_memory_stats()

    def ms(self):
        self.torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=1)
    ap.add_argument("--level", type=int, default=26)
    ap.add_argument("--scale-bits", type=int, default=44)
    ap.add_argument("--groups", default="3,3")
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--fresh-level", type=int, default=None)
    ap.add_argument("--dnum", type=int, default=0)
    ap.add_argument("--nvtx", action="store_true",
                    help="NVTX ranges between the stage / bootstrap-phase boundaries (range 'after_<boundary>'), for "
                         "ncu --nvtx --nvtx-include; e.g. after_boot_mod_raise = CoeffToSlot, after_boot_conjugate_split = EvalMod")
    args = ap.parse_args()
    import torch
    from aes_fhe_b200.params import make_params
    from aes_fhe_b200.services.aes_bits import AESBitService
    from aes_fhe_b200.services.key_expansion import expand_key
    from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
    from oracle import aes_plain as A

    from aes_fhe_b200.params import LOG_PQ_BUDGET_SPARSE
    P = make_params(16, args.level, scale_bits=args.scale_bits, dnum=args.dnum, log_pq_budget=LOG_PQ_BUDGET_SPARSE)
    groups = tuple(int(x) for x in args.groups.split(","))
    t0 = time.time()
    w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, seed=3, device_codec=True), rotation_steps=[])
    svc = AESBitService(w, boot_groups=groups)
    G = args.states
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    rng = np.random.default_rng(9)
    blocks = rng.integers(0, 256, (G * svc.Bs, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(bytes.fromhex("3243f6a8885a308d313198a2e0370734"), np.uint8)
    rks = expand_key(key)
    fresh = args.fresh_level if args.fresh_level is not None else svc.best_fresh_level(args.rounds)
    plan = svc.plan_levels(fresh, args.rounds)
    st = svc.encrypt_state(blocks, level=fresh)
    rkeys = svc.encrypt_round_keys(key, G, plan, rounds=args.rounds)
    for _ in range(max(1, args.warmup)):                                               # warm-up: keys, matrices, tables, and the
        out = svc.encrypt_blocks(st, key, rounds=args.rounds, round_keys=rkeys)        # caching allocator's growth (one pass is not
    torch.cuda.synchronize()                                                           # enough after a change of tensor lifetimes)
    setup = time.time() - t0
    c0, r0 = dict(w.engine.op_counts), svc.refreshes
    l0 = w.engine.backend.launch_count()
    resident = torch.cuda.memory_allocated() / 2 ** 30
    peak_all = torch.cuda.max_memory_allocated() / 2 ** 30
    torch.cuda.reset_peak_memory_stats()
    svc.timer = StageTimer()
    w.engine.phase_timer = svc.timer
    if args.nvtx:
        import torch.cuda.nvtx as nvtx
        timer, opened = svc.timer, [False]

        def hook(name):
            timer(name)
            if opened[0]:
                nvtx.range_pop()
            nvtx.range_push("after_" + name.replace(":", "_"))
            opened[0] = True
        svc.timer = w.engine.phase_timer = hook
        hook.ms, hook.peaks = timer.ms, timer.peaks
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = svc.encrypt_blocks(st, key, rounds=args.rounds, round_keys=rkeys)
    b.record()
    if args.nvtx:
        nvtx.range_pop()
    stages = svc.timer.ms()
    peaks = svc.timer.peaks
    svc.timer = None
    w.engine.phase_timer = None
    ms = a.elapsed_time(b)
    got = svc.decrypt_state(out)
    s = blocks ^ rks[0]
    for r in range(1, args.rounds + 1):
        s = A.round_fn(s, rks[r], last=(r == 10))
    slots = svc.decrypt_slots(out)
    err = float(np.abs(slots - (1.0 - 2.0 * svc.pack_bits(s))).max())
    cnt = {k: v - c0.get(k, 0) for k, v in w.engine.op_counts.items()}
    print(json.dumps({"states": G, "blocks": int(G * svc.Bs), "level": args.level, "groups": groups, "rounds": args.rounds,
                      "limbs": [P.n_q, P.n_p, P.dnum], "ms": ms, "blocks_per_s": G * svc.Bs / (ms * 1e-3), "stages_ms": stages,
                      "bytes_equal_plain_aes": bool(np.array_equal(got, s)), "max_slot_err": err,
                      "bootstrapped_ciphertexts": svc.refreshes - r0, "op_counts": cnt,
                      "launches": w.engine.backend.launch_count() - l0, "out_level": out.level, "fresh_level": fresh,
                      "refresh_before_rounds": plan["refresh_before_rounds"], "setup_s": setup,
                      "mem_gb_peak_first_run": peak_all, "mem_gb_resident_after_warmup": resident,
                      "mem_gb_peak_by_stage": peaks}))


if __name__ == "__main__":
    main()
