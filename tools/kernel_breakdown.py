"""Device time per kernel of one bench step (bit-sliced AES-128, or SubBytes), taken with CUPTI through
torch.profiler -- concurrent-safe, no replay, so the sum is the real busy time of the step.

    python tools/kernel_breakdown.py [--workload aes128|sbox] [--batch B] [--dnum D] [--lib path/to/other/libaesfhe_b200.so]

`--lib` loads another build of the library (A/B runs against aes_fhe_b200/csrc/variants/*.so).
Prints a markdown table; `--json FILE` also writes {kernel: [launches, total_us]}.
"""
import argparse
import json
import sys
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=None, help="sbox: ciphertexts (default 16); aes128: states of 8192 blocks (default 1)")
    ap.add_argument("--dnum", type=int, default=None)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--json", default=None)
    ap.add_argument("--workload", default="aes128", choices=["sbox", "aes128"],
                    help="aes128: the bench step (bit-sliced AES-128, BASELINE configs[4]); sbox: SubBytes (configs[1])")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 16 if args.workload == "sbox" else 1
    if args.lib:
        from aes_fhe_b200 import _capi
        _capi.LIB_PATH = Path(args.lib).resolve()
    import numpy as np
    if args.workload == "sbox":
        from aes_fhe_b200.params import make_params
        from aes_fhe_b200.services.engine_context import EngineContext
        from aes_fhe_b200.services.sbox_service import SBoxService
        from aes_fhe_b200.services.xor_service import ZetaEncoder
        P = make_params(16, 22) if args.dnum is None else make_params(16, 22, dnum=args.dnum)
        ctx = EngineContext(signature=2, max_level=22, mode="parallel", _engine_kwargs=dict(_params=P, seed=1), rotation_steps=[])
        svc = SBoxService(ctx)
        eng = ctx.engine
        data = np.random.default_rng(0).integers(0, 256, (args.batch, eng.slot_count), dtype=np.uint8)
        ct = eng.encrypt(ZetaEncoder.to_zeta(data, 256), ctx.public_key)

        def step():
            return svc.sub_bytes_array_bsgs(ct)
        warm = 2
    else:
        import bench
        from aes_fhe_b200.services.aes_bits import AESBitService
        from aes_fhe_b200.services.key_expansion import expand_key
        from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
        P = bench.aes_params()
        w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, seed=3, device_codec=True), rotation_steps=[])
        aes = AESBitService(w)
        G = args.batch
        key = bytes.fromhex(bench.KEY_HEX)
        rks = expand_key(key)
        fresh = aes.best_fresh_level()
        plan = aes.plan_levels(fresh)
        st = aes.encrypt_state(bench.rank_blocks(0, G * aes.Bs), level=fresh)
        rkeys = aes.encrypt_round_keys(key, G, plan)

        def step():
            return aes.encrypt_blocks(st, key, round_keys=rkeys)
        warm = 1
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    torch.cuda.synchronize()
    step_ms = a.elapsed_time(b) / args.steps
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = agg[ev.name]
            k[0] += 1
            k[1] += ev.device_time
    total = sum(v[1] for v in agg.values())
    print(f"# step {step_ms:.2f} ms (events, no profiler); kernel busy time under CUPTI {total / args.steps / 1e3:.2f} ms per step, "
          f"batch {args.batch}, dnum {args.dnum}, lib {args.lib or 'default'}")
    print("| kernel | launches/step | us/step | share | avg us |")
    print("|---|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n / args.steps:.0f} | {us / args.steps:.1f} | {100 * us / total:.1f}% | {us / n:.1f} |")
    if args.json:
        json.dump({k: [v[0] / args.steps, v[1] / args.steps] for k, v in agg.items()}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
