"""Device time per kernel of one SubBytes step (the bench workload), taken with CUPTI through
torch.profiler -- concurrent-safe, no replay, so the sum is the real busy time of the step.

    python tools/kernel_breakdown.py [--batch 16] [--dnum D] [--lib path/to/other/libaesfhe_b200.so]

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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--dnum", type=int, default=None)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--json", default=None)
    ap.add_argument("--workload", default="sbox", choices=["sbox", "aes128"],
                    help="sbox: the bench step (SubBytes, BASELINE configs[1]); aes128: ten rounds with refreshes (configs[4])")
    args = ap.parse_args()
    if args.lib:
        from aes_fhe_b200 import _capi
        _capi.LIB_PATH = Path(args.lib).resolve()
    import bench
    bench.DNUM = args.dnum
    from aes_fhe_b200.services.xor_service import ZetaEncoder
    if args.workload == "sbox":
        ctx, svc = bench._make_service(seed=1)
        eng = ctx.engine
        data = bench._inputs(eng.slot_count, args.batch, seed=0)
        ct = eng.encrypt(ZetaEncoder.to_zeta(data, 256), ctx.public_key)

        def step():
            return svc.sub_bytes_array_bsgs(ct)
        warm = 2
    else:
        import numpy as np
        from aes_fhe_b200.services.aes128 import AES128Service
        from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache
        cfg = XORConfig()
        w = EngineWrapper(cfg, _engine_kwargs=dict(seed=3), rotation_steps=[])
        aes = AES128Service(w, XORService(w, CoefficientCache(cfg.coeffs_path)))
        key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
        rng = np.random.default_rng(9)
        st = aes.encrypt_state([rng.integers(0, 256, (aes.B, 16), dtype=np.uint8) for _ in range(args.batch)])

        def step():
            return aes.encrypt_blocks(st, key)
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
