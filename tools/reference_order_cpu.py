#!/usr/bin/env python
"""CPU timing of the reference's OWN, UNMODIFIED Python services (imported from /root/reference through the `desilofhe`
shim, aes_fhe_b200/compat.py) on the oracle port: `XORService.xor_cipher` (92 key switches) and
`SBoxService.sub_bytes_array` (255 key switches) at N = 2^16 -- the stock operation order that BASELINE.md section 4
promises beside the GPU numbers.  Only runs where the reference tree is mounted (this container, not the GPU box), so its
output is committed under profiles/ with the core count of the machine it ran on.

    python tools/reference_order_cpu.py [--skip-sbox] > profiles/r02_reference_order_cpu.json
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REFERENCE = Path("/root/reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-sbox", action="store_true")
    args = ap.parse_args()
    if not REFERENCE.exists():
        print(json.dumps({"unavailable": "reference tree not mounted"}))
        return
    from aes_fhe_b200 import compat
    from aes_fhe_b200 import engine as E
    from aes_fhe_b200.params import make_params
    from oracle.refmod import RefBackend, build
    build()
    threads = len(os.sched_getaffinity(0))
    cfg = {"params": make_params(16, 30)}

    class OracleEngine(E.Engine):
        def __init__(self, *a, **k):
            P = cfg["params"]
            k.pop("max_level", None)
            super().__init__(_backend=RefBackend(P, threads=threads), _params=P)

        def create_rotation_key(self, sk, steps=None):
            return super().create_rotation_key(sk, steps=[])

    compat.install(OracleEngine)
    compat.mount_reference(REFERENCE)
    import aes_xor_fhe.xor_service as xs
    out = {"machine_cores": threads, "kind": "reference-python-on-port", "n": 65536,
           "note": "the reference's unmodified service code on oracle/refmod.cpp; one ciphertext = 2048 AES blocks"}
    c = xs.XORConfig(coeffs_path=REFERENCE / "xor_mono_coeffs.json")
    eng = xs.EngineWrapper(c)
    svc = xs.XORService(eng, xs.CoefficientCache(c.coeffs_path))
    rng = np.random.default_rng(0)
    sc = eng.engine.slot_count
    a = rng.integers(0, 16, size=sc, dtype=np.uint8)
    b = rng.integers(0, 16, size=sc, dtype=np.uint8)
    ea, eb = eng.encrypt(xs.ZetaEncoder.to_zeta(a)), eng.encrypt(xs.ZetaEncoder.to_zeta(b))
    svc.coeff_cache.get_plaintext_coeffs(eng)                       # the 64 plaintext encodings, outside the timed call
    k0 = dict(eng.engine.op_counts)
    t0 = time.perf_counter()
    res = svc.xor_cipher(ea, eb)
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(xs.ZetaEncoder.from_zeta(eng.decrypt(res)), a ^ b))
    ks = sum(v - k0.get(k, 0) for k, v in eng.engine.op_counts.items() if k.startswith("keyswitch"))
    out["xor_cipher"] = {"seconds": dt, "key_switches": int(ks), "nibbles_equal_a_xor_b": ok, "max_level": 30,
                         "call": "XORService.xor_cipher (xor_service.py:271-286), test_xor_random inputs"}
    if not args.skip_sbox:
        import aes_xor_fhe.sbox.sbox_service as sb
        import aes_xor_fhe.engine_context as ec
        from aes_xor_fhe.utils import zeta_decode, zeta_encode
        cfg["params"] = make_params(16, 22)
        ctx = ec.EngineContext(signature=2, max_level=22, mode="parallel", thread_count=threads, device_id=0)
        s2 = sb.SBoxService(ctx)
        x = np.tile(np.arange(256, dtype=np.uint8), sc // 256 + 1)[:sc]
        ct = ctx.engine.encrypt(zeta_encode(x.astype(np.int64), modulus=256), ctx.public_key)
        k0 = dict(ctx.engine.op_counts)
        t0 = time.perf_counter()
        o = s2.sub_bytes_array(ct)
        dt = time.perf_counter() - t0
        ok = bool(np.array_equal(zeta_decode(ctx.engine.decrypt(o, ctx.secret_key), modulus=256), np.array(sb.AES_SBOX, dtype=np.uint8)[x]))
        ks = sum(v - k0.get(k, 0) for k, v in ctx.engine.op_counts.items() if k.startswith("keyswitch"))
        out["sub_bytes_array"] = {"seconds": dt, "key_switches": int(ks), "bytes_equal_sbox": ok, "max_level": 22,
                                  "blocks_per_s": 2048 / dt,
                                  "call": "SBoxService.sub_bytes_array (sbox/sbox_service.py:116-138), test_sbox_array_simd inputs"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
