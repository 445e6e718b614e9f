#!/usr/bin/env python
"""Work of one bit-sliced AES-128 (one state, ten rounds, ten bit bootstraps) in length-N transforms, counted by
the CPU oracle running the whole schedule at a small ring with the limb structure of the N = 2^16 parameter set
(25 + 9 limbs, three digits of nine): the schedule's row count does not depend on N.  bench.py scales its bounded
CPU sample (seconds per transform row at N = 2^16) to a whole AES-128 with this number.

    python tools/aes_bits_work.py          # prints the small-ring count

The committed tests/golden/aes_bits_work.json holds the N = 2^16 count measured by bench.py on the B200 (the special
FFT has 15 butterfly layers there, 10 at N = 2^11, so the bootstrap's linear transforms have more diagonals and the
count is larger); this script reproduces the small-ring figure recorded beside it.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def measure(log_n: int = 11):
    from aes_fhe_b200.params import make_params
    from aes_fhe_b200.services.aes_bits import AESBitService
    from aes_fhe_b200.services.key_expansion import expand_key
    from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
    from oracle import aes_plain as A
    from oracle.refmod import RefBackend
    P = make_params(log_n, 24, dnum=3, scale_bits=44)
    be = RefBackend(P)
    w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, _backend=be, seed=2), rotation_steps=[])
    svc = AESBitService(w)
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    rks = expand_key(key)
    blocks = np.random.default_rng(1).integers(0, 256, (svc.Bs, 16), dtype=np.uint8)
    st = svc.encrypt_state(blocks, level=1 + svc.boot_in_levels)
    rkeys = [svc.encrypt_round_key(rks[r], level=12, half=(r == 10)) for r in range(11)]
    svc.prepare_keys()
    r0, c0 = be.ntt_row_count(), dict(w.engine.op_counts)
    out = svc.encrypt_blocks(st, key, round_keys=rkeys)
    rows = be.ntt_row_count() - r0
    assert np.array_equal(svc.decrypt_state(out), A.encrypt_blocks(blocks, key))
    counts = {k: v - c0.get(k, 0) for k, v in w.engine.op_counts.items() if v - c0.get(k, 0)}
    return {"limbs": [P.n_q, P.n_p, P.alpha, P.dnum], "max_level": 24, "boot_groups": [3, 3],
            "ntt_rows_per_state_aes128": int(rows), "blocks_per_state_at_2_16": 8192,
            "bootstrapped_ciphertexts": int(svc.refreshes), "batched_op_calls": counts,
            "counted_on": f"oracle/refmod.cpp at N=2^{log_n}, plaintext encodings and keys excluded"}


if __name__ == "__main__":
    d = measure()
    print(json.dumps(d))
