"""world_size-2 gloo run of the sharded path on CPU: rank 0 alone generates keys (OS entropy) and keeps the secret;
public / relinearisation / conjugation key and every Galois key (ShiftRows, bootstrap transforms) are broadcast;
each rank runs AddRoundKey_0 + one bit-sliced AES round with a bit bootstrap on its own state; the result
ciphertexts are gathered and decrypted by the key owner.  No collective on the data path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aes_fhe_b200.sharding import shard_range


def test_shard_range_partitions_everything():
    for n in (1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                got += list(range(lo, hi))
            assert got == list(range(n))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aes_fhe_b200.engine import Engine
        from aes_fhe_b200.params import make_params
        from aes_fhe_b200.services.aes_bits import AESBitService
        from aes_fhe_b200.services.key_expansion import expand_key
        from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
        from aes_fhe_b200.sharding import distribute_keys, gather_ciphertexts, ReceivedKeys
        from oracle.refmod import RefBackend
        from oracle import aes_plain as A
        P = make_params(11, 24, scale_bits=44)
        key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
        rks = expand_key(key)
        if rank == 0:
            # the key owner: OS-entropy keys (no seed), every Galois key of the service issued before shipping
            w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, _backend=RefBackend(P, threads=2)), rotation_steps=[])
            engine = w.engine
            AESBitService(w).prepare_keys()
        else:
            w = None
            engine = Engine(_params=P, _backend=RefBackend(P, threads=2), use_bootstrap=True)
        ctx, stats = distribute_keys(engine, w, src=0)
        assert stats["keys"] >= 3 + 3 and stats["bytes"] > 0
        if rank:
            assert isinstance(ctx.secret_key, ReceivedKeys) and not hasattr(ctx.secret_key, "coeffs")
            try:
                ctx.decrypt(None)
                raise AssertionError("a non-owner rank must not decrypt")
            except RuntimeError:
                pass
        svc = AESBitService(ctx)             # rank 1: bootstrap plan and ShiftRows keys come from the received set
        svc.prepare_keys()
        # each rank: its own state through ARK_0, ShiftRows, a bit bootstrap, SubBytes, MixColumns + ARK
        rng = np.random.default_rng(50 + rank)
        blocks = rng.integers(0, 256, (svc.Bs, 16), dtype=np.uint8)
        st = svc.encrypt_state(blocks, level=1 + svc.boot_in_levels)
        out = svc.encrypt_blocks(st, key, rounds=1)
        assert engine.op_counts["bootstrap"] == 1
        parts = gather_ciphertexts(engine, out, dst=0)
        if rank == 0:
            ok = len(parts) == world
            for r, ct in enumerate(parts):
                want = A.round_fn(np.random.default_rng(50 + r).integers(0, 256, (svc.Bs, 16), dtype=np.uint8) ^ rks[0], rks[1])
                ok &= bool(np.array_equal(svc.decrypt_state(ct), want))
            q.put(ok)
        else:
            assert parts is None
    finally:
        dist.destroy_process_group()


@pytest.mark.slow
def test_two_rank_gloo_key_owner_and_sharded_round():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=900)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
