"""world_size-2 gloo run of the sharded path on CPU: rank 0 owns the keys, evaluation keys are
broadcast, each rank evaluates the XOR LUT on its own shard of ciphertext batches, results are
gathered and decrypted by the key owner.  No collective on the data path."""
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
        from aes_fhe_b200.params import make_params
        from aes_fhe_b200.services.xor_service import (XORService, EngineWrapper, XORConfig, CoefficientCache, ZetaEncoder)
        from aes_fhe_b200.sharding import broadcast_evaluation_keys, gather_handles
        from aes_fhe_b200.engine import Ciphertext
        from oracle.refmod import RefBackend
        P = make_params(12, 9)
        cfg = XORConfig()
        # different seeds: rank 1's own keys are useless until rank 0's arrive
        w = EngineWrapper(cfg, _engine_kwargs=dict(_params=P, _backend=RefBackend(P, threads=2), seed=100 + rank),
                          rotation_steps=[])
        broadcast_evaluation_keys(w.ctx, src=0)
        w.public_key, w.relin_key, w.conj_key = w.ctx.public_key, w.ctx.relinearization_key, w.ctx.conjugation_key
        xs = XORService(w, CoefficientCache(cfg.coeffs_path))
        total = 4                                               # ciphertext batches in the job
        lo, hi = shard_range(total, rank, world)
        sc = w.engine.slot_count
        rng = np.random.default_rng(5)
        a = rng.integers(0, 16, (total, sc), dtype=np.uint8)
        b = rng.integers(0, 16, (total, sc), dtype=np.uint8)
        out = xs.xor_cipher_fused(w.encrypt(ZetaEncoder.to_zeta(a[lo:hi])), w.encrypt(ZetaEncoder.to_zeta(b[lo:hi])))
        parts = gather_handles(w.engine.backend, out.polys)
        if rank == 0:
            ok = True
            for r, polys in enumerate(parts):
                l, h = shard_range(total, r, world)
                dec = w.decrypt(Ciphertext(w.engine, polys, out.level))
                ok &= bool(np.array_equal(ZetaEncoder.from_zeta(np.atleast_2d(dec)), a[l:h] ^ b[l:h]))
            q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.slow
def test_two_rank_gloo_sharded_xor():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
