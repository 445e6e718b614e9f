"""Homomorphic AES round on the B200 at the BASELINE size (N = 2^16, 2048 blocks per
ciphertext, batch of ciphertexts), decoded bytes bit-exact against FIPS-197 and plain AES."""
import numpy as np
import pytest
import torch

from aes_fhe_b200.params import make_params
from test_aes_round import run_round_scenario

pytestmark = pytest.mark.gpu


def _gpu(P):
    from aes_fhe_b200.backend_cuda import CudaBackend
    return CudaBackend(P)


def test_round_one_fips197_full_size_with_stages(cuda_lib):
    P = make_params(16, 30)
    run_round_scenario(_gpu(P), P, batch=1, check_stages=True)


def test_round_one_fips197_full_size_batch(cuda_lib):
    P = make_params(16, 30)
    w, svc, r1 = run_round_scenario(_gpu(P), P, batch=2, check_stages=False)
    assert r1[0].batch == 2 and r1[0].level >= 4


def test_sbox_reference_order_and_bsgs_full_size(cuda_lib):
    from aes_fhe_b200.services.engine_context import EngineContext
    from aes_fhe_b200.services.sbox_service import SBoxService, AES_SBOX
    from aes_fhe_b200.services.xor_service import ZetaEncoder
    P = make_params(16, 22)
    ctx = EngineContext(signature=2, max_level=22, _engine_kwargs=dict(_params=P, seed=1), rotation_steps=[])
    svc = SBoxService(ctx)
    sc = ctx.engine.slot_count
    x = np.tile(np.arange(256, dtype=np.uint8), sc // 256 + 1)[:sc]          # test_sbox_array_simd
    ct = ctx.encrypt(ZetaEncoder.to_zeta(x, 256))
    exp = np.array(AES_SBOX, dtype=np.uint8)[x]
    for fn in (svc.sub_bytes_array, svc.sub_bytes_array_bsgs):
        dec = ctx.decrypt(fn(ct))
        assert np.abs(dec - ZetaEncoder.to_zeta(exp, 256)).max() < 1e-3
        assert np.array_equal(ZetaEncoder.from_zeta(dec, 256), exp)
