"""FIPS-197 bit-exactness of the homomorphic round (north_star check 3), CPU oracle on a small
ring; the same scenario runs on the B200 at N = 2^16 in test_gpu_aes.py."""
import numpy as np
import pytest

from aes_fhe_b200.params import make_params
from aes_fhe_b200.services.aes_round import AESRoundService
from aes_fhe_b200.services.key_expansion import expand_key
from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache
from oracle import aes_plain as A

KEY_B = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
PT_B = bytes.fromhex("3243f6a8885a308d313198a2e0370734")


def test_plain_aes_oracle_matches_fips197():
    assert A.encrypt_blocks(np.frombuffer(PT_B, np.uint8)[None], KEY_B)[0].tobytes().hex() == "3925841d02dc09fbdc118597196a0b32"
    k = bytes(range(16))
    p = np.frombuffer(bytes.fromhex("00112233445566778899aabbccddeeff"), np.uint8)
    assert A.encrypt_blocks(p[None], k)[0].tobytes().hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"
    assert np.array_equal(expand_key(KEY_B), A.key_schedule(KEY_B))
    assert expand_key(KEY_B)[1].tobytes().hex() == "a0fafe1788542cb123a339392a6c7605"


def run_round_scenario(backend, P, batch=1, check_stages=True):
    cfg = XORConfig()
    w = EngineWrapper(cfg, _engine_kwargs=dict(_params=P, _backend=backend, seed=2), rotation_steps=[])
    svc = AESRoundService(w, XORService(w, CoefficientCache(cfg.coeffs_path)))
    rks = expand_key(KEY_B)
    rng = np.random.default_rng(7)
    blocks = [rng.integers(0, 256, (svc.B, 16), dtype=np.uint8) for _ in range(batch)]
    blocks[0][0] = np.frombuffer(PT_B, np.uint8)
    plain = np.stack(blocks) if batch > 1 else blocks[0]
    st = svc.encrypt_state(blocks if batch > 1 else blocks[0])
    k0, k1 = svc.encrypt_round_key(rks[0]), svc.encrypt_round_key(rks[1])
    s0 = svc.add_round_key(st, k0)
    if check_stages:
        d0 = svc.decrypt_state(s0)
        assert np.array_equal(d0, plain ^ rks[0])
        assert d0.reshape(-1, 16)[0].tobytes().hex() == "193de3bea0f4e22b9ac68d2ae9f84808"
        pl = svc.sbox_planes(s0)
        sb = svc.decrypt_state((pl["S_hi"], pl["S_lo"]))
        assert sb.reshape(-1, 16)[0].tobytes().hex() == "d42711aee0bf98f1b8b45de51e415230"
        assert np.array_equal(sb, A.sub_bytes(plain ^ rks[0]))
        sr = svc.decrypt_state(svc.shift_rows((pl["S_hi"], pl["S_lo"])))
        assert sr.reshape(-1, 16)[0].tobytes().hex() == "d4bf5d30e0b452aeb84111f11e2798e5"
        mc = svc.shift_rows_mix_columns(pl)
        dm = svc.decrypt_state(mc)
        assert dm.reshape(-1, 16)[0].tobytes().hex() == "046681e5e0cb199a48f8d37a2806264c"
        r1 = svc.add_round_key(mc, k1)
    else:
        r1 = svc.round(s0, k1)
    d1 = svc.decrypt_state(r1)
    assert d1.reshape(-1, 16)[0].tobytes().hex() == "a49c7ff2689f352b6b5bea43026a5049"    # FIPS-197 App. B, round 2 input
    assert np.array_equal(d1, A.round_fn(plain ^ rks[0], rks[1]))
    # last-round form (no MixColumns) on the same state
    return w, svc, r1


@pytest.mark.slow
def test_round_one_matches_fips197_appendix_b(ref_backend_cls):
    P = make_params(12, 27)
    run_round_scenario(ref_backend_cls(P), P)
