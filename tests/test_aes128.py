"""Chained rounds (BASELINE config 5) on the CPU oracle at a small ring: the refresh step
(bootstrap + clean-up polynomial) of aes_fhe_b200/services/aes128.py.  The ten-round run at
N = 2^16 is in test_gpu_aes.py."""
import numpy as np

from aes_fhe_b200.params import make_params
from aes_fhe_b200.services.aes128 import AES128Service
from aes_fhe_b200.services.key_expansion import expand_key
from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache, ZetaEncoder
from oracle import aes_plain as A

KEY_B = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
PT_B = bytes.fromhex("3243f6a8885a308d313198a2e0370734")


def make_service(backend, P, seed=2):
    cfg = XORConfig()
    w = EngineWrapper(cfg, _engine_kwargs=dict(_params=P, _backend=backend, seed=seed), rotation_steps=[])
    return w, AES128Service(w, XORService(w, CoefficientCache(cfg.coeffs_path)))


def plain_rounds(blocks, key, rounds):
    rks = expand_key(key)
    s = np.asarray(blocks, dtype=np.uint8) ^ rks[0]
    for r in range(1, rounds + 1):
        s = A.round_fn(s, rks[r], last=(r == 10))
    return s


def test_clean_up_polynomial_contracts_errors(ref_backend_cls):
    """(17 x - x^17) / 16 maps zeta (1 + e) to zeta (1 - 8.5 e^2): five levels, error squared."""
    P = make_params(11, 8, scale_bits=44)
    w, svc = make_service(ref_backend_cls(P), P)
    rng = np.random.default_rng(0)
    k = rng.integers(0, 16, w.engine.slot_count)
    noise = (rng.standard_normal(k.size) + 1j * rng.standard_normal(k.size)) * 2e-3
    ct = w.encrypt(ZetaEncoder.to_zeta(k) + noise)
    out = svc.clean(ct)
    assert out.level == ct.level - AES128Service.CLEAN_LEVELS
    err_in, err_out = np.abs(noise).max(), np.abs(w.decrypt(out) - ZetaEncoder.to_zeta(k)).max()
    assert err_in > 5e-3 and err_out < 8.5 * err_in ** 2 * 1.2 + 1e-6


def test_two_rounds_with_refresh_on_oracle(ref_backend_cls):
    """AddRoundKey_0 + two full rounds through `encrypt_blocks`: round 2 can only run because the
    state is bootstrapped and cleaned between the LUT layers (four bootstrap calls, 14 refreshed
    ciphertexts); every byte must equal plain AES, block 0 is FIPS-197 Appendix B."""
    P = make_params(11, 30, scale_bits=44)
    w, svc = make_service(ref_backend_cls(P), P)
    rng = np.random.default_rng(7)
    blocks = rng.integers(0, 256, (svc.B, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(PT_B, np.uint8)
    out = svc.encrypt_blocks(svc.encrypt_state(blocks), KEY_B, rounds=2)
    got = svc.decrypt_state(out)
    assert np.array_equal(got, plain_rounds(blocks, KEY_B, 2))
    assert got[0].tobytes().hex() == "aa8f5f0361dde3ef82d24ad26832469a"      # FIPS-197 App. B, start of round 3
    assert w.engine.op_counts["bootstrap"] == 4 and svc.refreshes == 14
    # slots are back on the unit circle: the clean-up leaves ~1e-7, not the bootstrap's 1e-3
    assert np.abs(np.abs(w.decrypt(out[0])) - 1).max() < 1e-4
