"""Pin the CPU oracle + facade against the reference's own golden data: the
slot / decoded-integer level is the only level the reference observes."""
import re
from pathlib import Path

import numpy as np
import pytest

from aes_fhe_b200.params import make_params
from aes_fhe_b200.services import lut
from aes_fhe_b200.services.engine_context import EngineContext
from aes_fhe_b200.services.sbox_service import SBoxService, AES_SBOX
from aes_fhe_b200.services.xor_service import (XORService, EngineWrapper, XORConfig, CoefficientCache, ZetaEncoder)
from conftest import REFERENCE

GOLD = Path(__file__).resolve().parent / "golden"


def _wrap(P, ref_backend_cls, steps=()):
    cfg = XORConfig()
    w = EngineWrapper(cfg, _engine_kwargs=dict(_params=P, _backend=ref_backend_cls(P), seed=11), rotation_steps=list(steps))
    return w, XORService(w, CoefficientCache(cfg.coeffs_path))


def test_coefficients_equal_committed_golden_vectors():
    """tests/golden/*.json were generated from the reference's JSONs by tests/golden/make_golden.py"""
    g = __import__("json").loads((GOLD / "coeff_digest.json").read_text())
    xor = lut.xor4_coeffs()
    assert np.count_nonzero(np.abs(xor) > 1e-12) == g["xor_nonzero"] == 64
    for i, j, re_, im_ in g["xor_entries"]:
        assert abs(xor[i, j] - complex(re_, im_)) < 1e-14
    hi, lo = lut.sbox_hi_lo_coeffs()
    assert np.allclose(hi[g["sbox_probe_idx"]], np.array(g["sbox_hi_probe_re"]) + 1j * np.array(g["sbox_hi_probe_im"]), atol=1e-14)
    assert np.allclose(lo[g["sbox_probe_idx"]], np.array(g["sbox_lo_probe_re"]) + 1j * np.array(g["sbox_lo_probe_im"]), atol=1e-14)
    assert abs(np.abs(hi).sum() - g["sbox_hi_l1"]) < 1e-9 and abs(np.abs(lo).sum() - g["sbox_lo_l1"]) < 1e-9
    assert [int(v) for v in lut.AES_SBOX] == g["aes_sbox"]


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree not mounted")
def test_coefficients_equal_reference_jsons():
    ref = lut.load_entries(REFERENCE / "xor_mono_coeffs.json")
    mine = lut.xor4_coeffs()
    assert set(ref) == {(i, j) for i in range(16) for j in range(16) if abs(mine[i, j]) > 1e-12}
    assert max(abs(ref[k] - mine[k]) for k in ref) < 1e-15
    hi, lo = lut.sbox_hi_lo_coeffs()
    assert np.abs(lut.load_json_coeffs(REFERENCE / "sbox/coeffs/sbox_hi_coeffs.json") - hi).max() < 1e-14
    assert np.abs(lut.load_json_coeffs(REFERENCE / "sbox/coeffs/sbox_lo_coeffs.json") - lo).max() < 1e-14
    src = (REFERENCE / "sbox/sbox_service.py").read_text()
    tab = [int(x, 16) for x in re.findall(r"0x[0-9a-f]{2}", src[src.index("AES_SBOX = ["):src.index("]", src.index("AES_SBOX = ["))])]
    assert tab == AES_SBOX


def test_engine_semantics_pinned_by_test_engine_rot(ref_backend_cls):
    """/root/reference/test/test_engine_rot.py:21-62 on the oracle engine."""
    P = make_params(12, 9)
    w, _ = _wrap(P, ref_backend_cls, steps=[1, 4])
    sc = w.engine.slot_count
    vec = np.linspace(0.0, 1.0, num=sc)
    assert np.allclose(w.decrypt(w.encrypt(vec)), vec, atol=1e-6)
    base = np.arange(sc, dtype=np.float64) / sc
    assert np.allclose(w.decrypt(w.rotate(w.encrypt(base), 5)), np.roll(base, 5), atol=1e-6)
    v = np.random.RandomState(0).rand(sc)
    ct = w.encrypt(v)
    assert np.allclose(w.decrypt(w.relinearize(ct)), v, atol=1e-6)          # no-op through the wrapper
    with pytest.raises(RuntimeError, match="should have 3 polynomials"):
        w.engine.relinearize(ct, w.relin_key)
    v = np.random.RandomState(1).rand(sc)
    ct = w.encrypt(v)
    sq = w.relinearize(w.multiply(ct, ct, w.relin_key))
    assert np.allclose(w.decrypt(sq), v * v, atol=1e-5)
    # short inputs are zero padded (test_xor_service.py:32-34)
    short = w.decrypt(w.encrypt(np.array([1.0, 2.0, 3.0])))
    assert np.allclose(short[:3], [1, 2, 3], atol=1e-6) and np.allclose(short[3:], 0, atol=1e-6)


def test_xor_all_256_nibble_pairs_reference_order(ref_backend_cls):
    """test_nibble_xor_bruteforce (test_xor_service.py:106-123) in one SIMD ciphertext pair."""
    P = make_params(12, 9)
    w, xs = _wrap(P, ref_backend_cls)
    a = np.repeat(np.arange(16, dtype=np.uint8), 16)
    b = np.tile(np.arange(16, dtype=np.uint8), 16)
    out = xs.xor(a, b)[:256]
    assert np.array_equal(out, a ^ b)
    c = w.engine.op_counts
    assert c["keyswitch_relin"] == 78 and c["keyswitch_galois"] == 14        # SURVEY 3.2 op counts


def test_xor_random_and_fused_schedule(ref_backend_cls):
    P = make_params(12, 9)
    w, xs = _wrap(P, ref_backend_cls)
    rng = np.random.default_rng(0)
    sc = w.engine.slot_count
    a = rng.integers(0, 16, size=sc, dtype=np.uint8)
    b = rng.integers(0, 16, size=sc, dtype=np.uint8)
    ea, eb = w.encrypt(ZetaEncoder.to_zeta(a)), w.encrypt(ZetaEncoder.to_zeta(b))
    r1, r2 = xs.xor_cipher(ea, eb), xs.xor_cipher_fused(ea, eb)
    assert r1.level == r2.level == 9 - 5
    z = ZetaEncoder.to_zeta(a ^ b)
    for r in (r1, r2):
        d = w.decrypt(r)
        assert np.abs(d - z).max() < 1e-3                                    # north_star slot bound
        assert np.array_equal(ZetaEncoder.from_zeta(d), a ^ b)


def test_sbox_all_256_bytes_both_schedules(ref_backend_cls):
    """test_sbox_array_simd (test_sbox_service.py:55-66) on a small ring."""
    P = make_params(12, 12)
    ctx = EngineContext(signature=2, max_level=12, _engine_kwargs=dict(_params=P, _backend=ref_backend_cls(P)),
                        rotation_steps=[])
    svc = SBoxService(ctx)
    sc = ctx.engine.slot_count
    x = np.tile(np.arange(256, dtype=np.uint8), sc // 256 + 1)[:sc]
    ct = ctx.encrypt(ZetaEncoder.to_zeta(x, 256))
    exp = np.array(AES_SBOX, dtype=np.uint8)[x]
    for fn, lvl, ks, cj in ((svc.sub_bytes_array, 2, 255, 0), (svc.sub_bytes_array_bsgs_hilo, 1, 32, 0),
                            (svc.sub_bytes_array_bsgs, 2, 22, 1)):
        ctx.engine.op_counts.clear()
        out = fn(ct)
        assert out.level == lvl and ctx.engine.op_counts["keyswitch_relin"] == ks
        assert ctx.engine.op_counts.get("keyswitch_galois", 0) == cj
        dec = ctx.decrypt(out)
        assert np.abs(dec - ZetaEncoder.to_zeta(exp, 256)).max() < 1e-3
        assert np.array_equal(ZetaEncoder.from_zeta(dec, 256), exp)


def test_nibble_pair_add_round_key_matches_reference_inputs(ref_backend_cls):
    """AESFHERound.full_round's scenario (new.py:186-227) with test_all_process.py:12-17 seeds."""
    from aes_fhe_b200.services.new import AESFHERound
    P = make_params(12, 9)
    w, xs = _wrap(P, ref_backend_cls)
    np.random.seed(25073101)
    state = np.random.randint(0, 256, 16, dtype=np.uint8)
    np.random.seed(25073102)
    key = np.random.randint(0, 256, 16, dtype=np.uint8)
    assert state.tobytes().hex() == "a5ca0aac94d89019f5f9dcb343297b67"
    assert key.tobytes().hex() == "1a7fe95f3c417065d0be3f41059adf59"
    out = AESFHERound(w, xs).full_round(state, key, recombine=True)
    assert out.tobytes().hex() == "bfb5e3f3a899e07c2547e3f246b3a43e"
