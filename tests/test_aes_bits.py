"""Bit-sliced AES-128 (aes_fhe_b200/services/aes_bits.py) on the CPU oracle at small rings: every stage
against the plain +-1 model and plain AES, the bit bootstrap, and chained rounds.  The full-size run
(N = 2^16, ten rounds) is in test_gpu_aes.py."""
import numpy as np
import pytest

from aes_fhe_b200 import bootstrap as B
from aes_fhe_b200.params import make_params
from aes_fhe_b200.services import aes_bits as AB
from aes_fhe_b200.services.key_expansion import expand_key
from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
from oracle import aes_plain as A

KEY_B = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
PT_B = bytes.fromhex("3243f6a8885a308d313198a2e0370734")


def make_service(backend, P, seed=2, **kw):
    w = EngineWrapper(XORConfig(), _engine_kwargs=dict(_params=P, _backend=backend, seed=seed), rotation_steps=[])
    return w, AB.AESBitService(w, **kw)


def test_walsh_form_of_the_sbox_is_exact_and_multilinear():
    W = AB.sbox_walsh()
    x = np.arange(256)
    s = AB.bits_pm(x)
    out = np.einsum("kab,ax,bx->kx", W, AB.monomials(s[4:]), AB.monomials(s[:4]))
    assert np.abs(out - AB.bits_pm(A.SBOX[x])).max() < 1e-12
    assert np.allclose(np.abs(W).sum(axis=(1, 2)), 13.5)           # sum |w| of every output bit


def test_plain_bit_model_is_aes128():
    rng = np.random.default_rng(1)
    blocks = rng.integers(0, 256, (64, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(PT_B, np.uint8)
    rks = A.key_schedule(KEY_B)
    st = AB.PlainBits.xor(AB.PlainBits.from_blocks(blocks), AB.PlainBits.from_key(rks[0], 64))
    for r in range(1, 11):
        st = AB.PlainBits.sub_bytes(AB.PlainBits.shift_rows(st))
        st = AB.PlainBits.mix_ark(st, AB.PlainBits.from_key(rks[r], 64), last=(r == 10))
    got = AB.PlainBits.to_blocks(st)
    assert np.array_equal(got, A.encrypt_blocks(blocks, KEY_B))
    assert got[0].tobytes().hex() == "3925841d02dc09fbdc118597196a0b32"          # FIPS-197 Appendix B


def test_pack_unpack_round_trip_and_layout():
    class _E:                                                            # just enough engine for the packing
        slot_count = 512
    svc = AB.AESBitService.__new__(AB.AESBitService)
    svc.sc, svc.Bs = 512, 128
    rng = np.random.default_rng(0)
    blocks = rng.integers(0, 256, (2 * 128, 16), dtype=np.uint8)
    planes = svc.pack_bits(blocks)
    assert planes.shape == (64, 512)
    assert np.array_equal(svc.unpack_bits(planes), blocks)
    # bit k of FIPS byte 4 c + r of block b of state g: row (k * 4 + r) * G + g, slot c * Bs + b
    k, r, g, c, b = 5, 2, 1, 3, 17
    assert planes[(k * 4 + r) * 2 + g, c * 128 + b] == (blocks[g * 128 + b, 4 * c + r] >> k) & 1
    assert np.array_equal(svc.unpack_bits(svc.pack_bits(blocks[:100]), nb=100), blocks[:100])     # ragged


def test_round_stages_on_oracle(ref_backend_cls):
    """ARK_0, ShiftRows, SubBytes, MixColumns + ARK of round 1 on two states, each stage against plain AES and
    the slot error against the +-1 model; then the last-round form (no MixColumns)."""
    P = make_params(10, 9, scale_bits=44)
    w, svc = make_service(ref_backend_cls(P), P, boot_key=object())
    rng = np.random.default_rng(3)
    G = 2
    blocks = rng.integers(0, 256, (G * svc.Bs, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(PT_B, np.uint8)
    rks = expand_key(KEY_B)
    k0, k1 = svc.encrypt_round_key(rks[0], G), svc.encrypt_round_key(rks[1], G)
    st = svc.add_round_key(svc.encrypt_state(blocks), k0)
    assert st.level == 8 and np.array_equal(svc.decrypt_state(st), blocks ^ rks[0])
    st = svc.shift_rows(st)
    want = A.shift_rows(blocks ^ rks[0])
    assert st.level == 8 and np.array_equal(svc.decrypt_state(st), want)
    st = svc.sub_bytes(st)
    want = A.sub_bytes(want)
    assert st.level == 8 - svc.SBOX_LEVELS and np.array_equal(svc.decrypt_state(st), want)
    assert np.abs(svc.decrypt_slots(st) - (1.0 - 2.0 * svc.pack_bits(want))).max() < 1e-6
    last = svc.add_round_key(st, k1)
    assert np.array_equal(svc.decrypt_state(last), want ^ rks[1])
    st = svc.mix_columns_ark(st, k1)
    want = A.mix_columns(want) ^ rks[1]
    assert st.level == 8 - svc.SBOX_LEVELS - svc.MIX_LEVELS and np.array_equal(svc.decrypt_state(st), want)
    assert svc.decrypt_state(st)[0].tobytes().hex() == "a49c7ff2689f352b6b5bea43026a5049"       # FIPS-197 App. B, round 2 input
    assert np.abs(svc.decrypt_slots(st) - (1.0 - 2.0 * svc.pack_bits(want))).max() < 1e-5
    # 1 + 22 + 8 + 140 products per state, 3 batched rotations
    n = w.engine.op_counts
    assert n["keyswitch_galois"] == 3 and n["mul_ct"] == 1 + 3 + 1 + 3          # batched calls (pairs, triples, quadruples of both nibbles together)


def test_evalmod_design_for_bits_is_flat_at_the_bits():
    """sin(2 pi x) by the degree-18 interpolant + 5 double-angle steps: exact at x = I +- 1/4 to 1e-4, and an
    input error e comes out as pi^2 e^2 / 8"""
    poly, alphas = B._evalmod_design(2 * np.pi, "monomial", B.DOUBLE_ANGLES_BITS, B.POLY_DEGREE_BITS)
    rng = np.random.default_rng(0)
    I = rng.integers(-(B.K_NORM - 1), B.K_NORM, 40000)
    s = rng.choice([-1.0, 1.0], I.size)

    def f(e):
        c = np.polynomial.polynomial.polyval((I + (s + e) / 4) / B.K_NORM, poly)
        for i in range(B.DOUBLE_ANGLES_BITS):
            c = c * c - alphas[i + 1]
        return c

    assert np.abs(f(0.0) - s).max() < 1e-4
    e = rng.normal(0, 0.02, I.size)
    assert np.abs(f(e) - s * (1 - np.pi ** 2 * e ** 2 / 8)).max() < 2e-4


@pytest.mark.parametrize("log_n", [11])
def test_bit_bootstrap_on_oracle(log_n, ref_backend_cls):
    P = make_params(log_n, 20, scale_bits=44)
    w, svc = make_service(ref_backend_cls(P), P)
    eng = w.engine
    rng = np.random.default_rng(5)
    u = rng.choice([-1.0, 1.0], (2, eng.slot_count))
    v = rng.choice([-1.0, 1.0], (2, eng.slot_count))
    noise = rng.normal(0, 0.005, (2, 2, eng.slot_count))
    ct = eng.encrypt((u + noise[0]) + 1j * (v + noise[1]), w.public_key, level=svc.boot_in_levels + 1)
    out = eng.bootstrap_bits(ct, w.relin_key, w.conj_key, svc.boot_key)
    plan = svc.boot_key.plan
    assert out.batch == 4 and out.level == 20 - plan.depth_bits
    got = eng.decrypt(out, w.secret_key)
    err = np.abs(got - np.concatenate([u, v])).max()
    assert err < 1e-3, err                      # input error 5e-3 rms (2e-2 max) came out squared: pi^2 e^2 / 8
    # the same refresh raised only to level 17: three limbs fewer in every step, result at 17 - 13, same cleaning
    low = eng.bootstrap_bits(ct, w.relin_key, w.conj_key, svc.boot_key, top_level=17)
    assert low.level == 17 - plan.depth_bits
    assert np.abs(eng.decrypt(low, w.secret_key) - np.concatenate([u, v])).max() < 1e-3
    with pytest.raises(RuntimeError):
        eng.bootstrap_bits(ct, w.relin_key, w.conj_key, svc.boot_key, top_level=21)           # above max_level
    with pytest.raises(RuntimeError):
        eng.bootstrap_bits(ct, w.relin_key, w.conj_key, svc.boot_key, top_level=plan.depth_bits)     # no level left for the result


def test_final_round_key_squares_the_error(ref_backend_cls):
    """Last AddRoundKey with the half-amplitude key: (o k / 2)(3 - o^2) = s k (1 - 1.5 e^2) for o = s (1 + e); a
    full-amplitude key gives the plain product (error e kept)."""
    P = make_params(10, 4, scale_bits=44)
    w, svc = make_service(ref_backend_cls(P), P, boot_key=object())
    eng = w.engine
    rng = np.random.default_rng(13)
    blocks = rng.integers(0, 256, (svc.Bs, 16), dtype=np.uint8)
    rk = expand_key(KEY_B)[10]
    s = 1.0 - 2.0 * svc.pack_bits(blocks)
    e = rng.normal(0, 0.01, s.shape)
    o = eng.encrypt(s * (1.0 + e), w.public_key, level=2)
    want = 1.0 - 2.0 * svc.pack_bits(blocks ^ rk)
    half = svc.final_round_key(o, svc.encrypt_round_key(rk, level=2, half=True))
    assert half.level == 0
    err_half = np.abs(svc.decrypt_slots(half) - want * (1.0 + e) * (3.0 - (1.0 + e) ** 2) / 2.0).max()     # = 1 - 1.5 e^2 - e^3 / 2
    assert err_half < 1e-6, err_half
    assert np.abs(svc.decrypt_slots(half) - want).max() < 1.6 * np.abs(e).max() ** 2 + 2e-5
    full = svc.final_round_key(o, svc.encrypt_round_key(rk, level=2))
    assert full.level == 1 and np.abs(svc.decrypt_slots(full) - want * (1.0 + e)).max() < 1e-5
    assert svc.round_levels(True) == 6 and svc.round_levels(False) == 7


def test_two_rounds_with_refresh_on_oracle(ref_backend_cls):
    P = make_params(11, 24, scale_bits=44)          # 13 (bit bootstrap) + 7 (round) + 4 (entry of the next bootstrap)
    w, svc = make_service(ref_backend_cls(P), P)
    rng = np.random.default_rng(7)
    blocks = rng.integers(0, 256, (svc.Bs, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(PT_B, np.uint8)
    rks = expand_key(KEY_B)
    st = svc.encrypt_state(blocks, level=1 + svc.boot_in_levels)
    out = svc.encrypt_blocks(st, KEY_B, rounds=2)
    s = blocks ^ rks[0]
    for r in (1, 2):
        s = A.round_fn(s, rks[r])
    got = svc.decrypt_state(out)
    assert np.array_equal(got, s)
    assert got[0].tobytes().hex() == "aa8f5f0361dde3ef82d24ad26832469a"      # FIPS-197 App. B, start of round 3
    assert w.engine.op_counts["bootstrap"] == 2 and svc.refreshes == 32
    assert np.abs(svc.decrypt_slots(out) - (1.0 - 2.0 * svc.pack_bits(s))).max() < 1e-3


def test_block_io_helpers_follow_the_reference_interface():
    """utils mirror (row f-3): column-major state, PKCS#7 always pads, chunking"""
    from aes_fhe_b200.services import utils as U
    b = bytes(range(16))
    s = U.bytes_to_state(b)
    assert s[1, 0] == 1 and s[0, 1] == 4 and U.state_to_bytes(s) == b
    assert U.pkcs7_unpad(U.pkcs7_pad(b"abc")) == b"abc" and len(U.pkcs7_pad(b"a" * 16)) == 32
    assert U.chunk_bytes(b"x" * 40) == [b"x" * 16, b"x" * 16, b"x" * 8]
    with pytest.raises(ValueError):
        U.pkcs7_unpad(b"abc\x05")
    with pytest.raises(ValueError):
        U.bytes_to_state(b"short")
    assert np.array_equal(U.zeta_decode(U.zeta_encode(np.arange(256), 256), 256), np.arange(256))


def test_ctr_counter_blocks(monkeypatch):
    """encrypt_ctr builds nonce || big-endian counter blocks and XORs the keystream (checked with plain AES in
    place of the homomorphic evaluation; NIST SP 800-38A F.5.1 uses a 128-bit counter, so the vector here is
    plain AES of the same blocks)"""
    seen = {}

    def fake(svc, blocks, key16):
        seen["blocks"] = blocks.copy()
        return A.encrypt_blocks(blocks, key16)

    monkeypatch.setattr(AB, "_run_blocks", fake)
    nonce = bytes(range(100, 112))
    data = bytes(range(50))
    out = AB.encrypt_ctr(None, data, KEY_B, nonce, counter0=0xFFFFFFFE)
    assert seen["blocks"].shape == (4, 16) and seen["blocks"][0, :12].tobytes() == nonce
    assert [bytes(b[12:]).hex() for b in seen["blocks"]] == ["fffffffe", "ffffffff", "00000000", "00000001"]
    ks = A.encrypt_blocks(seen["blocks"], KEY_B).reshape(-1)[:50]
    assert out == (np.frombuffer(data, np.uint8) ^ ks).tobytes()
    assert AB.encrypt_ctr(None, out, KEY_B, nonce, counter0=0xFFFFFFFE) == data


def test_ecb_ten_rounds_on_oracle(ref_backend_cls):
    """bytes in, bytes out: AES-128-ECB of a 40-byte message (PKCS#7 -> 3 blocks) through ten homomorphic rounds
    with nine bit bootstraps on the oracle at N = 2^11; FIPS-197 Appendix C.1 is the first block"""
    P = make_params(11, 24, scale_bits=44)
    w, svc = make_service(ref_backend_cls(P), P)
    key = bytes(range(16))
    msg = bytes.fromhex("00112233445566778899aabbccddeeff") + b"twenty-four more bytes.."
    ct = AB.encrypt_ecb(svc, msg, key)
    assert len(ct) == 48 and ct[:16].hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"
    from aes_fhe_b200.services.utils import pkcs7_pad
    want = A.encrypt_blocks(np.frombuffer(pkcs7_pad(msg), np.uint8).reshape(-1, 16), key).tobytes()
    assert ct == want
    assert w.engine.op_counts["bootstrap"] == 10


def test_level_plan_picks_the_input_level_with_the_fewest_refreshes():
    """host logic of AESBitService.plan_levels / best_fresh_level on the 24-level chain: a full-height refresh leaves
    11 levels, a round takes 7 (the last 6, polish included) and must leave 4 for the next refresh; a refresh is
    raised only as high as the rounds up to the next one need"""
    class _Eng:
        max_level = 24

    class _Key:
        _groups = 3
    svc = AB.AESBitService.__new__(AB.AESBitService)
    svc.engine, svc.boot_key, svc.boot_in_levels = _Eng(), _Key(), 4
    low = svc.plan_levels(5)
    assert low["refresh_before_rounds"] == list(range(1, 11)) and low["key_levels"] == [5] + [7] * 9 + [2] and low["out_level"] == 0
    assert low["refresh_top_levels"] == {**{r: 24 for r in range(1, 10)}, 10: 19}      # the last one: 13 + 6, not 24
    best = svc.plan_levels(19)
    assert best["refresh_before_rounds"] == [3, 4, 5, 6, 7, 8, 9, 10] and best["key_levels"][:3] == [19, 14, 7]
    svc.engine.max_level = 26                                                                # the default chain
    deep = svc.plan_levels(26)
    assert deep["refresh_before_rounds"] == [4, 5, 6, 7, 8, 9] and deep["out_level"] == 0
    assert deep["refresh_top_levels"] == {4: 24, 5: 24, 6: 24, 7: 24, 8: 24, 9: 26} and deep["key_levels"][4:] == [7, 7, 7, 7, 7, 9, 2]
    svc.engine.max_level = 24
    assert svc.plan_levels(24)["refresh_before_rounds"] == best["refresh_before_rounds"]       # more levels buy nothing
    assert svc.best_fresh_level() == 19
    with pytest.raises(RuntimeError):
        svc.plan_levels(3)
