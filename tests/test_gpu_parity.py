"""Parity tests proper: the sm_100a build on a real B200, through the C ABI, against the CPU
oracle on the same seeded inputs -- bit-exact residues, then slot / byte level checks."""
import numpy as np
import pytest
import torch

import kernel_parity as kp
from aes_fhe_b200.params import make_params
from conftest import make_engines

pytestmark = pytest.mark.gpu


def _gpu(P):
    from aes_fhe_b200.backend_cuda import CudaBackend
    assert torch.cuda.is_available(), "these tests need the B200"
    return CudaBackend(P)


@pytest.mark.parametrize("log_n,lvl", [(12, 5), (13, 4), (14, 3), (15, 3), (16, 4)])
def test_primitives_small_chain(log_n, lvl, ref_backend_cls, cuda_lib):
    P = make_params(log_n, lvl)
    kp.check_primitives(P, _gpu(P), ref_backend_cls(P))


def test_primitives_full_parameters(ref_backend_cls, cuda_lib):
    P = make_params(16, 30)          # N = 2^16, 31 + 10 limbs, 3 digits of 11: the default Engine() chain
    kp.check_primitives(P, _gpu(P), ref_backend_cls(P))


def test_rescale(ref_backend_cls, cuda_lib):
    for P in (make_params(12, 6), make_params(16, 30)):
        kp.check_rescale(P, _gpu(P), ref_backend_cls(P))


@pytest.mark.parametrize("log_n,lvl,dnum", [(12, 6, 4), (12, 7, 3), (13, 4, 2), (14, 9, 1)])
def test_keyswitch_phases_small(log_n, lvl, dnum, ref_backend_cls, cuda_lib):
    P = make_params(log_n, lvl, dnum=dnum)
    kp.check_keyswitch(P, _gpu(P), ref_backend_cls(P))


def test_keyswitch_full_parameters(ref_backend_cls, cuda_lib):
    P = make_params(16, 30)
    kp.check_keyswitch(P, _gpu(P), ref_backend_cls(P), levels=[31, 30, 25, 17, 16, 9, 8, 2, 1])


def _benchmarked_sets():
    """the parameter sets bench.py times at N = 2^16: SubBytes (configs[1]: 23 + 11 limbs, two digits of 12 ->
    k_bconv<12>, k_ks_inner<2,...>) and bit-sliced AES-128 (configs[4]: 27 + 7 limbs, four digits of 7, 44-bit scale,
    sized against the sparse-secret bound)"""
    from aes_fhe_b200.params import LOG_PQ_BUDGET_SPARSE
    return {"subbytes": make_params(16, 22),
            "aes128": make_params(16, 26, scale_bits=44, log_pq_budget=LOG_PQ_BUDGET_SPARSE)}


@pytest.mark.parametrize("name", ["subbytes", "aes128"])
def test_primitives_at_benchmarked_parameters(name, ref_backend_cls, cuda_lib):
    P = _benchmarked_sets()[name]
    assert (P.n_q, P.n_p, P.alpha, P.dnum) == {"subbytes": (23, 11, 12, 2), "aes128": (27, 7, 7, 4)}[name]
    kp.check_primitives(P, _gpu(P), ref_backend_cls(P))
    kp.check_rescale(P, _gpu(P), ref_backend_cls(P))


@pytest.mark.parametrize("name", ["subbytes", "aes128"])
def test_keyswitch_phases_at_benchmarked_parameters(name, ref_backend_cls, cuda_lib):
    """residues of ModUp / key inner product / ModDown and of the whole key switch, bit for bit, at the top level,
    at digit boundaries and at the bottom of the chain"""
    P = _benchmarked_sets()[name]
    a = P.alpha
    levels = sorted({P.n_q, P.n_q - 1, 2 * a + 1, 2 * a, a + 1, a, 2, 1} & set(range(1, P.n_q + 1)), reverse=True)
    kp.check_keyswitch(P, _gpu(P), ref_backend_cls(P), levels=levels)


@pytest.mark.parametrize("log_n,lvl", [(12, 6), (14, 6)])
def test_engine_ops(log_n, lvl, ref_backend_cls, cuda_lib):
    P = make_params(log_n, lvl)
    eg, er = make_engines(P, ref_backend_cls, _gpu(P))
    kp.check_engine_ops(eg, er)


def test_engine_ops_full_ring(ref_backend_cls, cuda_lib):
    P = make_params(16, 8)
    eg, er = make_engines(P, ref_backend_cls, _gpu(P))
    kp.check_engine_ops(eg, er)


def test_linearity_property_full_size(cuda_lib):
    """Size-independent property at the BASELINE size: NTT(a) + NTT(b) == NTT(a + b) and the
    transform inverts, on 2 x 37 limbs of N = 2^16."""
    P = make_params(16, 30)
    gb = _gpu(P)
    rng = np.random.default_rng(5)
    from conftest import rand_poly
    a = gb.from_numpy(rand_poly(P, rng, 2, P.n_q, True, batch=2))
    b = gb.from_numpy(rand_poly(P, rng, 2, P.n_q, True, batch=2))
    nq, K = P.n_q, P.n_p
    lhs = gb.add(gb.ntt(a, nq, K), gb.ntt(b, nq, K), nq, K)
    rhs = gb.ntt(gb.add(a, b, nq, K), nq, K)
    assert torch.equal(lhs, rhs)
    assert torch.equal(gb.intt(gb.ntt(a, nq, K), nq, K), a)
    # negacyclic convolution theorem against the oracle-free schoolbook on a sparse operand
    x = np.zeros(P.n, dtype=np.int64); x[1] = 1                      # X
    y = rng.integers(-5, 6, P.n).astype(np.int64)
    px, py = gb.from_i64(x, nq, False), gb.from_i64(y, nq, False)
    prod = gb.intt(gb.mul(px, py, nq, 0), nq, 0)
    shifted = np.roll(y, 1); shifted[0] = -shifted[0]                # X * y  (mod X^N + 1)
    assert torch.equal(prod, gb.intt(gb.from_i64(shifted, nq, False), nq, 0))


@pytest.mark.parametrize("log_n,batch", [(12, 2), (14, 3)])
def test_fused_lut_services(log_n, batch, ref_backend_cls, cuda_lib):
    P = make_params(log_n, 13)
    kp.check_fused_services(P, _gpu(P), ref_backend_cls(P), batch=batch)


def test_device_codec_roundtrip_and_compat(cuda_lib):
    """GPU-side encode / sampling / decode (throughput path) agrees with the host codec."""
    from aes_fhe_b200.engine import Engine
    P = make_params(16, 6)
    eng = Engine(_params=P, seed=9)
    sk = eng.create_secret_key(); pk = eng.create_public_key(sk); rlk = eng.create_relinearization_key(sk)
    rng = np.random.default_rng(0)
    v = np.exp(-2j * np.pi * rng.integers(0, 256, (3, eng.slot_count)) / 256)
    eng.device_codec = True
    ct = eng.encrypt(v, pk)
    d_dev = eng.decrypt(ct, sk)
    eng.device_codec = False
    d_host = eng.decrypt(ct, sk)
    # fresh public-key encryption noise at N = 2^16: ~2e-7 rms per slot, ~1.2e-6 max over 10^5 slots
    assert np.abs(d_dev - v).max() < 5e-6 and np.abs(d_host - v).max() < 5e-6
    assert np.abs(d_dev - d_host).max() < 1e-9
    sq = eng.multiply(ct, ct, rlk)
    eng.device_codec = True
    assert np.abs(eng.decrypt(sq, sk) - v * v).max() < 1e-5
    # byte-level codec on the device: zeta_256 encode + encrypt, decrypt + decode, ragged input zero-padded
    x = rng.integers(0, 256, (3, eng.slot_count), dtype=np.uint8)
    assert np.array_equal(eng.decrypt_zeta(eng.encrypt_zeta(x, pk, 256), sk, 256), x)
    assert np.abs(eng.decrypt(eng.encrypt_zeta(x[0, :100], pk, 16), sk)[:100] - np.exp(-2j * np.pi * (x[0, :100] % 16) / 16)).max() < 5e-6
    assert np.abs(eng.decrypt(eng.encrypt_zeta(x[0, :100], pk, 16), sk)[100:]).max() < 5e-6
    short = eng.decrypt(eng.encrypt(np.array([1.0, 2.0, 3.0]), pk), sk)
    assert np.allclose(short[:3], [1, 2, 3], atol=5e-6) and np.allclose(short[3:], 0, atol=5e-6)


def test_fused_ntt_equals_two_pass_at_scale(ref_backend_cls, cuda_lib):
    """The single-launch NTT (csrc/ntt_fused.cuh: persistent groups, L2-resident scratch, group
    barriers) against the two-pass kernels on far more rows than there are groups, including the
    fused load/store functors (rescale, key switch, mod-raise); no barrier may have timed out."""
    from conftest import rand_poly
    P = make_params(16, 30)
    gb = _gpu(P)
    rng = np.random.default_rng(11)
    nq, K = P.n_q, P.n_p
    a = gb.from_numpy(rand_poly(P, rng, 2, nq, True, batch=12))            # 24 x 39 = 936 rows
    d = gb.from_numpy(rand_poly(P, rng, 1, 25, False, batch=5))
    ksk = gb.from_numpy(np.stack([rand_poly(P, rng, 2, nq, True) for _ in range(P.dnum)]))
    r = gb.from_numpy(rand_poly(P, rng, 3, 17, False, batch=3))
    outs = []
    for fused in (1, 2, 0):        # persistent single-launch, chained single-launch, two-pass
        assert gb.lib.fhe_set_ntt_fused(gb.ctx, fused) == 0
        f = gb.ntt(a, nq, K)
        outs.append([f, gb.intt(a, nq, K), gb.intt(f, nq, K), gb.keyswitch(d, ksk, 25), gb.rescale(r, 17),
                     gb.mod_raise(r[:, :, :1].contiguous(), 9)])
    assert gb.lib.fhe_ntt_fused_status(gb.ctx) == 0
    for x, y, z in zip(*outs):
        assert torch.equal(x, z) and torch.equal(y, z)
    assert torch.equal(outs[0][2], a)
    # one row against the oracle
    one = rand_poly(P, rng, 1, nq, True)
    assert gb.lib.fhe_set_ntt_fused(gb.ctx, 1) == 0
    assert np.array_equal(gb.to_numpy(gb.ntt(gb.from_numpy(one), nq, K)), ref_backend_cls(P).ntt(one, nq, K))


def test_bit_sliced_round_residues_equal_oracle(ref_backend_cls, cuda_lib):
    """ARK_0 + ShiftRows + SubBytes + MixColumns/ARK of the bit-sliced AES path (services/aes_bits.py) on the
    B200 and on the oracle from the same seed: the 32 output ciphertexts must agree residue for residue, and
    decode to plain AES."""
    from test_aes_bits import make_service, KEY_B
    from aes_fhe_b200.services.key_expansion import expand_key
    from oracle import aes_plain as A
    P = make_params(13, 9, scale_bits=44)
    res = []
    for be in (_gpu(P), ref_backend_cls(P)):
        w, svc = make_service(be, P, boot_key=object())
        rng = np.random.default_rng(3)
        blocks = rng.integers(0, 256, (svc.Bs, 16), dtype=np.uint8)
        rks = expand_key(KEY_B)
        k0, k1 = svc.encrypt_round_key(rks[0]), svc.encrypt_round_key(rks[1])
        st = svc.add_round_key(svc.encrypt_state(blocks), k0)
        st = svc.mix_columns_ark(svc.sub_bytes(svc.shift_rows(st)), k1)
        assert np.array_equal(svc.decrypt_state(st), A.round_fn(blocks ^ rks[0], rks[1]))
        res.append(be.to_numpy(st.polys))
    assert np.array_equal(res[0], res[1])


def test_multiply_gather_on_gpu(ref_backend_cls, cuda_lib):
    """fhe_mul_relin_rescale_ptrs on the B200 against the oracle, residue for residue, at a small ring and at N = 2^16"""
    for P in (make_params(13, 7, dnum=4), make_params(16, 5)):
        kp.check_multiply_gather(P, _gpu(P), ref_backend_cls(P))


def test_bit_bootstrap_residues_equal_oracle(ref_backend_cls, cuda_lib):
    """the whole bit bootstrap (SlotToCoeff, ModRaise, CoeffToSlot with hoisted rotations and shared ModDowns,
    EvalMod) on the B200 against the oracle, residue for residue, at N = 2^12"""
    from test_aes_bits import make_service
    P = make_params(12, 18, scale_bits=44)
    res, rows = [], []
    for be in (_gpu(P), ref_backend_cls(P)):
        w, svc = make_service(be, P)
        rng = np.random.default_rng(5)
        u = rng.choice([-1.0, 1.0], (2, w.engine.slot_count)) + 1j * rng.choice([-1.0, 1.0], (2, w.engine.slot_count))
        ct = w.engine.encrypt(u, w.public_key, level=svc.boot_in_levels + 1)
        svc.prepare_keys()
        r0 = be.ntt_row_count()
        out = w.engine.bootstrap_bits(ct, w.relin_key, w.conj_key, svc.boot_key)
        rows.append(be.ntt_row_count() - r0)
        got = w.engine.decrypt(out, w.secret_key)
        assert np.abs(got - np.concatenate([u.real, u.imag])).max() < 1e-4
        res.append(be.to_numpy(out.polys))
    assert np.array_equal(res[0], res[1])
    # the work unit of bench.py's CPU extrapolation: library and oracle do the same number of length-N transforms
    # for the same schedule (plaintext encodings are cached on both sides by now, keys exist)
    assert rows[0] == rows[1], rows


def test_byte_nibble_bridge_gf_service_on_gpu(ref_backend_cls, cuda_lib):
    """rows a8 / a10 / f-2 on the B200 (the CPU-oracle versions are in test_services_plain_and_oracle.py):
    extract_nibbles / recombine_nibbles / byte-domain add_round_key (xor_service.py:256-269, 434-547) and
    GFService.mul2 / mul3 (gf_service.py:55-78): decoded values exact, slots within 1e-3 of the zeta encoding."""
    from aes_fhe_b200.services.gf_service import GFService, gf_tables
    from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache, ZetaEncoder
    P = make_params(13, 24)
    cfg = XORConfig()
    w = EngineWrapper(cfg, _engine_kwargs=dict(_params=P, _backend=_gpu(P), seed=21), rotation_steps=[])
    xs = XORService(w, CoefficientCache(cfg.coeffs_path))
    sc = w.engine.slot_count
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, sc, dtype=np.uint8)
    x[:256] = np.arange(256)
    ct = w.encrypt(ZetaEncoder.to_zeta(x, 256))
    hi, lo = xs.extract_nibbles(ct)
    assert np.abs(w.decrypt(hi) - ZetaEncoder.to_zeta(x >> 4)).max() < 1e-3
    assert np.array_equal(ZetaEncoder.from_zeta(w.decrypt(hi)), x >> 4)
    assert np.array_equal(ZetaEncoder.from_zeta(w.decrypt(lo)), x & 15)
    back = xs.recombine_nibbles(hi, lo)
    assert np.array_equal(ZetaEncoder.from_zeta(w.decrypt(back), 256), x)
    key = rng.integers(0, 256, sc, dtype=np.uint8)
    out = xs.add_round_key(ct, key)
    assert np.array_equal(ZetaEncoder.from_zeta(w.decrypt(out), 256), x ^ key)
    gf = GFService(w, xs)
    t2, t3 = gf_tables()
    for fn, tab in ((gf.mul2_bsgs, t2), (gf.mul3_bsgs, t3)):
        h2, l2 = fn(ct)
        assert np.array_equal(ZetaEncoder.from_zeta(w.decrypt(h2)), tab[x] >> 4)
        assert np.array_equal(ZetaEncoder.from_zeta(w.decrypt(l2), 256), tab[x] & 15)


@pytest.mark.parametrize("log_n,lvl,half_width,batch", [(13, 8, 31, 5), (16, 6, 31, 2)])
def test_double_hoisted_transform_on_gpu(log_n, lvl, half_width, batch, ref_backend_cls, cuda_lib):
    """fhe_bsgs_inner on the B200 against the oracle's primitives, residue for residue (62 diagonals: 16 baby x 4
    giant steps, the shape of a CoeffToSlot factor at N = 2^16)"""
    P = make_params(log_n, lvl)
    kp.check_double_hoisted_transform(P, _gpu(P), ref_backend_cls(P), batch=batch, half_width=half_width)
