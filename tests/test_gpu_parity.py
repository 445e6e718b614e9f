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
    P = make_params(16, 30)          # N = 2^16, 31 + 6 limbs: the BASELINE configuration
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
