"""The CUDA kernel sources, compiled for the host simulator (csrc/compat.h, -DFHE_EMU), must
match the CPU oracle bit for bit through the C ABI.  Runs without a GPU; the same checks run
on the real sm_100a build in test_gpu_parity.py."""
import numpy as np
import pytest

import kernel_parity as kp
from aes_fhe_b200.backend_cuda import CudaBackend
from aes_fhe_b200.params import make_params
from conftest import make_engines


def _emu(P, emu_lib, fused=1):
    """NTT variant: 1 = persistent single-launch (csrc/ntt_fused.cuh, the simulator default), 0 = two-pass
    (csrc/ntt.cuh), 2 = chained single-launch (csrc/ntt_chained.cuh)."""
    b = CudaBackend(P, _lib_path=emu_lib, _device="cpu")
    assert b.lib.fhe_set_ntt_fused(b.ctx, fused) == 0
    return b


@pytest.mark.parametrize("fused", [1, 0, 2])
@pytest.mark.parametrize("log_n,lvl", [(12, 5), (13, 4), (14, 2)])
def test_primitives(log_n, lvl, fused, emu_lib, ref_backend_cls):
    P = make_params(log_n, lvl)
    b = _emu(P, emu_lib, fused)
    kp.check_primitives(P, b, ref_backend_cls(P))
    assert b.lib.fhe_ntt_fused_status(b.ctx) == 0


def test_primitives_full_ring(emu_lib, ref_backend_cls):
    P = make_params(16, 2)
    kp.check_primitives(P, _emu(P, emu_lib), ref_backend_cls(P))


@pytest.mark.parametrize("fused", [1, 0, 2])
def test_rescale(fused, emu_lib, ref_backend_cls):
    P = make_params(12, 6)
    kp.check_rescale(P, _emu(P, emu_lib, fused), ref_backend_cls(P))


@pytest.mark.parametrize("fused", [1, 0, 2])
@pytest.mark.parametrize("log_n,lvl,dnum", [(12, 6, 4), (12, 7, 3), (13, 4, 2)])
def test_keyswitch_phases(log_n, lvl, dnum, fused, emu_lib, ref_backend_cls):
    P = make_params(log_n, lvl, dnum=dnum)
    kp.check_keyswitch(P, _emu(P, emu_lib, fused), ref_backend_cls(P))


def test_keyswitch_wide_digits(emu_lib, ref_backend_cls):
    """Digits of 10 limbs (dnum = 2 at 20 limbs): the base conversion's one-coefficient-per-thread form
    (more than six source limbs) and its 16-term exact dot product."""
    P = make_params(12, 19, dnum=2)
    assert P.alpha == 10
    kp.check_keyswitch(P, _emu(P, emu_lib, 2), ref_backend_cls(P), levels=[19, 12, 4])


def test_engine_ops(emu_lib, ref_backend_cls):
    P = make_params(12, 6)
    eg, er = make_engines(P, ref_backend_cls, _emu(P, emu_lib))
    kp.check_engine_ops(eg, er)


def test_fused_lut_services(emu_lib, ref_backend_cls):
    P = make_params(12, 13)
    kp.check_fused_services(P, _emu(P, emu_lib), ref_backend_cls(P), batch=2)


@pytest.mark.parametrize("log_n,lvl,dnum", [(12, 6, 0), (12, 7, 4)])
def test_multiply_gather(log_n, lvl, dnum, emu_lib, ref_backend_cls):
    """fused multiply with gathered operands (pointer table as a kernel parameter) against the oracle"""
    P = make_params(log_n, lvl, dnum=dnum)
    kp.check_multiply_gather(P, _emu(P, emu_lib), ref_backend_cls(P))


@pytest.mark.parametrize("views", [False, True])
def test_bit_sliced_round_on_emulator(views, emu_lib, ref_backend_cls, monkeypatch):
    """ARK_0, ShiftRows, SubBytes, MixColumns + ARK of the bit-sliced AES path through the real kernels (gathered
    products, fused LUT sums; views: the LUT inputs as strided batch slices, in_poly_stride of fhe_lincomb /
    fhe_tensor_acc) against the oracle at N = 2^12: identical residues, plain AES bytes"""
    from test_aes_bits import make_service, KEY_B
    from aes_fhe_b200.services import aes_bits
    monkeypatch.setattr(aes_bits, "LUT_VIEWS", views)
    from aes_fhe_b200.services.key_expansion import expand_key
    from oracle import aes_plain as A
    P = make_params(12, 9, scale_bits=44)
    res = []
    for be in (_emu(P, emu_lib), ref_backend_cls(P)):
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


@pytest.mark.parametrize("chunk", [1, 3, 5])
def test_chained_ntt_many_chunks(chunk, emu_lib, ref_backend_cls, monkeypatch):
    """The chained single-launch NTT with tiny chunks (FHE_CHAIN_ROWS): several segments, a ragged last
    chunk, padding CTAs and skipped rows (ModUp layout) all occur at these sizes."""
    monkeypatch.setenv("FHE_CHAIN_ROWS", str(chunk))
    P = make_params(12, 6, dnum=4)
    b = _emu(P, emu_lib, 2)
    kp.check_primitives(P, b, ref_backend_cls(P))
    kp.check_keyswitch(P, b, ref_backend_cls(P))
    kp.check_rescale(P, b, ref_backend_cls(P))
    assert b.lib.fhe_ntt_fused_status(b.ctx) == 0


@pytest.mark.parametrize("log_n,lvl,dnum,half_width,batch,max_baby", [(12, 5, 4, 20, 5, 32), (12, 5, 4, 20, 5, 16), (12, 6, 2, 6, 2, 32),
                                                                      (13, 4, 1, 31, 1, 32), (13, 4, 1, 31, 1, 16)])
def test_double_hoisted_transform(log_n, lvl, dnum, half_width, batch, max_baby, emu_lib, ref_backend_cls, monkeypatch):
    """fhe_bsgs_inner (all baby steps of a BSGS transform in one pass, extended basis) against the oracle: 40
    diagonals -> 32 baby x 2 giant steps (two kernel passes summed) or 16 x 3, with a ragged batch chunk; 12 -> 8 x 2;
    62 -> 32 x 2 or 16 x 4 with one digit"""
    from aes_fhe_b200 import bootstrap as B
    monkeypatch.setattr(B, "BSGS_MAX_BABY", max_baby)
    P = make_params(log_n, lvl, dnum=dnum)
    kp.check_double_hoisted_transform(P, _emu(P, emu_lib), ref_backend_cls(P), batch=batch, half_width=half_width)
