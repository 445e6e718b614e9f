"""The reference's UNMODIFIED Python files, imported from /root/reference through the
`desilofhe` shim, running on the facade with the CPU oracle backend.  Skipped where the
reference tree is not mounted (the GPU box)."""
import sys

import numpy as np
import pytest

from aes_fhe_b200 import compat
from aes_fhe_b200 import engine as E
from aes_fhe_b200.params import make_params
from conftest import REFERENCE

pytestmark = pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref_modules(ref_backend_cls):
    P = make_params(12, 12)

    class OracleEngine(E.Engine):
        def __init__(self, *a, **k):
            k.pop("max_level", None)
            super().__init__(_backend=ref_backend_cls(P), _params=P, seed=3)

        def create_rotation_key(self, sk, steps=None):
            return super().create_rotation_key(sk, steps=[4, 8, 12, -4, -8, -12, 5])

    compat.install(OracleEngine)
    compat.mount_reference(REFERENCE)
    import aes_xor_fhe.xor_service as xs
    import aes_xor_fhe.sbox.sbox_service as sb
    import aes_xor_fhe.shiftrows_service as sr
    import aes_xor_fhe.engine_context as ec
    import new as newmod
    return dict(xs=xs, sb=sb, sr=sr, ec=ec, new=newmod)


def test_reference_xor_service_unmodified(ref_modules):
    xs = ref_modules["xs"]
    cfg = xs.XORConfig(coeffs_path=REFERENCE / "xor_mono_coeffs.json")
    eng = xs.EngineWrapper(cfg)
    svc = xs.XORService(eng, xs.CoefficientCache(cfg.coeffs_path))
    rng = np.random.default_rng(0)
    sc = eng.engine.slot_count
    a = rng.integers(0, 16, size=sc, dtype=np.uint8)
    b = rng.integers(0, 16, size=sc, dtype=np.uint8)
    assert np.array_equal(svc.xor(a, b), a ^ b)                       # test_xor_random
    assert np.array_equal(svc.xor(np.array([0, 1, 2, 3], np.uint8), np.array([3, 2, 1, 0], np.uint8))[:4],
                          [3, 3, 3, 3])                               # test_xor_simple


def test_reference_sbox_service_unmodified(ref_modules):
    sb, ec = ref_modules["sb"], ref_modules["ec"]
    ctx = ec.EngineContext(signature=2, max_level=22, mode="parallel", thread_count=8, device_id=0)
    svc = sb.SBoxService(ctx)
    sc = ctx.engine.slot_count
    x = np.tile(np.arange(256, dtype=np.uint8), sc // 256 + 1)[:sc]
    from aes_xor_fhe.utils import zeta_encode, zeta_decode
    out = svc.sub_bytes_array(ctx.engine.encrypt(zeta_encode(x, modulus=256), ctx.public_key))
    res = zeta_decode(ctx.engine.decrypt(out, ctx.secret_key), modulus=256)
    assert np.array_equal(res, np.array(sb.AES_SBOX, dtype=np.uint8)[x])          # test_sbox_array_simd


def test_reference_full_round_and_shiftrows_unmodified(ref_modules):
    xs, newmod, sr = ref_modules["xs"], ref_modules["new"], ref_modules["sr"]
    cfg = xs.XORConfig(coeffs_path=REFERENCE / "xor_mono_coeffs.json")
    eng = xs.EngineWrapper(cfg)
    svc = xs.XORService(eng, xs.CoefficientCache(cfg.coeffs_path))
    rng = np.random.default_rng(1)
    sc = eng.engine.slot_count
    state = rng.integers(0, 256, size=sc, dtype=np.uint8)
    key = rng.integers(0, 256, size=sc, dtype=np.uint8)
    out = newmod.AESFHERound(eng, svc).full_round(state, key, recombine=True)    # new.py:248-261
    assert np.array_equal(out, state ^ key)
    # shiftrows_service on a 16-slot state: the engine must reproduce the reference's own
    # (mask, rotate by -4r, add) arithmetic exactly as a plain-complex evaluation does (D6
    # documents that this is not AES ShiftRows on a 2^k-slot ring; parity is with the op sequence)
    shr = sr.AESFHEShiftRows(eng, svc)
    v = xs.ZetaEncoder.to_zeta(np.arange(16, dtype=np.uint8))
    dec = eng.decrypt(shr.shift_rows(eng.encrypt(v)))
    full = np.zeros(sc, dtype=np.complex128); full[:16] = v
    expect = np.zeros(sc, dtype=np.complex128)
    for r in range(4):
        m = np.zeros(sc); m[r:16:4] = 1.0
        expect += np.roll(full * m, -4 * r)
    assert np.abs(dec - expect).max() < 1e-5
