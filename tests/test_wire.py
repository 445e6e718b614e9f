"""Wire format: round trips on the CPU oracle, rejection of foreign / damaged data."""
import numpy as np
import pytest

from aes_fhe_b200 import wire
from aes_fhe_b200.engine import Engine
from aes_fhe_b200.params import make_params


def _engine(ref_backend_cls, lvl=5, seed=4):
    P = make_params(12, lvl)
    return Engine(_params=P, _backend=ref_backend_cls(P), seed=seed)


def test_round_trip_of_ciphertexts_and_evaluation_keys(ref_backend_cls):
    owner = _engine(ref_backend_cls)
    sk = owner.create_secret_key(); pk = owner.create_public_key(sk)
    rlk = owner.create_relinearization_key(sk); cj = owner.create_conjugation_key(sk)
    rk = owner.create_fixed_rotation_key(sk, 3)
    server = _engine(ref_backend_cls, seed=99)                       # another process: same parameters, no secret key
    pk2, rlk2, cj2, rk2 = (wire.loads(server, wire.dumps(owner, k)) for k in (pk, rlk, cj, rk))
    assert rk2.delta == 3 and rk2.galois == rk.galois
    rng = np.random.default_rng(0)
    v = np.exp(-2j * np.pi * rng.integers(0, 16, (2, owner.slot_count)) / 16)
    ct = wire.loads(server, wire.dumps(owner, owner.encrypt(v, pk)))
    assert ct.level == 5 and ct.batch == 2
    out = server.rotate(server.conjugate(server.multiply(ct, ct, rlk2), cj2), rk2)
    back = wire.loads(owner, wire.dumps(server, out))
    assert np.abs(owner.decrypt(back, sk) - np.roll(np.conj(v * v), 3, axis=-1)).max() < 1e-5
    # a client encrypts with the shipped public key
    assert np.abs(owner.decrypt(server.encrypt(v[0], pk2), sk) - v[0]).max() < 1e-5


def test_rejects_foreign_parameters_damage_and_secret_keys(ref_backend_cls):
    a, b = _engine(ref_backend_cls, 5), _engine(ref_backend_cls, 4)
    sk = a.create_secret_key(); pk = a.create_public_key(sk)
    blob = wire.dumps(a, a.encrypt(np.ones(8), pk))
    with pytest.raises(wire.WireError, match="different parameter set"):
        wire.loads(b, blob)
    with pytest.raises(wire.WireError, match="bad magic"):
        wire.loads(a, b"XXXXXXXX" + blob[8:])
    with pytest.raises(wire.WireError, match="payload size"):
        wire.loads(a, blob[:-8])
    bad = bytearray(blob); bad[-1] = 0xFF                              # top byte of the last residue: >= modulus
    with pytest.raises(wire.WireError, match="out of range"):
        wire.loads(a, bytes(bad))
    with pytest.raises(wire.WireError, match="not serialisable"):
        wire.dumps(a, sk)


def test_rejects_wrong_shapes_before_they_reach_a_kernel(ref_backend_cls):
    """undersized keys, 1-polynomial ciphertexts, wrong batch / digit dimensions and Galois elements
    that do not match their rotation must raise WireError, never IndexError or an out-of-bounds device read."""
    eng = _engine(ref_backend_cls)
    P = eng.params
    sk = eng.create_secret_key(); pk = eng.create_public_key(sk)
    rlk = eng.create_relinearization_key(sk); rk = eng.create_fixed_rotation_key(sk, 3)
    base = {"params": wire.params_digest(P)}
    good = eng.backend.to_numpy(rlk.data)

    def load(kind, header, arr):
        return wire.loads(eng, wire._pack(kind, dict(base, **header), arr))

    assert isinstance(load("relinearization_key", {}, good), type(rlk))
    for bad in (good[:, :, :, :P.n_q],                       # q-limbs only: kernels would read n_q + n_p rows
                good[:P.dnum - 1],                           # a digit short
                good[:, :1],                                 # one polynomial
                np.concatenate([good, good], axis=2),        # batch 2
                good[..., : P.n // 2]):                      # half a ring
        with pytest.raises(wire.WireError):
            load("relinearization_key", {}, bad)
    ct = eng.backend.to_numpy(eng.encrypt(np.ones(8), pk).polys)
    assert load("ciphertext", {"level": 5}, ct).npoly == 2
    for hdr, bad in (({"level": 5}, ct[:1]),                 # 1 polynomial
                     ({"level": 5}, np.concatenate([ct, ct], axis=0)),      # 4 polynomials
                     ({"level": 5}, ct[:, :0]),              # empty batch
                     ({"level": 4}, ct),                     # level / limb mismatch
                     ({"level": 99}, ct),
                     ({"level": "5"}, ct),
                     ({"level": 5}, ct[0])):                 # 3-d
        with pytest.raises(wire.WireError):
            load("ciphertext", hdr, bad)
    with pytest.raises(wire.WireError):
        load("public_key", {}, eng.backend.to_numpy(pk.polys)[0])           # was an IndexError
    with pytest.raises(wire.WireError):
        load("public_key", {}, eng.backend.to_numpy(pk.polys)[:, :, :3])
    rot = eng.backend.to_numpy(rk.data)
    assert load("fixed_rotation_key", {"galois": int(rk.galois), "delta": 3}, rot).delta == 3
    for hdr in ({"galois": int(rk.galois), "delta": 4}, {"galois": 4, "delta": 3}, {"galois": int(rk.galois)}, {"delta": 3}):
        with pytest.raises(wire.WireError):
            load("fixed_rotation_key", hdr, rot)
    with pytest.raises(wire.WireError):
        load("conjugation_key", {"galois": int(rk.galois)}, rot)


def test_byte_codec_fallback_matches_zeta_encoder(ref_backend_cls):
    """Engine.encrypt_zeta / decrypt_zeta (device codec on the GPU) on a backend without a device codec:
    same values as ZetaEncoder.to_zeta + encrypt / decrypt + from_zeta (xor_service.py:132-145)."""
    from aes_fhe_b200.services.xor_service import ZetaEncoder
    eng = _engine(ref_backend_cls)
    sk = eng.create_secret_key(); pk = eng.create_public_key(sk)
    x = np.random.default_rng(2).integers(0, 256, (2, eng.slot_count), dtype=np.uint8)
    ct = eng.encrypt_zeta(x, pk, 256)
    assert np.array_equal(ZetaEncoder.from_zeta(eng.decrypt(ct, sk), 256), x)
    assert np.array_equal(eng.decrypt_zeta(ct, sk, 256), x)
    assert np.array_equal(eng.decrypt_zeta(eng.encrypt_zeta(x[0, :10] % 16, pk), sk)[:10], x[0, :10] % 16)
