"""Bootstrapping on the CPU oracle at small rings; the B200 run at N = 2^16 is in
test_gpu_aes.py."""
import numpy as np
import pytest

from aes_fhe_b200 import bootstrap as B
from aes_fhe_b200.engine import Engine
from aes_fhe_b200.params import make_params


def test_special_fft_factorisation_and_grouping():
    n = 64
    M = 4 * n
    xi = np.exp(2j * np.pi / M)
    U = np.array([[xi ** ((pow(5, j, M) * k) % M) for k in range(n)] for j in range(n)])
    layers, inv_layers = B._fft_layers(n)
    br = np.array([int(format(i, "06b")[::-1], 2) for i in range(n)])
    c = np.random.default_rng(0).standard_normal(n) + 1j * np.random.default_rng(1).standard_normal(n)

    def apply(D, x):
        return sum(v * np.roll(x, -d) for d, v in D.items())

    for groups in (1, 2, 3):
        y = c[br]
        for G in B._group(layers, groups, n):
            y = apply(G, y)
        assert np.abs(y - U @ c).max() < 1e-11
        x = U @ c
        for G in B._group(list(reversed(inv_layers)), groups, n):
            x = apply(G, x)
        assert np.abs(x - c[br]).max() < 1e-11


def test_evalmod_polynomial_recovers_the_message():
    rho = 32.0
    poly, alphas = B._evalmod_design(rho, "monomial")
    rng = np.random.default_rng(0)
    I = rng.integers(-(B.K_NORM - 1), B.K_NORM, 20000)
    msg = rng.uniform(-0.2, 0.2, I.size)
    y = (msg / rho + I) / B.K_NORM
    c = np.polynomial.polynomial.polyval(y, poly)
    for i in range(B.DOUBLE_ANGLES):
        c = c * c - alphas[i + 1]
    assert np.abs(c - msg).max() < 1e-4            # msg * (2 pi eps)^2 / 6 sine error at eps <= 0.2/32


@pytest.mark.parametrize("log_n,scale_bits,tol", [(11, 40, 3e-3), (12, 44, 5e-4)])
def test_bootstrap_refreshes_levels_on_oracle(log_n, scale_bits, tol, ref_backend_cls):
    P = make_params(log_n, 22, scale_bits=scale_bits)
    eng = Engine(_params=P, _backend=ref_backend_cls(P), seed=5, use_bootstrap=True)
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    rlk = eng.create_relinearization_key(sk)
    cj = eng.create_conjugation_key(sk)
    bk = eng.create_bootstrap_key(sk)
    rng = np.random.default_rng(0)
    v = np.exp(-2j * np.pi * rng.integers(0, 16, (2, eng.slot_count)) / 16)       # batch of 2
    ct = eng.encrypt(v, pk, level=2)
    out = eng.bootstrap(ct, rlk, cj, bk)
    assert out.level == 22 - 18 and out.batch == 2
    d = eng.decrypt(out, sk)
    assert np.abs(d - v).max() < tol
    # the refreshed ciphertext is usable: one more product
    sq = eng.multiply(out, out, rlk)
    assert np.abs(eng.decrypt(sq, sk) - v * v).max() < 4 * tol
    with pytest.raises(RuntimeError):
        eng.bootstrap(ct, rlk, cj, eng.create_small_bootstrap_key(sk))


def test_chebyshev_paterson_stockmeyer_matches_numpy(ref_backend_cls):
    """The Chebyshev-basis Paterson-Stockmeyer evaluation used by EvalMod (10 products for degree
    22) against numpy's chebval on the decrypted slots."""
    from numpy.polynomial import chebyshev as C
    P = make_params(11, 8)
    eng = Engine(_params=P, _backend=ref_backend_cls(P), seed=3, use_bootstrap=True)
    sk = eng.create_secret_key(); pk = eng.create_public_key(sk); rlk = eng.create_relinearization_key(sk)
    rng = np.random.default_rng(1)
    yv = rng.uniform(-1, 1, eng.slot_count)
    coeffs = rng.standard_normal(23) / (1 + np.arange(23))
    ct = eng.encrypt(yv, pk)
    c0 = eng.op_counts.get("mul_ct", 0)
    out = B.chebyshev_eval_ps(eng, rlk, ct, coeffs)
    assert eng.op_counts["mul_ct"] - c0 == 10 and out.level == ct.level - 5
    assert np.abs(eng.decrypt(out, sk).real - C.chebval(yv, coeffs)).max() < 1e-6
