import math
from fractions import Fraction

import pytest

from aes_fhe_b200.params import make_params, is_prime, bit_reverse


@pytest.mark.parametrize("log_n,lvl", [(12, 8), (13, 12), (16, 30), (16, 22)])
def test_chain_is_ntt_friendly_and_distinct(log_n, lvl):
    P = make_params(log_n, lvl)
    assert len(set(P.moduli)) == len(P.moduli)
    for q, psi in zip(P.moduli, P.psi):
        assert is_prime(q) and q % (2 * P.n) == 1 and q < (1 << 45)
        assert pow(psi, P.n, q) == q - 1            # primitive 2N-th root
    assert P.n_q == lvl + 1 and P.n_p >= 1


def test_scale_recursion_and_drift():
    P = make_params(16, 30)
    for l in range(P.max_level, 0, -1):
        assert abs(float(P.delta[l] * P.delta[l] / P.moduli[l] / P.delta[l - 1]) - 1) < 1e-30
        assert abs(P.scale(l) / 2.0 ** 40 - 1) < 1e-4    # drift stays bounded
    # special modulus exceeds the largest digit: hybrid key switching noise condition
    pprod = math.prod(P.p)
    for j in range(P.dnum):
        assert pprod > math.prod(P.q[j * P.alpha:(j + 1) * P.alpha])
    # 128-bit security budget at N = 2^16 (log PQ <= ~1770 for ternary secrets)
    assert sum(math.log2(m) for m in P.moduli) < 1770


def test_rotation_galois_elements():
    P = make_params(12, 4)
    assert P.galois_for_rotation(0) == 1
    assert P.galois_for_rotation(-1) == 5
    assert (P.galois_for_rotation(1) * 5) % (2 * P.n) == 1
    assert P.galois_conj == 2 * P.n - 1


def test_bit_reverse():
    assert [bit_reverse(i, 3) for i in range(8)] == [0, 4, 2, 6, 1, 5, 3, 7]
