"""Service-level evaluation schedules (the K8 row of SURVEY.md section 2.2): bivariate
16x16 LUT polynomials in two zeta_16-valued ciphertexts, with lazy relinearisation.

    f(x, y) = sum_i x^i * ( sum_j c_ij y^j )

* the inner sums are pure constant multiplications (no key switch);
* the level / scale alignment of every operand is folded into the constants, so no
  operand is rescaled on its own;
* the 15 outer products of one output are accumulated as a 3-polynomial ciphertext and
  relinearised once, then rescaled twice.

Depth: 3 (power basis) + 2 = 5 levels, the same as the reference's ``xor_cipher``
(/root/reference/xor_service.py:271-286), with 1 relinearisation per output instead of one
per monomial.  Several outputs (the six S-box / 2S / 3S nibble planes) share the bases.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, List, Sequence

import numpy as np

from .engine import Ciphertext


def power_basis_16(eng_wrap, ct) -> Dict[int, Ciphertext]:
    """{k: ct^k} for k = 1..15 of a unit-modulus (zeta_16) ciphertext: 7 products and 7
    conjugations (t^(16-k) = conj(t^k)), as /root/reference/xor_service.py:245-254."""
    pos = eng_wrap.make_power_basis(ct, 8)
    basis = {k: c for k, c in enumerate(pos, 1)}
    for k in range(1, 8):
        basis[16 - k] = eng_wrap.conjugate(pos[k - 1])
    return basis


def _const_to_residues(eng, value: complex, scale: Fraction, nq: int):
    re = int(round(Fraction(float(value.real)) * scale))
    im = int(round(Fraction(float(value.imag)) * scale))
    return eng._const_residues(re, im, nq)


def bivariate_lut(eng_wrap, ct_x, ct_y, coeff_mats: Sequence[np.ndarray],
                  bx: Dict[int, Ciphertext] | None = None,
                  by: Dict[int, Ciphertext] | None = None) -> List[Ciphertext]:
    eng = eng_wrap.engine
    be, P = eng.backend, eng.params
    if bx is None:
        bx = power_basis_16(eng_wrap, ct_x)
    if by is None:
        by = power_basis_16(eng_wrap, ct_y)
    lo = min(min(c.level for c in bx.values()), min(c.level for c in by.values()))
    if lo < 2:
        raise RuntimeError("bivariate_lut: not enough levels left")
    nq = lo + 1
    bt = max(ct_x.batch, ct_y.batch)
    # scale of the accumulated products before the two closing rescales
    target = P.delta[lo - 2] * P.moduli[lo] * P.moduli[lo - 1]
    xs = {i: (be.take_limbs(c.polys, nq, False) if c.level > lo else c.polys) for i, c in bx.items()}
    ys = {j: (be.take_limbs(c.polys, nq, False) if c.level > lo else c.polys) for j, c in by.items()}
    outs = []
    for C in coeff_mats:
        C = np.asarray(C, dtype=np.complex128)
        acc = be.zeros(3, bt, nq, False)
        for i in range(16):
            row = C[i]
            if not np.any(np.abs(row) > 1e-13):
                continue
            # scale the inner sum must carry so that x^i * inner sits at `target`
            s_in = target if i == 0 else target / P.delta[bx[i].level]
            inner = None
            for j in range(1, 16):
                if abs(row[j]) <= 1e-13:
                    continue
                cp, cm = _const_to_residues(eng, row[j], s_in / P.delta[by[j].level], nq)
                term = be.mul_const(ys[j], cp, cm, nq)
                inner = term if inner is None else be.add(inner, term, nq, 0)
            if inner is None:
                inner = be.zeros(2, bt, nq, False)
            if abs(row[0]) > 1e-13:
                cp, cm = _const_to_residues(eng, row[0], s_in, nq)
                inner = be.add_const(inner, cp, cm, nq)
            if i == 0:
                acc = be.concat([be.add(be.take_polys(acc, 2), inner, nq, 0), be.select_poly(acc, 2)])
            else:
                acc = be.add(acc, be.tensor(xs[i], inner, nq), nq, 0)
        ct3 = Ciphertext(eng, acc, lo)
        ct2 = eng._relin(ct3, eng_wrap.relin_key)
        outs.append(eng._rescale(eng._rescale(ct2)))
    return outs
