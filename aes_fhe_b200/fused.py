"""Service-level evaluation schedules (rows K7/K8 of SURVEY.md section 2.2): LUT polynomials
evaluated with baby-step/giant-step structure, constant-only inner sums and lazy
relinearisation.  Everything is batched: the ciphertext operands may carry any batch size.

    univariate (Paterson-Stockmeyer):  p(t)    = sum_m  (t^B)^m * ( sum_j c[B m + j] t^j )
    bivariate 16x16:                   f(x, y) = sum_i  x^i     * ( sum_j c[i][j]  y^j )

* the inner sums are pure constant multiplications -- one pass of ``fhe_lincomb`` reads every
  power once and writes all inner sums of all outputs; the level / scale alignment of every
  operand is folded into the constants, so nothing is rescaled on its own;
* the outer products of one output are accumulated as a 3-polynomial ciphertext
  (``fhe_tensor_acc``), relinearised ONCE together with the first rescale (one ModDown by P q_l), then
  rescaled once more.

Depth: baby basis + giant basis + 2.  For the reference's degree-255 S-box pair
(/root/reference/sbox/sbox_service.py:116-138: 254 relinearisations, 10 levels) that is
4 + 4 + 2 (+1 for hi x lo) = 11 levels and 15 + 14 + 2 + 1 = 32 key switches (23 and 10 levels in the
folded single-polynomial form of SBoxService.sub_bytes_array_bsgs); for the 4-bit XOR
(/root/reference/xor_service.py:271-286: 92 key switches, 5 levels) 5 levels and 17 key switches (odd powers only: 5 products per operand, 4 conjugations for y, two
lazily relinearised outer sums and one conjugation).
Same slot values within CKKS noise; residues differ from the reference operation order.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import Ciphertext

_EPS = 1e-13


def _lim(x: Fraction) -> Fraction:
    return x.limit_denominator(1 << 600)


def power_basis_16(eng_wrap, ct, wanted: Optional[Sequence[int]] = None, with_dev: bool = False):
    """{k: ct^k} of a unit-modulus (zeta_16) ciphertext for the exponents in `wanted` (default: all
    of 1..15), with t^(16-k) = conj(t^k) as /root/reference/xor_service.py:245-254.  Only the powers
    that are wanted, or needed to build a wanted one (t^k = t^ceil(k/2) * t^floor(k/2), depth 3), are
    computed: the XOR polynomial has odd exponents only, which takes 5 products and 4 conjugations
    instead of 7 and 7.

    ``with_dev``: products of operands at different levels use the higher one in place (upper limbs
    ignored, Engine._mul_ct_dropped) instead of bringing it down with a constant multiply + rescale;
    returns ``(basis, dev)`` where ``dev[k]`` is the factor by which the scale of ``basis[k]`` exceeds
    the table scale of its level (the LUT constants absorb it)."""
    eng, rlk = eng_wrap.engine, eng_wrap.relin_key
    wanted = sorted(set(range(1, 16) if wanted is None else wanted))
    low = sorted({k if k <= 8 else 16 - k for k in wanted})          # positive powers 1..8 behind them
    need = set()

    def req(k):
        if k > 1 and k not in need:
            need.add(k)
            req((k + 1) // 2)
            req(k // 2)

    for k in low:
        req(k)
    pw: Dict[int, Ciphertext] = {1: ct}
    dv: Dict[int, Fraction] = {1: Fraction(1)}
    for k in sorted(need):
        hi, lo = (k + 1) // 2, k // 2
        if with_dev:
            pw[k], d = eng._mul_ct_dropped(pw[hi], pw[lo], rlk)
            dv[k] = _lim(d * dv[hi] * dv[lo])
        else:
            pw[k] = eng.multiply(pw[hi], pw[lo], rlk)
            dv[k] = Fraction(1)
    basis, dev = {}, {}
    for k in wanted:
        basis[k] = pw[k] if k <= 8 else eng_wrap.conjugate(pw[16 - k])
        dev[k] = dv[k] if k <= 8 else dv[16 - k]
    return (basis, dev) if with_dev else basis


def lazy_power_basis(eng, rlk, ct, degree: int, dev0: Fraction = Fraction(1)):
    """[ct^1 .. ct^degree] with the recursion of Engine.make_power_basis (ct^k = ct^ceil(k/2) ct^floor(k/2)),
    mixed-level products done with Engine._mul_ct_dropped.  Returns (powers, devs)."""
    pw: List[Optional[Ciphertext]] = [None] * (degree + 1)
    dv: List[Fraction] = [Fraction(1)] * (degree + 1)
    pw[1], dv[1] = ct, dev0
    for k in range(2, degree + 1):
        hi, lo = (k + 1) // 2, k // 2
        pw[k], d = eng._mul_ct_dropped(pw[hi], pw[lo], rlk)
        dv[k] = _lim(d * dv[hi] * dv[lo])
    return pw[1:], dv[1:]


def _const_pair(eng, value: complex, scale: Fraction, nq: int):
    re = int(round(Fraction(float(value.real)) * scale))
    im = int(round(Fraction(float(value.imag)) * scale))
    return eng._const_residues(re, im, nq)


def _outer_sum(eng, relin_key, outer: Dict[int, Ciphertext], inner_basis: Dict[int, Ciphertext],
               coeff_mats: Sequence[np.ndarray], cache_key, outer_dev: Optional[Dict[int, Fraction]] = None,
               inner_dev: Optional[Dict[int, Fraction]] = None, batched: bool = False):
    """sum_i outer[i] * (sum_j C[i][j] inner_basis[j]),  i = 0 meaning the constant 1 (same for
    j = 0).  ``coeff_mats`` is a list of (n_outer+1) x (n_inner+1) complex matrices; all outputs
    share the two bases.  Returns one ciphertext per matrix, two levels below the lowest
    operand.  ``outer_dev`` / ``inner_dev``: scale deviations of the basis elements (power_basis_16,
    lazy_power_basis); the constants of the inner sums are divided by them, so the outputs sit exactly on the
    scale table.  ``batched``: return ONE ciphertext whose batch holds the outputs one after the other (what a
    concatenation of the list would give) -- on the B200 the sums of all outputs then land in one accumulator and are
    relinearised and rescaled together (one key switch over n_out x batch ciphertexts instead of n_out small ones)."""
    be, P = eng.backend, eng.params
    lo = min([c.level for c in outer.values()] + [c.level for c in inner_basis.values()])
    if lo < 2:
        raise RuntimeError("LUT evaluation: not enough levels left")
    nq = lo + 1
    target = P.delta[lo - 2] * P.moduli[lo] * P.moduli[lo - 1]     # scale before the two rescales
    n_out = len(coeff_mats)
    i_list = sorted(outer)                       # outer powers actually present
    j_list = sorted(inner_basis)
    odev = {i: Fraction(1) for i in i_list} if outer_dev is None else outer_dev
    idev = {j: Fraction(1) for j in j_list} if inner_dev is None else inner_dev
    rows = [(m, i) for m in range(n_out) for i in [0] + i_list
            if np.any(np.abs(np.asarray(coeff_mats[m])[i]) > _EPS)]

    cache = eng.__dict__.setdefault("_lut_cache", {})
    key = (cache_key, lo, tuple(outer[i].level for i in i_list), tuple(inner_basis[j].level for j in j_list),
           tuple(odev[i] for i in i_list), tuple(idev[j] for j in j_list))
    prep = cache.get(key)
    if prep is None:
        const_res, c0_res = [], []
        for (m, i) in rows:
            C = np.asarray(coeff_mats[m], dtype=np.complex128)
            s_in = target if i == 0 else target / (P.delta[outer[i].level] * odev[i])
            const_res.append([_const_pair(eng, C[i, j], s_in / (P.delta[inner_basis[j].level] * idev[j]), nq)
                              for j in j_list])
            c0_res.append(_const_pair(eng, C[i, 0], s_in, nq))
        prep = be.prepare_lincomb(const_res, c0_res, nq)
        cache[key] = prep
    inner = be.lincomb([inner_basis[j].polys for j in j_list], prep)      # one pass, all inner sums

    bt = max([inner[0].shape[1]] + [c.batch for c in outer.values()])
    per_out = [[(i, inner[r]) for r, (mm, i) in enumerate(rows) if mm == m] for m in range(n_out)]
    if batched and getattr(be, "tensor_acc_into", False) and all(any(i for i, _ in terms) for terms in per_out):
        acc_all = be.alloc((3, n_out * bt, nq, P.n))
        for m, terms in enumerate(per_out):
            zero = [t for i, t in terms if i == 0]
            be.tensor_acc(None, [outer[i].polys for i, _ in terms if i], [t for i, t in terms if i], nq,
                          out=acc_all[:, m * bt:(m + 1) * bt], init=zero[0] if zero else None)
        eng._count('keyswitch_relin')
        eng._count('rescale')
        return eng._rescale(Ciphertext(eng, be.relin_rescale(acc_all, relin_key.data, nq), lo - 1))

    outs = []
    for m in range(n_out):
        a_list, b_list, acc = [], [], None
        zero_term = None
        for r, (mm, i) in enumerate(rows):
            if mm != m:
                continue
            if i == 0:
                zero_term = inner[r]
            else:
                a_list.append(outer[i].polys)
                b_list.append(inner[r])
        bt = max([inner[0].shape[1]] + [c.batch for c in outer.values()])
        if a_list:
            acc = be.tensor_acc(None, a_list, b_list, nq)
        else:
            acc = be.zeros(3, bt, nq, False)
        if zero_term is not None:
            if hasattr(be, "add_into_polys"):
                acc = be.add_into_polys(acc, zero_term, nq)              # in place on polynomials 0 and 1
            else:
                acc = be.concat([be.add(be.take_polys(acc, 2), be.expand_batch(zero_term, bt), nq, 0), be.select_poly(acc, 2)])
        # relinearise and do the first of the two rescales in one ModDown by P * q_lo (as Engine.multiply does)
        eng._count('keyswitch_relin')
        eng._count('rescale')
        ct2 = Ciphertext(eng, be.relin_rescale(acc, relin_key.data, nq), lo - 1)
        outs.append(eng._rescale(ct2))
    if batched:
        return Ciphertext(eng, be.concat_batch([c.polys for c in outs]), outs[0].level)
    return outs


def bivariate_lut(eng_wrap, ct_x, ct_y, coeff_mats: Sequence[np.ndarray],
                  bx: Optional[Dict[int, Ciphertext]] = None,
                  by: Optional[Dict[int, Ciphertext]] = None, cache_key="biv") -> List[Ciphertext]:
    """f_m(x, y) = sum_ij C_m[i][j] x^i y^j for 16x16 coefficient matrices (zeta_16 inputs)."""
    eng = eng_wrap.engine
    used_i = [i for i in range(1, 16) if any(np.any(np.abs(np.asarray(C)[i]) > _EPS) for C in coeff_mats)]
    used_j = [j for j in range(1, 16) if any(np.any(np.abs(np.asarray(C)[:, j]) > _EPS) for C in coeff_mats)]
    hi_i = [i for i in used_i if i > 8]
    if bx is None and by is None and 2 * len(coeff_mats) < len(hi_i):
        # Few outputs: do not build conj(x^k) at all.  With x^i = conj(x^(16-i)) for i > 8,
        #     f = sum_{i<=8} x^i g_i(y)  +  conj( sum_{i>8} x^(16-i) conj(g_i)(y) ),
        # and conj(g_i)(y) = sum_j conj(c_ij) y^((16-j) % 16) is again a constant-only sum over the
        # y basis.  Two outer sums and ONE conjugation per output replace the conjugations of the
        # x basis (XOR: 17 key switches instead of 19).
        by, dy = power_basis_16(eng_wrap, ct_y, sorted(set(used_j) | {(16 - j) % 16 for j in used_j if j}), with_dev=True)
        bx, dx = power_basis_16(eng_wrap, ct_x, sorted({i if i <= 8 else 16 - i for i in used_i}), with_dev=True)
        mats = []
        for C in coeff_mats:
            C = np.asarray(C, dtype=np.complex128)
            A = C.copy(); A[9:] = 0.0
            B = np.zeros_like(C)
            for i in hi_i:
                for j in range(16):
                    B[16 - i, (16 - j) % 16] = np.conj(C[i, j])
            mats += [A, B]
        outs = _outer_sum(eng, eng_wrap.relin_key, {i: bx[i] for i in sorted(bx)}, {j: by[j] for j in sorted(by)},
                          mats, (cache_key, "conj-split", len(coeff_mats)), dx, dy)
        return [eng.add(outs[2 * m], eng_wrap.conjugate(outs[2 * m + 1])) for m in range(len(coeff_mats))]
    dx = dy = None
    if bx is None:
        bx, dx = power_basis_16(eng_wrap, ct_x, used_i, with_dev=True)
    if by is None:
        by, dy = power_basis_16(eng_wrap, ct_y, used_j, with_dev=True)
    outer = {i: bx[i] for i in used_i}
    inner = {j: by[j] for j in used_j}
    return _outer_sum(eng, eng_wrap.relin_key, outer, inner, coeff_mats, (cache_key, len(coeff_mats)), dx, dy)


def poly_eval_bsgs(engine, relin_key, ct, coeff_vecs: Sequence[np.ndarray], baby: int = 16,
                   cache_key="ps") -> List[Ciphertext]:
    """Paterson-Stockmeyer evaluation of several polynomials of degree < baby^2 on one
    ciphertext, sharing the baby basis t^1..t^baby and the giant basis (t^baby)^1..(baby-1)."""
    deg = max(len(c) for c in coeff_vecs) - 1
    n_giant = deg // baby                                     # highest giant power needed
    pw, dv = lazy_power_basis(engine, relin_key, ct, baby)    # t^1 .. t^baby (+ scale deviations)
    babies = {j: pw[j - 1] for j in range(1, baby)}
    bdev = {j: dv[j - 1] for j in range(1, baby)}
    giants: Dict[int, Ciphertext] = {}
    gdev: Dict[int, Fraction] = {}
    if n_giant >= 1:
        gp, gd = lazy_power_basis(engine, relin_key, pw[baby - 1], n_giant, dv[baby - 1])
        giants = {m: gp[m - 1] for m in range(1, n_giant + 1)}
        gdev = {m: gd[m - 1] for m in range(1, n_giant + 1)}
    mats = []
    for c in coeff_vecs:
        c = np.asarray(c, dtype=np.complex128)
        M = np.zeros((n_giant + 1, baby), dtype=np.complex128)
        for k, v in enumerate(c):
            M[k // baby, k % baby] = v
        mats.append(M)
    return _outer_sum(engine, relin_key, giants, babies, mats, (cache_key, len(mats), baby), gdev, bdev)
