"""CKKS bootstrapping for the full-slot ring (row f-1 of SURVEY.md section 8; the reference
calls ``engine.bootstrap(ct, rlk, conj_key, boot_key)`` from xor_service.py:120-129, :274-277
and mixcolumns_service.py:72-75).

Pipeline (Cheon-Han-Kim-Kim-Song 2018, Han-Ki 2020 double-angle variant):

  0. bring the ciphertext to level 0 with the message scaled so that q_0 / |m| >= 32
  1. ModRaise (``fhe_mod_raise``): plaintext becomes m + q_0 * I, |I| <= K (sparse secret)
  2. CoeffToSlot: the inverse special FFT as `groups` BSGS linear transforms; slots then hold
     the coefficients (a_k + i b_k) / (2 K_n) in bit-reversed order (never undone: EvalMod is
     slot-wise and SlotToCoeff consumes the same order)
  3. real / imaginary parts via one conjugation (multiplication by -i is the monomial
     X^(N/2): free), stacked on the batch axis so EvalMod runs once
  4. EvalMod: alpha_0 cos(2 pi (K_n y - 1/4) / 2^r) by a degree-22 Chebyshev interpolant
     (Paterson-Stockmeyer in the Chebyshev basis: 10 products, depth 5), then r double-angle steps c <- c^2 - alpha_{i+1}; the constants alpha_i fold
     the factor rho / 2 pi so the result is the message coefficient itself
  5. SlotToCoeff: the forward special FFT, `groups` BSGS linear transforms.

Only the first CoeffToSlot matrix sees a non-standard scale (the raised ciphertext has scale
q_0); its plaintext diagonals are encoded at Delta_{L-1} q_L / q_0 so that everything
afterwards sits on the engine's per-level scales.  Depth: 1 + groups + 5 + r + groups
(= 18 for groups = 3, r = 6).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
from numpy.polynomial import chebyshev as _cheb

from .engine import BootstrapKey, Ciphertext, FixedRotationKey, Plaintext

# bit bootstrap (bootstrap_bits): the message sits at +-q_0/4, where sin(2 pi x) is flat, so the refresh squares the
# incoming error and only EvalMod's own error stays.  A double-angle step amplifies an error at the crest of the
# cosine 4x, so the interpolant has to be accurate to 1e-3 / 4^r: r = 5 with degree 18 (5 levels) is one level
# cheaper than r = 6 with degree 12..22 (tools/bits_study.py prints the table)
DOUBLE_ANGLES_BITS = 5
POLY_DEGREE_BITS = 18
import os as _os
_EAGER_PS = _os.environ.get("FHE_EAGER_PS") == "1"        # A/B switch: EvalMod polynomial with level adjustments
_EVEN_EVALMOD = _os.environ.get("FHE_EVEN_EVALMOD", "1") != "0"   # A/B switch: EvalMod polynomial in w = 2 u^2 - 1 (7 products) or in y (9)
BSGS_MAX_BABY = int(_os.environ.get("FHE_BSGS_BABY", "32"))   # baby steps of a double-hoisted transform (16: 16 x 4, 32: 32 x 2 for 63 diagonals)
_NO_DH = _os.environ.get("FHE_NO_DH") == "1"              # A/B switch: single-hoisted linear transforms (square BSGS)
_NO_DH_FUSE = _os.environ.get("FHE_NO_DH_FUSE") == "1"    # A/B switch: double hoisting from separate primitives
RHO_TARGET = 32.0         # q_0 / (scaled message): the message is divided by RHO_TARGET * Delta_0 / q_0
K_NORM = 33               # |I| <= 32 for hamming weight <= 192 (8 sigma)
DOUBLE_ANGLES = 6
POLY_DEGREE = 22


# --------------------------------------------------------------------------- matrices
def _fft_layers(n: int):
    """Butterfly layers of the special FFT U[j,k] = zeta_j^k (zeta_j = xi^(5^j), xi a primitive
    4n-th root): U = L_last ... L_1 BitRev.  A layer is {rotation d: diagonal}, meaning
    out[p] = sum_d diag_d[p] * in[(p + d) % n]."""
    M = 4 * n
    rot = [pow(5, j, M) for j in range(n)]
    ksi = np.exp(2j * np.pi * np.arange(M) / M)
    layers, inv_layers = [], []
    ln = 2
    while ln <= n:
        lenh, lenq = ln // 2, ln * 4
        j = np.arange(lenh)
        w = ksi[(np.array(rot[:lenh]) % lenq) * (M // lenq)]
        first = (np.arange(n) % ln) < lenh
        wfull = np.tile(np.concatenate([w, w]), n // ln)
        d0 = np.where(first, 1.0, -wfull).astype(np.complex128)
        dp = np.where(first, wfull, 0.0).astype(np.complex128)
        dm = np.where(first, 0.0, 1.0).astype(np.complex128)
        i0 = np.where(first, 0.5, -0.5 / wfull).astype(np.complex128)
        ip = np.where(first, 0.5, 0.0).astype(np.complex128)
        im = np.where(first, 0.0, 0.5 / wfull).astype(np.complex128)
        if lenh == n - lenh:
            layers.append({0: d0, lenh: dp + dm})
            inv_layers.append({0: i0, lenh: ip + im})
        else:
            layers.append({0: d0, lenh: dp, n - lenh: dm})
            inv_layers.append({0: i0, lenh: ip, n - lenh: im})
        ln *= 2
    return layers, inv_layers


def _mat_mul(A: Dict[int, np.ndarray], B: Dict[int, np.ndarray], n: int) -> Dict[int, np.ndarray]:
    """A * B (B applied first) in diagonal form."""
    C: Dict[int, np.ndarray] = {}
    for a, va in A.items():
        for b, vb in B.items():
            d = (a + b) % n
            t = va * np.roll(vb, -a)
            C[d] = C[d] + t if d in C else t
    return {d: v for d, v in C.items() if np.abs(v).max() > 1e-15}


def _group(mats: List[Dict[int, np.ndarray]], groups: int, n: int) -> List[Dict[int, np.ndarray]]:
    """mats are applied in list order; merge them into `groups` consecutive products."""
    k = len(mats)
    groups = min(groups, k)
    sizes = [k // groups + (1 if i < k % groups else 0) for i in range(groups)]
    out, pos = [], 0
    for s in sizes:
        G = mats[pos]
        for m in mats[pos + 1:pos + s]:
            G = _mat_mul(m, G, n)
        out.append(G)
        pos += s
    return out


def _bsgs_split(rots: List[int], n: int, double_hoist: bool = False):
    """d = g + b with b in [0, bm): returns (bm, {g: [(b, d), ...]}).  Square split (n1 ~ sqrt of the number of
    diagonals) for the single-hoisted evaluation; with double hoisting the baby steps are cheap (one fused pass,
    no ModDown) and every giant step costs a ModDown + ModUp, so the split is as lopsided as the fused kernel
    takes: up to BSGS_MAX_BABY baby steps (32: the 63 diagonals of a radix-32 factor then need ONE key-switched giant
    step; the fused kernel takes 16 baby steps per pass, so two passes are summed)."""
    nz = [d for d in rots if d]
    stride = n
    for d in nz:
        stride = np.gcd(stride, d)
    stride = int(stride) if nz else 1
    n1 = 1
    if double_hoist:
        while n1 < BSGS_MAX_BABY and 2 * n1 < len(rots):
            n1 *= 2
    else:
        while n1 * n1 < len(rots):
            n1 *= 2
    bm = stride * n1
    plan: Dict[int, List[Tuple[int, int]]] = {}
    for d in rots:
        b = d % bm
        plan.setdefault((d - b) % n, []).append((b, d))
    return bm, plan


def _evalmod_design(rho: float, basis: str = "monomial", r: int = None, degree: int = None, even: bool = False):
    """coefficients of alpha_0 cos(2 pi (K_n y - 1/4) / 2^r) on y in [-1, 1] (monomial or
    Chebyshev basis) and the constants alpha_1..alpha_r of the double-angle steps
    c <- c^2 - alpha.  The final amplitude is rho / 2 pi.

    even: the cosine is an even function of u = y - 1 / (4 K_n) (|u| <= 32.5 / 33 < 1), so its Chebyshev series in u
    has even terms only, T_2k(u) = T_k(w) with w = 2 u^2 - 1: returns the degree / 2 coefficients in w."""
    r = DOUBLE_ANGLES if r is None else r
    degree = POLY_DEGREE if degree is None else degree
    A = 2 * np.pi * K_NORM / 2 ** r
    phi = 2 * np.pi * 0.25 / 2 ** r
    alphas = [rho / (2 * np.pi)]
    for _ in range(r):
        alphas.append(np.sqrt(2 * alphas[-1]))
    alphas = alphas[::-1]                       # alphas[0] scales the base polynomial
    if even:
        if basis != "chebyshev":
            raise ValueError("the even form is evaluated in the Chebyshev basis")
        return _cheb.chebinterpolate(lambda w: alphas[0] * np.cos(A * np.sqrt(np.maximum(w + 1.0, 0.0) / 2.0)), degree // 2), alphas
    cheb = _cheb.chebinterpolate(lambda y: alphas[0] * np.cos(A * y - phi), degree)
    return (cheb if basis == "chebyshev" else _cheb.cheb2poly(cheb)), alphas


def chebyshev_basis(engine, relin_key, y: Ciphertext, degree: int) -> Dict[int, Ciphertext]:
    """{k: T_k(y)}, k = 1..degree, by T_{a+b} = 2 T_a T_b - T_{a-b}; T_k sits ceil(log2 k) levels
    below y.  Bounded basis (|T_k| <= 1): the interpolant's coefficients decay instead of
    alternating with magnitude ~10 as in the monomial basis, so ciphertext noise is not amplified
    by cancellation."""
    T: Dict[int, Ciphertext] = {1: y}
    for k in range(2, degree + 1):
        a = (k + 1) // 2
        b = k // 2
        prod = engine.multiply(T[a], T[b], relin_key)
        twice = engine.add(prod, prod)
        if a == b:
            T[k] = engine.add_plain(twice, -1.0)
        else:
            T[k] = engine.subtract(twice, T[a - b])
    return T


def chebyshev_eval_ps(engine, relin_key, y: Ciphertext, coeffs, baby: int = 4, even: bool = False) -> Ciphertext:
    """sum_k coeffs[k] T_k(y) (even: sum_k coeffs[k] T_k(2 y^2 - 1), i.e. an even polynomial of twice the degree in y
    for one extra product) with the Paterson-Stockmeyer recursion in the Chebyshev basis:
    p = q * T_n + r with T_k = 2 T_n T_(k-n) - T_(2n-k) for n < k < 2n, n = baby * 2^j.  The leaves
    (degree < baby) are constant-only linear combinations of T_1..T_(baby-1).  Degree 22 costs
    10 ciphertext products (T_2, T_3, T_4, T_8, T_16 and five recombinations) instead of the 21 of
    the full basis, at the same depth (6)."""
    c = [float(v) for v in coeffs]
    while len(c) > 1 and abs(c[-1]) < 1e-300:
        c.pop()
    deg = len(c) - 1

    def double_minus(prod, sub):          # 2 * prod - sub  (sub: ciphertext or the constant 1)
        twice = engine.add(prod, prod)
        return engine.add_plain(twice, -1.0) if sub is None else engine.subtract(twice, sub)

    if even:
        y = double_minus(engine.multiply(y, y, relin_key), None)
    T: Dict[int, Ciphertext] = {1: y}

    for k in range(2, baby + 1):
        a, b = (k + 1) // 2, k // 2
        T[k] = double_minus(engine.multiply(T[a], T[b], relin_key), None if a == b else T[a - b])
    n = baby
    while 2 * n <= deg:
        T[2 * n] = double_minus(engine.multiply(T[n], T[n], relin_key), None)
        n *= 2

    def leaf(cc):
        terms = {k: cc[k] for k in range(1, len(cc)) if abs(cc[k]) > 0.0}
        if not terms:
            return float(cc[0])
        return lincomb_const(engine, {k: T[k] for k in terms}, terms, const=cc[0])

    def ev(cc):
        d = len(cc) - 1
        if d < baby:
            return leaf(cc)
        n = baby
        while 2 * n <= d:
            n *= 2
        q = [0.0] * (d - n + 1)
        r = list(cc[:n])
        q[0] = cc[n]
        for k in range(n + 1, d + 1):
            q[k - n] += 2.0 * cc[k]
            r[2 * n - k] -= cc[k]
        qv, rv = ev(q), ev(r)
        prod = engine.multiply(T[n], qv) if isinstance(qv, float) else engine.multiply(qv, T[n], relin_key)
        return engine.add_plain(prod, rv) if isinstance(rv, float) else engine.add(prod, rv)

    out = ev(c)
    if isinstance(out, float):
        raise ValueError("chebyshev_eval_ps: constant polynomial")
    return out


def chebyshev_eval_ps_lazy(engine, relin_key, y: Ciphertext, coeffs, baby: int = 4, even: bool = False) -> Ciphertext:
    """The same polynomial and the same products as chebyshev_eval_ps (Paterson-Stockmeyer in the Chebyshev basis,
    baby = 4; even: in w = 2 y^2 - 1) without a single level adjustment.  Every ciphertext carries its exact scale as a rational factor
    `dev` on the scale table (true scale = delta[level] * dev):

    * a product of operands at different levels uses the higher one in place (Engine._mul_ct_dropped: upper limbs
      ignored, an exact modulus switch) -- the product's scale leaves the table by a known factor;
    * T_2n = 2 T_n^2 - 1 costs no doubling: the product T_n^2 at scale s IS the value 2 T_n^2 at scale s / 2, then
      the constant -1 is added at that scale;
    * the leaves use {T_1, T_2, T_1 T_2} (T_3 = 2 T_1 T_2 - T_1 folded into the constants), one constant-only
      pass each, and a leaf can be produced at ANY scale;
    * so a node  q T_n + r  first fixes the scale S it must deliver, asks q for S q_l / scale(T_n) and r for S:
      the sum needs no alignment, and the root is asked for the table scale.

    Same depth (5 levels for degree <= 23) and 9 products for degree 18, but none of the 19 constant-multiply +
    rescale alignments per evaluation that the eager form spends on mixed-level adds and products."""
    from fractions import Fraction
    from .fused import _const_pair
    be, P = engine.backend, engine.params
    c = [float(v) for v in coeffs]
    while len(c) > 1 and abs(c[-1]) < 1e-300:
        c.pop()
    deg = len(c) - 1
    if baby != 4 or deg < baby:
        return chebyshev_eval_ps(engine, relin_key, y, coeffs, baby, even)

    def scale_of(ct, dev):
        return P.delta[ct.level] * dev

    def mul(a, da, b, db):
        out, d = engine._mul_ct_dropped(a, b, relin_key)
        return out, (d * da * db).limit_denominator(1 << 600)

    def two_x_minus_one(a, da):
        """2 a^2 - 1 from a: the square read as twice its value at half the scale"""
        sq, d = mul(a, da, a, da)
        d = d / 2
        return engine.add_plain(sq, -float(d), inplace=True), d

    one = Fraction(1)
    T: Dict[int, Tuple[Ciphertext, Fraction]] = {1: two_x_minus_one(y, one) if even else (y, one)}
    T[2] = two_x_minus_one(*T[1])
    p12 = mul(*T[1], *T[2])                                    # T_1 T_2 = (T_3 + T_1) / 2
    T[4] = two_x_minus_one(*T[2])
    n = 4
    while 2 * n <= deg:
        T[2 * n] = two_x_minus_one(*T[n])
        n *= 2
    leaf_in = {1: T[1], 2: T[2], 3: p12}                       # 3 = the product T_1 T_2

    def leaf_terms(cc):
        cc = list(cc) + [0.0] * (4 - len(cc))
        terms = {1: cc[1] - cc[3], 2: cc[2], 3: 2.0 * cc[3]}
        return {k: v for k, v in terms.items() if abs(v) > 0.0}, cc[0]

    def leaf_level(cc):
        terms, _ = leaf_terms(cc)
        return min(leaf_in[k][0].level for k in terms) - 1 if terms else None

    def plan_level(cc):
        d = len(cc) - 1
        if d < baby:
            return leaf_level(cc)
        n = baby
        while 2 * n <= d:
            n *= 2
        q, _ = split(cc, n)
        lq = plan_level(q)
        return min(lq if lq is not None else T[n][0].level, T[n][0].level) - 1

    def split(cc, n):
        d = len(cc) - 1
        q = [0.0] * (d - n + 1)
        r = list(cc[:n])
        q[0] = cc[n]
        for k in range(n + 1, d + 1):
            q[k - n] += 2.0 * cc[k]
            r[2 * n - k] -= cc[k]
        return q, r

    def leaf(cc, S):
        """the leaf polynomial at true scale S (default: the table scale of its level), or a float"""
        terms, c0 = leaf_terms(cc)
        if not terms:
            return float(c0), None
        lo = min(leaf_in[k][0].level for k in terms)
        nq = lo + 1
        if S is None:
            S = P.delta[lo - 1]
        target = S * P.moduli[lo]                               # scale in front of the rescale
        keys = sorted(terms)
        res = [[_const_pair(engine, complex(terms[k]), target / scale_of(*leaf_in[k]), nq) for k in keys]]
        c0r = [_const_pair(engine, complex(c0), target, nq)]
        out = be.lincomb([leaf_in[k][0].polys for k in keys], be.prepare_lincomb(res, c0r, nq))[0]
        ct = engine._rescale(Ciphertext(engine, out, lo))
        return ct, (S / P.delta[ct.level]).limit_denominator(1 << 600)

    def ev(cc, S=None):
        d = len(cc) - 1
        if d < baby:
            return leaf(cc, S)
        n = baby
        while 2 * n <= d:
            n *= 2
        q, r = split(cc, n)
        tn, dn = T[n]
        lq = plan_level(q)
        if lq is None:
            # constant quotient: q[0] * T_n as a constant multiple, delivered at scale S one level below T_n
            lvl = tn.level
            if S is None:
                S = P.delta[lvl - 1]
            k = S * P.moduli[lvl] / scale_of(tn, dn) * Fraction(float(q[0]))
            src = tn
            prod = engine._rescale(engine._mul_int_const(src, int(round(k)), 0))
        else:
            lvl = min(lq, tn.level)
            if S is None:
                S = P.delta[lvl - 1]
            qv, dq = ev(q, S * P.moduli[lvl] / scale_of(tn, dn))
            prod, dp = mul(qv, dq, tn, dn)
            assert prod.level == lvl - 1
        rv, dr = ev(r, S)
        if isinstance(rv, float):
            return engine.add_plain(prod, rv * float(S / P.delta[prod.level])), (S / P.delta[prod.level]).limit_denominator(1 << 600)
        nq = prod.level + 1
        rp = rv.polys if rv.level == prod.level else be.take_limbs(rv.polys, nq, False)
        engine._count('add_ct')
        return Ciphertext(engine, be.add(prod.polys, rp, nq, 0), prod.level), (S / P.delta[prod.level]).limit_denominator(1 << 600)

    out, dev = ev(c, None)
    assert dev == 1, dev
    return out


def lincomb_const(engine, cts: Dict[int, Ciphertext], coeffs, const: float = 0.0) -> Ciphertext:
    """const + sum_k coeffs[k] * cts[k]: constants absorb the level / scale alignment, one pass of
    fhe_lincomb per 16 inputs, ONE rescale."""
    from .fused import _const_pair
    be, P = engine.backend, engine.params
    lo = min(c.level for c in cts.values())
    nq = lo + 1
    target = P.delta[lo - 1] * P.moduli[lo]
    keys = sorted(cts)
    acc = None
    for s in range(0, len(keys), 16):
        part = keys[s:s + 16]
        res = [[_const_pair(engine, complex(coeffs[k]), target / P.delta[cts[k].level], nq) for k in part]]
        c0 = [_const_pair(engine, complex(const), target, nq)] if s == 0 else None
        out = be.lincomb([cts[k].polys for k in part], be.prepare_lincomb(res, c0, nq))[0]
        acc = out if acc is None else be.add(acc, out, nq, 0)
    return engine._rescale(Ciphertext(engine, acc, lo))


def _ps_depth(degree: int, baby: int = 4) -> int:
    """levels chebyshev_eval_ps uses for a dense polynomial of this degree (mirrors its recursion: a leaf is one
    constant-only sum below T_(baby-1), a node is one product with T_n below its quotient)"""
    def depth(d):
        if d < baby:
            return (max(d, 1) - 1).bit_length() + 1 if d >= 1 else 0
        n = baby
        while 2 * n <= d:
            n *= 2
        return max(max(depth(d - n), (n - 1).bit_length()) + 1, depth(n - 1))
    return depth(degree)


# --------------------------------------------------------------------------- key
class _Plan:
    pass


def make_bootstrap_key(engine, sk, groups: int = 3, groups_stc: Optional[int] = None) -> BootstrapKey:
    """Lazy: the (large) set of Galois keys and encoded matrices is built on first use.  The
    key object keeps a reference to the secret key for that purpose, like the reference's
    EngineContext keeps every key in one process (engine_context.py:62-73)."""
    bk = BootstrapKey(small=False)
    bk._sk = sk
    bk._groups = groups                                            # CoeffToSlot matrices
    bk._groups_stc = groups if groups_stc is None else groups_stc   # SlotToCoeff matrices (fewer = one level back)
    return bk


def _materialise(engine, bk: BootstrapKey):
    if bk.plan is not None:
        return bk.plan
    if engine.secret_hamming_weight <= 0:
        raise RuntimeError("bootstrapping needs the sparse secret: construct the engine with use_bootstrap=True "
                           "(a uniform ternary secret makes the ModRaise overflow |I| ~ sqrt(N), not <= 32)")
    P = engine.params
    n = engine.slot_count
    L = P.max_level
    groups = bk._groups
    groups_stc = getattr(bk, "_groups_stc", groups)
    depth = 1 + groups + 5 + DOUBLE_ANGLES + groups_stc    # extra rescale of the first matrix, EvalMod polynomial (5), the rest
    depth_bits = groups + _ps_depth(POLY_DEGREE_BITS) + DOUBLE_ANGLES_BITS
    if L < min(depth, depth_bits) + 1:
        raise RuntimeError(f"bootstrapping needs max_level >= {min(depth, depth_bits) + 1}, engine has {L}")
    layers, inv_layers = _fft_layers(n)
    # CoeffToSlot: L_1^-1 ... L_last^-1 with L_last^-1 applied first; fold 1 / (2 K_n)
    cts = _group(list(reversed(inv_layers)), groups, n)
    cts[0] = {d: v / (2.0 * K_NORM) for d, v in cts[0].items()}
    stc = _group(layers, groups_stc, n)
    plan = _Plan()
    plan.depth = depth
    plan.rot_keys: Dict[int, FixedRotationKey] = {}
    plan.cts, plan.stc = [], []

    dh = bool(getattr(bk, "_double_hoist", not _NO_DH))

    def prepare(mat, scale_fn=None):
        return prepare_matrix(engine, bk._sk, mat, plan.rot_keys, scale_fn, dh and scale_fn is not first_scale)

    # first matrix: two rescales, plaintext scale ~ Delta * q / q_0 * q: rounding of the diagonals
    # must be far below 2^-40 because the raised ciphertext's slots are ~ sqrt(n) * |I| large
    first_scale = lambda lvl: P.delta[lvl - 2] * P.moduli[lvl] * P.moduli[lvl - 1] / P.moduli[0]      # noqa: E731
    for gi, m in enumerate(cts):
        plan.cts.append(prepare(m, first_scale if gi == 0 else None))
    # bit bootstrap: everything in front of EvalMod comes out squared, so its first matrix needs no extra precision:
    # one rescale (plaintext scale Delta q / q_0), one level saved
    plan.cts_bits0 = prepare(cts[0], lambda lvl: P.delta[lvl - 1] * P.moduli[lvl] / P.moduli[0])
    for m in stc:
        plan.stc.append(prepare(m))
    plan.shift = max(1.0, RHO_TARGET * float(P.delta[0]) / P.moduli[0])        # message divisor
    plan.rho = plan.shift * float(P.moduli[0] / P.delta[0])
    plan.poly, plan.alphas = _evalmod_design(plan.rho, "chebyshev")
    # bit bootstrap: message * q_0 / 4 at level 0, unit amplitude out
    plan.shift_bits = 4.0 * float(P.delta[0]) / P.moduli[0]
    plan.poly_bits, plan.alphas_bits = _evalmod_design(2.0 * np.pi, "chebyshev", DOUBLE_ANGLES_BITS, POLY_DEGREE_BITS)
    plan.poly_bits_even, _ = _evalmod_design(2.0 * np.pi, "chebyshev", DOUBLE_ANGLES_BITS, POLY_DEGREE_BITS, even=True)
    plan.depth_bits = groups + _ps_depth(POLY_DEGREE_BITS) + DOUBLE_ANGLES_BITS          # levels above the ModRaise
    bk.plan = plan
    return plan


def prepare_matrix(engine, sk, mat: Dict[int, np.ndarray], rot_keys: Dict[int, FixedRotationKey], scale_fn=None,
                   double_hoist: bool = False):
    """A slot-domain matrix in diagonal form ({rotation d: diagonal}, out[p] = sum_d diag_d[p] in[(p + d) % n]) made
    ready for _linear_transform: baby-step/giant-step split, pre-rotated plaintext diagonals, Galois keys (created
    with `sk`, or fetched from a ReceivedKeys object, into `rot_keys`)."""
    n = engine.slot_count
    bm, split = _bsgs_split(sorted(mat), n, double_hoist)
    entry = dict(bm=bm, giants={}, dh=bool(double_hoist), _keys=rot_keys)
    for g, items in split.items():
        lst = []
        for b, d in items:
            # inner sum is rotated by g afterwards, so the diagonal is pre-rotated the other way
            lst.append((b, Plaintext(engine, np.roll(mat[d], g), scale_fn=scale_fn)))
            if b and b not in rot_keys:
                rot_keys[b] = engine.create_fixed_rotation_key(sk, -b)
        entry["giants"][g] = lst
        if g and g not in rot_keys:
            rot_keys[g] = engine.create_fixed_rotation_key(sk, -g)
    return entry


# --------------------------------------------------------------------------- evaluation
def _linear_transform_dh(engine, ct: Ciphertext, entry) -> Ciphertext:
    """The same sum with DOUBLE hoisting (Bossuat-Mouchet-Troncoso-Pastoriza-Hubaux 2021): the baby-step rotations
    share one ModUp AND skip their ModDowns -- the rotated ciphertexts stay in the extended basis Q u P, are
    multiplied there with extended-basis plaintexts and summed; only the G giant-step sums are brought down (one
    ModDown each), rotated (one ModUp each) and accumulated, with one final ModDown merged with the rescale.
    1 + (G-1) ModUps and G ModDowns instead of n1 + n2 each.  On the B200 the baby side is ONE kernel
    (fhe_bsgs_inner): inner products with all baby keys, the lift of c0, the automorphisms (as gathers) and the
    plaintext products, without any intermediate in memory."""
    be, P = engine.backend, engine.params
    keys = entry["_keys"]
    lvl = ct.level
    nq, K = lvl + 1, P.n_p
    giants = entry["giants"]
    glist = list(giants)
    blist = sorted({b for items in giants.values() for b, _ in items})
    rows = [[dict(items).get(b) for b in blist] for items in giants.values()]

    def pt(p):
        return None if p is None else p.at_level(lvl, ext=True)

    c0, c1 = be.select_poly(ct.polys, 0), be.select_poly(ct.polys, 1)
    engine._count('keyswitch_galois', sum(1 for b in blist if b))
    engine._count('mul_pt', sum(1 for row in rows for p in row if p is not None))
    if hasattr(be, "bsgs_inner") and not _NO_DH_FUSE and P.digits_at(nq) <= 4:
        ext = be.modup_raw(c1, nq)
        inners = []
        for s in range(0, len(glist), 4):
            chunk = rows[s:s + 4]
            out = None
            for b0 in range(0, len(blist), 16):                # the kernel takes 16 baby steps per pass
                bl = blist[b0:b0 + 16]
                out = be.bsgs_inner(ext, ct.polys, [None if b == 0 else keys[b].data for b in bl],
                                    [1 if b == 0 else int(keys[b].galois) for b in bl],
                                    [[pt(p) for p in row[b0:b0 + 16]] for row in chunk], nq, acc=out)
            inners += [out[i] for i in range(len(chunk))]
    else:
        # the same values from the primitives both backends have (the oracle's path):
        #   Z_b = sigma_b( <ModUp(c1), sigma_b^-1 key_b> + P (c0, 0) )
        ext = be.modup(c1, nq)
        two_n = 2 * P.n
        Z = {}
        for b in blist:
            if b == 0:
                Z[b] = be.ks_accum(None, None, None, ct.polys, nq)
                continue
            key = keys[b]
            pre = getattr(key, "_pre_permuted", None)
            if pre is None:
                kd = key.data
                flat = be.automorphism(kd.reshape((-1, 1) + tuple(kd.shape[-2:])), pow(int(key.galois), -1, two_n), kd.shape[-2], 0)
                pre = key._pre_permuted = flat.reshape(kd.shape)
            acc = be.ks_accum(be.ks_inner(ext, c1, pre, nq), None, None, c0, nq)
            Z[b] = be.automorphism(acc, key.galois, nq, K)
        inners = []
        for row in rows:
            acc = None
            for b, p in zip(blist, row):
                if p is None:
                    continue
                term = be.mul(Z[b], pt(p), nq, K)
                acc = term if acc is None else be.add(acc, term, nq, K)
            inners.append(acc)
    acc = None
    for g, inner in zip(glist, inners):
        if g == 0:
            acc = inner if acc is None else be.add(acc, inner, nq, K)
            continue
        key = keys[g]
        low = be.moddown_inplace(inner, nq) if hasattr(be, "moddown_inplace") else be.moddown(inner, nq)
        rot = be.automorphism(low, key.galois, nq, 0)
        acc = be.ks_accum(acc, be.select_poly(rot, 1), key.data, be.select_poly(rot, 0), nq)
        engine._count('keyswitch_galois')
    engine._count('rescale')
    return Ciphertext(engine, be.moddown_rescale(acc, nq), lvl - 1)


def _linear_transform(engine, ct: Ciphertext, entry) -> Ciphertext:
    """sum_d diag_d (.) roll(x, -d) with baby-step / giant-step rotations; one level."""
    if entry.get("dh") and hasattr(engine.backend, "ks_accum"):
        return _linear_transform_dh(engine, ct, entry)
    plan_keys = entry["_keys"]
    babies: Dict[int, Ciphertext] = {0: ct}
    need = sorted({b for items in entry["giants"].values() for b, _ in items if b})
    for b, r in zip(need, engine.rotate_hoisted(ct, [plan_keys[b] for b in need])):      # one ModUp for all baby steps
        babies[b] = r
    be = engine.backend
    shared = hasattr(be, "ks_accum")          # giant-step key switches accumulate in the extended basis: ONE ModDown
    out, acc, lvl = None, None, None
    for items in entry["giants"].values():
        for b, _ in items:
            if b not in babies:
                babies[b] = engine.rotate(ct, plan_keys[b])
    # all diagonal sums in one pass over the baby rotations (each read once, not once per giant step); they stay
    # un-rescaled (scale delta * q_level) through the giant-step rotation and the final sum: ONE rescale per
    # transform instead of one per giant step
    blist = sorted(babies)
    rows = []
    for items in entry["giants"].values():
        have = dict(items)
        rows.append([have.get(b) for b in blist])
    inners = engine.multiply_plain_sums([babies[b] for b in blist], rows)
    for g, inner in zip(entry["giants"], inners):
        if shared:
            # ... and the giant rotations share their ModDown: rot_g(x) = ModDown(<ModUp(sigma x1), key_g> + P sigma x0),
            # ModDown is linear up to rounding, so the extended accumulators are summed (inside the inner-product
            # kernel) and divided by P q_level once -- G - 1 ModDowns and the separate rescale disappear
            lvl = inner.level
            nq = lvl + 1
            if g:
                key = plan_keys[g]
                rot = be.automorphism(inner.polys, key.galois, nq, 0)
                acc = be.ks_accum(acc, be.select_poly(rot, 1), key.data, be.select_poly(rot, 0), nq)
                engine._count('keyswitch_galois')
            else:
                acc = be.ks_accum(acc, None, None, inner.polys, nq)
            continue
        if g:
            inner = engine.rotate(inner, plan_keys[g])
        out = inner if out is None else engine.add(out, inner)
    if shared:
        engine._count('rescale')
        return Ciphertext(engine, be.moddown_rescale(acc, lvl + 1), lvl - 1)
    return engine._rescale(out)


def bootstrap(engine, ct: Ciphertext, relin_key, conj_key, boot_key: BootstrapKey) -> Ciphertext:
    if getattr(boot_key, "small", False) or not hasattr(boot_key, "_sk"):
        raise RuntimeError("bootstrap needs the key from create_bootstrap_key")
    if ct.npoly != 2:
        raise RuntimeError("bootstrap: ciphertext must have 2 polynomials")
    plan = _materialise(engine, boot_key)
    for e in plan.cts + plan.stc:
        e["_keys"] = plan.rot_keys
    be, P = engine.backend, engine.params
    L = P.max_level
    engine._count("bootstrap")

    # 0. level 0, message / 2^5
    if ct.level == 0:
        raise RuntimeError("bootstrap: call before the ciphertext reaches level 0")
    x = ct if ct.level == 1 else Ciphertext(engine, be.take_limbs(ct.polys, 2, False), 1)
    from fractions import Fraction
    c = P.delta[0] * P.moduli[1] / (P.delta[ct.level] * Fraction(plan.shift))
    x = engine._rescale(engine._mul_int_const(x, int(round(c)), 0))                  # level 0

    # 1. ModRaise: declared scale q_0, values = eps * msg_coeff + I
    raised = Ciphertext(engine, be.mod_raise(x.polys, L + 1), L)

    # 2. CoeffToSlot (first matrix absorbs the q_0 scale)
    t = engine._rescale(_linear_transform(engine, raised, plan.cts[0]))     # second rescale: scale is standard from here
    for entry in plan.cts[1:]:
        t = _linear_transform(engine, t, entry)

    # 3. real and imaginary parts, stacked on the batch axis
    tc = engine.conjugate(t, conj_key)
    re = engine.add(t, tc)
    im = engine.multiply_by_i(engine.subtract(t, tc), -1)
    bt = t.batch
    y = Ciphertext(engine, be.concat_batch([re.polys, im.polys]), t.level)

    # 4. EvalMod
    cpoly = chebyshev_eval_ps(engine, relin_key, y, plan.poly)
    for i in range(DOUBLE_ANGLES):
        cpoly = engine.add_plain(engine.multiply(cpoly, cpoly, relin_key), -plan.alphas[i + 1])
    re_p, im_p = be.split_batch(cpoly.polys, [bt, bt])
    u = engine.add(Ciphertext(engine, re_p, cpoly.level),
                   engine.multiply_by_i(Ciphertext(engine, im_p, cpoly.level), 1))

    # 5. SlotToCoeff
    for entry in plan.stc:
        u = _linear_transform(engine, u, entry)
    return u


def _phase(engine, name: str):
    """optional hook for profiling tools (tools/aes_bits_probe.py): engine.phase_timer(name) at phase boundaries"""
    hook = getattr(engine, "phase_timer", None)
    if hook is not None:
        hook(name)


def bootstrap_bits(engine, ct: Ciphertext, relin_key, conj_key, boot_key: BootstrapKey,
                   top_level: Optional[int] = None) -> Ciphertext:
    """Refresh of +-1-valued slots ("bit bootstrap").  ``ct`` [batch B] holds u + i v with u, v real and close to
    +-1 (two bit planes per ciphertext); the result [batch 2 B: all u, then all v] holds the cleaned values at
    level  top_level - plan.depth_bits  (top_level = max_level unless the caller needs fewer levels afterwards: the
    ModRaise then goes to top_level + 1 limbs only and every step of the refresh runs on fewer limbs).

    Order of operations (binary-message bootstrapping, Bae-Cheon-Kim-Stehle 2024, with this repo's pieces):
      0. SlotToCoeff FIRST, at the bottom of the chain (levels groups_stc + 1 -> 1, a few limbs: almost free); the
         plaintext polynomial then has the coefficients Delta u_k and Delta v_k themselves
      1. down to level 0 with the message scaled to q_0 / 4, ModRaise: coefficient = q_0 (I + s / 4) + e
      2. CoeffToSlot, real / imaginary split by one conjugation (stacked on the batch axis)
      3. EvalMod = sin(2 pi x) with unit amplitude: at x = I +- 1/4 the derivative vanishes, so an input error e
         comes out as pi^2 e^2 / 8 -- the refresh is also the clean-up -- and the result is already slot-encoded:
         no SlotToCoeff at the top, no separate clean-up polynomial."""
    if getattr(boot_key, "small", False) or not hasattr(boot_key, "_sk"):
        raise RuntimeError("bootstrap needs the key from create_bootstrap_key")
    if ct.npoly != 2:
        raise RuntimeError("bootstrap: ciphertext must have 2 polynomials")
    from fractions import Fraction
    plan = _materialise(engine, boot_key)
    for e in plan.cts + plan.stc + [plan.cts_bits0]:
        e["_keys"] = plan.rot_keys
    be, P = engine.backend, engine.params
    L = P.max_level if top_level is None else int(top_level)
    if L > P.max_level:
        raise RuntimeError(f"bootstrap_bits: top_level {L} above max_level {P.max_level}")
    if L < plan.depth_bits + 1:
        raise RuntimeError(f"bit bootstrapping needs {plan.depth_bits + 1} levels at the top, got {L}")
    engine._count("bootstrap")
    lvl_in = len(plan.stc) + 1
    if ct.level < lvl_in:
        raise RuntimeError(f"bootstrap_bits: {ct.level} levels left, SlotToCoeff and the step to level 0 need {lvl_in}")

    # 0. SlotToCoeff on the lowest limbs.  Dropping limbs is an exact modulus switch: same message, the scale stays
    # delta[ct.level]; the factor is carried to step 1
    dev = P.delta[ct.level] / P.delta[lvl_in]
    u = ct if ct.level == lvl_in else Ciphertext(engine, be.take_limbs(ct.polys, lvl_in + 1, False), lvl_in)
    _phase(engine, "boot:entry")
    for entry in plan.stc:
        u = _linear_transform(engine, u, entry)                                     # ends at level 1
    _phase(engine, "boot:slot_to_coeff")

    # 1. level 0, plaintext coefficient = s * q_0 / 4; ModRaise (declared scale q_0: values I + s / 4)
    c = P.delta[0] * P.moduli[1] / (P.delta[1] * dev * Fraction(plan.shift_bits))
    x = engine._rescale(engine._mul_int_const(u, int(round(c)), 0))
    raised = Ciphertext(engine, be.mod_raise(x.polys, L + 1), L)
    _phase(engine, "boot:mod_raise")

    # 2. CoeffToSlot; y = (I + s / 4) / K_n for the first and the second half of the coefficients
    t = _linear_transform(engine, raised, plan.cts_bits0)
    for entry in plan.cts[1:]:
        t = _linear_transform(engine, t, entry)
    _phase(engine, "boot:coeff_to_slot")
    tc = engine.conjugate(t, conj_key)
    re = engine.add(t, tc)
    im = engine.multiply_by_i(engine.subtract(t, tc), -1)
    y = Ciphertext(engine, be.concat_batch([re.polys, im.polys]), t.level)

    _phase(engine, "boot:conjugate_split")

    # 3. EvalMod with unit amplitude.  The cosine is even in u = y - 1 / (4 K_n): a degree-9 polynomial in
    # w = 2 u^2 - 1 instead of degree 18 in y -- 7 products instead of 9 at the same depth (1 + 4 levels)
    ps = chebyshev_eval_ps_lazy if hasattr(engine, "_mul_ct_dropped") and not _EAGER_PS else chebyshev_eval_ps
    if _EVEN_EVALMOD:
        cpoly = ps(engine, relin_key, engine.add_plain(y, -0.25 / K_NORM, inplace=True), plan.poly_bits_even, even=True)
    else:
        cpoly = ps(engine, relin_key, y, plan.poly_bits)
    _phase(engine, "boot:evalmod_polynomial")
    for i in range(DOUBLE_ANGLES_BITS):
        cpoly = engine.add_plain(engine.multiply(cpoly, cpoly, relin_key), -plan.alphas_bits[i + 1], inplace=True)
    _phase(engine, "boot:evalmod_double_angle")
    return cpoly
