"""Shared bit-exact parity checks: a CUDA-source backend (the real GPU build, or the host
simulation of the same kernels) against the CPU oracle, through the C ABI."""
import numpy as np

from conftest import rand_poly


def check_primitives(P, gb, rb, seed=1):
    rng = np.random.default_rng(seed)
    nq, K = P.n_q, P.n_p
    up = gb.from_numpy
    a = rand_poly(P, rng, 2, nq, True)
    assert np.array_equal(gb.to_numpy(gb.ntt(up(a), nq, K)), rb.ntt(a, nq, K)), "ntt"
    assert np.array_equal(gb.to_numpy(gb.intt(up(a), nq, K)), rb.intt(a, nq, K)), "intt"
    assert np.array_equal(gb.to_numpy(gb.intt(gb.ntt(up(a), nq, K), nq, K)), a), "ntt round trip"
    b = rand_poly(P, rng, 2, nq, False, batch=2)
    c = rand_poly(P, rng, 2, nq, False, batch=2)
    c1 = rand_poly(P, rng, 1, nq, False, batch=2)
    c11 = rand_poly(P, rng, 1, nq, False, batch=1)
    c21 = rand_poly(P, rng, 2, nq, False, batch=1)
    for name in ("add", "sub", "mul"):
        assert np.array_equal(gb.to_numpy(getattr(gb, name)(up(b), up(c), nq, 0)), getattr(rb, name)(b, c, nq, 0)), name
        for other, tag in ((c1, "poly"), (c11, "poly+batch"), (c21, "batch")):
            assert np.array_equal(gb.to_numpy(getattr(gb, name)(up(b), up(other), nq, 0)),
                                  getattr(rb, name)(b, other, nq, 0)), f"{name} broadcast {tag}"
    assert np.array_equal(gb.to_numpy(gb.mul(up(c21), up(c1), nq, 0)), rb.mul(c21, c1, nq, 0)), "mul cross broadcast"
    assert np.array_equal(gb.to_numpy(gb.sub(up(c11), up(b), nq, 0)), rb.sub(c11, b, nq, 0)), "sub small - big"
    assert np.array_equal(gb.to_numpy(gb.neg(up(b), nq, 0)), rb.neg(b, nq, 0)), "neg"
    assert np.array_equal(gb.to_numpy(gb.tensor(up(b), up(c), nq)), rb.tensor(b, c, nq)), "tensor"
    for g in (5, 2 * P.n - 1, pow(5, 7, 2 * P.n), pow(5, P.slot_count - 3, 2 * P.n)):
        assert np.array_equal(gb.to_numpy(gb.automorphism(up(b), g, nq, 0)), rb.automorphism(b, g, nq, 0)), f"auto {g}"
    co = rng.integers(-2 ** 55, 2 ** 55, size=(2, P.n), dtype=np.int64)
    co[0, :4] = [0, -1, 1, -(2 ** 62)]
    assert np.array_equal(gb.to_numpy(gb.from_i64(co, nq, True)), rb.from_i64(co, nq, True)), "from_i64"
    cp = [int(rng.integers(0, P.moduli[l])) for l in range(nq)]
    cm = [int(rng.integers(0, P.moduli[l])) for l in range(nq)]
    assert np.array_equal(gb.to_numpy(gb.mul_const(up(b), cp, cm, nq)), rb.mul_const(b, cp, cm, nq)), "mul_const"
    assert np.array_equal(gb.to_numpy(gb.add_const(up(b), cp, cm, nq)), rb.add_const(b, cp, cm, nq)), "add_const"
    assert np.array_equal(gb.to_numpy(gb.add_poly0(up(b), up(c1), nq)), rb.add_poly0(b, c1, nq)), "add_poly0"
    if hasattr(gb, "mul_plain_multi"):
        # diagonal sums of several giant steps in one pass: out_g = sum_t a_t (.) p[g][t], absent terms skipped
        # (T, G) = (3, 2): eight rows per sweep; (6, 5): four per sweep, ragged last sweep; (9, 3): two per sweep
        for T_, G_ in ((3, 2), (6, 5), (9, 3)):
            al = [rand_poly(P, rng, 2, nq, False, batch=2) for _ in range(T_)]
            pl = [[rand_poly(P, rng, 1, nq, False, batch=1) for _ in range(T_)] for _ in range(G_)]
            pl[1][1] = None
            pl[G_ - 1][0] = None
            got = gb.to_numpy(gb.mul_plain_multi([up(a_) for a_ in al],
                                                 [[None if p_ is None else up(p_) for p_ in row] for row in pl], nq))
            for g_, row in enumerate(pl):
                want = None
                for a_, p_ in zip(al, row):
                    if p_ is not None:
                        term = rb.mul(a_, p_, nq, 0)
                        want = term if want is None else rb.add(want, term, nq, 0)
                assert np.array_equal(got[g_], want), f"mul_plain_multi T={T_} G={G_} row {g_}"
    # constant-weighted sums straight through fhe_lincomb: one / two outputs (plain FP64 kernel) and more (tensor cores)
    for M_, T_ in ((1, 3), (2, 5), (3, 2), (9, 13)):
        ins = [rand_poly(P, rng, 2, nq, False, batch=2) for _ in range(T_)]
        rc = lambda: [int(rng.integers(0, P.moduli[l])) for l in range(nq)]
        cres = [[(rc(), rc()) for _ in range(T_)] for _ in range(M_)]
        c0 = [(rc(), rc()) for _ in range(M_)] if M_ != 2 else None
        got = gb.lincomb([up(a_) for a_ in ins], gb.prepare_lincomb(cres, c0, nq))
        want = rb.lincomb(ins, rb.prepare_lincomb(cres, c0, nq))
        for m_ in range(M_):
            assert np.array_equal(gb.to_numpy(got[m_]), want[m_]), f"lincomb M={M_} T={T_} output {m_}"
    x = rand_poly(P, rng, 1, 2, False, batch=3)
    assert np.array_equal(gb.crt_centered(up(x), 2), rb.crt_centered(x, 2)), "crt2"
    x1 = np.ascontiguousarray(x[:, :, :1])
    assert np.array_equal(gb.crt_centered(up(x1), 1), rb.crt_centered(x1, 1)), "crt1"


def check_rescale(P, gb, rb, seed=2):
    rng = np.random.default_rng(seed)
    for nq in sorted({P.n_q, max(2, P.n_q - 3), 2}):
        for npoly, batch in ((2, 1), (3, 2)):
            b = rand_poly(P, rng, npoly, nq, False, batch=batch)
            # edge values around the centring threshold of the dropped limb
            ql = P.moduli[nq - 1]
            b[0, 0, nq - 1, :4] = [0, ql >> 1, (ql >> 1) + 1, ql - 1]
            assert np.array_equal(gb.to_numpy(gb.rescale(gb.from_numpy(b), nq)), rb.rescale(b, nq)), f"rescale nq={nq}"


def check_keyswitch(P, gb, rb, levels=None, seed=3):
    rng = np.random.default_rng(seed)
    ksk = np.stack([rand_poly(P, rng, 2, P.n_q, True) for _ in range(P.dnum)])
    gk = gb.from_numpy(ksk)
    if levels is None:
        levels = sorted({P.n_q, P.n_q - 1, P.alpha, P.alpha + 1, 1, 2})
    for it, nq in enumerate(levels):
        if nq < 1 or nq > P.n_q:
            continue
        d = rand_poly(P, rng, 1, nq, False, batch=1 + it % 3)
        ge, re_ = gb.modup(gb.from_numpy(d), nq), rb.modup(d, nq)
        assert np.array_equal(gb.to_numpy(ge), re_), f"modup nq={nq}"
        ga, ra = gb.ks_inner(ge, gb.from_numpy(d), gk, nq), rb.ks_inner(re_, d, ksk, nq)
        assert np.array_equal(gb.to_numpy(ga), ra), f"ks_inner nq={nq}"
        gm, rm = gb.moddown(ga, nq), rb.moddown(ra, nq)
        assert np.array_equal(gb.to_numpy(gm), rm), f"moddown nq={nq}"
        assert np.array_equal(gb.to_numpy(gb.keyswitch(gb.from_numpy(d), gk, nq)), rm), f"keyswitch nq={nq}"
        if nq >= 2:
            # ct x ct + relinearise + rescale: the fused call (product formed inside the key switch) against
            # the oracle's tensor product followed by its relinearise + rescale, incl. a batch-broadcast operand
            bt = 1 + it % 3
            a = rand_poly(P, rng, 2, nq, False, batch=bt)
            b = rand_poly(P, rng, 2, nq, False, batch=1 if it % 2 else bt)
            want = rb.relin_rescale(rb.tensor(a, b, nq), ksk, nq)
            got = gb.mul_relin_rescale(gb.from_numpy(a), gb.from_numpy(b), gk, nq)
            assert np.array_equal(gb.to_numpy(got), want), f"mul_relin_rescale nq={nq}"
            assert np.array_equal(gb.to_numpy(gb.relin_rescale(gb.tensor(gb.from_numpy(a), gb.from_numpy(b), nq), gk, nq)),
                                  want), f"relin_rescale nq={nq}"
            if nq < P.n_q:
                # an operand that still carries the limbs of a higher level is used in place
                extra = min(2, P.n_q - nq)
                a_hi = np.concatenate([a, rand_poly(P, rng, 2, nq + extra, False, batch=bt)[:, :, nq:]], axis=2)
                got = gb.mul_relin_rescale(gb.from_numpy(np.ascontiguousarray(a_hi)), gb.from_numpy(b), gk, nq)
                assert np.array_equal(gb.to_numpy(got), want), f"mul_relin_rescale with a higher-level operand nq={nq}"
                got = gb.mul_relin_rescale(gb.from_numpy(b), gb.from_numpy(np.ascontiguousarray(a_hi)), gk, nq)
                assert np.array_equal(gb.to_numpy(got), rb.relin_rescale(rb.tensor(b, a, nq), ksk, nq)), f"... as b, nq={nq}"
            # key switches that share one ModDown: acc = KS_ext(d) + P (l0, 0); acc += KS_ext(d') + P (l0', l1'); acc += P u
            d2 = rand_poly(P, rng, 1, nq, False, batch=bt)
            l1 = rand_poly(P, rng, 1, nq, False, batch=bt)
            l2 = rand_poly(P, rng, 2, nq, False, batch=bt)
            ga = gb.ks_accum(None, gb.from_numpy(d2), gk, gb.from_numpy(l1), nq)
            ra = rb.ks_accum(None, d2, ksk, l1, nq)
            assert np.array_equal(gb.to_numpy(ga), ra), f"ks_accum nq={nq}"
            ga = gb.ks_accum(ga, gb.from_numpy(a[1:2]), gk, gb.from_numpy(l2), nq)
            ra = rb.ks_accum(ra, a[1:2], ksk, l2, nq)
            assert np.array_equal(gb.to_numpy(ga), ra), f"ks_accum accumulate nq={nq}"
            ga = gb.ks_accum(ga, None, None, gb.from_numpy(l2), nq)
            ra = rb.ks_accum(ra, None, None, l2, nq)
            assert np.array_equal(gb.to_numpy(ga), ra), f"ks_accum lift only nq={nq}"
            assert np.array_equal(gb.to_numpy(gb.moddown_rescale(ga, nq)), rb.moddown_rescale(ra, nq)), f"moddown_rescale nq={nq}"
            sq = gb.mul_relin_rescale(gb.from_numpy(a), gb.from_numpy(a), gk, nq)
            assert np.array_equal(gb.to_numpy(sq), rb.relin_rescale(rb.tensor(a, a, nq), ksk, nq)), f"square nq={nq}"


def check_engine_ops(eg, er, slot_tol=1e-5):
    """Same seed => same keys and ciphertexts; every engine op must give identical residues."""
    def same(cg, cr, what):
        assert cg.level == cr.level, what
        assert np.array_equal(eg.backend.to_numpy(cg.polys), er.backend.to_numpy(cr.polys)), what

    keys = []
    for e in (eg, er):
        sk = e.create_secret_key()
        keys.append(dict(sk=sk, pk=e.create_public_key(sk), rlk=e.create_relinearization_key(sk),
                         cj=e.create_conjugation_key(sk), rot=e.create_rotation_key(sk, steps=[1, -1, -2, 4])))
    kg, kr = keys
    assert np.array_equal(eg.backend.to_numpy(kg['rlk'].data), er.backend.to_numpy(kr['rlk'].data)), "rlk"
    sc = eg.slot_count
    rng = np.random.default_rng(11)
    v = np.exp(-2j * np.pi * rng.integers(0, 16, sc) / 16)
    w = rng.random(sc) - 0.5
    cg, cr = eg.encrypt(v, kg['pk']), er.encrypt(v, kr['pk'])
    same(cg, cr, "encrypt")
    dg, dr = eg.encrypt(w, kg['pk']), er.encrypt(w, kr['pk'])
    assert np.allclose(eg.decrypt(cg, kg['sk']), v, atol=1e-6)
    assert np.array_equal(eg.decrypt(cg, kg['sk']), er.decrypt(cr, kr['sk'])), "decrypt"
    m_g, m_r = eg.multiply(cg, dg, kg['rlk']), er.multiply(cr, dr, kr['rlk'])
    same(m_g, m_r, "multiply relin")
    assert np.allclose(eg.decrypt(m_g, kg['sk']), v * w, atol=slot_tol)
    same(eg.multiply(cg, dg), er.multiply(cr, dr), "multiply no relin (3 polys)")
    same(eg.relinearize(eg.multiply(cg, dg), kg['rlk']), er.relinearize(er.multiply(cr, dr), kr['rlk']), "relinearize")
    same(eg.multiply(cg, 0.25 - 0.5j), er.multiply(cr, 0.25 - 0.5j), "multiply const")
    pt_g, pt_r = eg.encode(w), er.encode(w)
    same(eg.multiply(cg, pt_g), er.multiply(cr, pt_r), "multiply plaintext")
    same(eg.add(cg, pt_g), er.add(cr, pt_r), "add plaintext")
    same(eg.add(m_g, cg), er.add(m_r, cr), "add mixed level")
    same(eg.add_plain(cg, 1.0), er.add_plain(cr, 1.0), "add_plain")
    same(eg.conjugate(cg, kg['cj']), er.conjugate(cr, kr['cj']), "conjugate")
    for delta in (1, -2, 3, 5, -1):
        rg, rr = eg.rotate(cg, kg['rot'], delta), er.rotate(cr, kr['rot'], delta)
        same(rg, rr, f"rotate {delta}")
        assert np.allclose(eg.decrypt(rg, kg['sk']), np.roll(v, delta), atol=slot_tol), f"rotate {delta} = np.roll"
    # a batch of 3 ciphertexts goes through the same calls in lockstep
    vb = np.exp(-2j * np.pi * rng.integers(0, 16, (3, sc)) / 16)
    bg, br = eg.encrypt(vb, kg['pk']), er.encrypt(vb, kr['pk'])
    same(bg, br, "encrypt batch")
    assert bg.batch == 3
    sq_g, sq_r = eg.multiply(bg, bg, kg['rlk']), er.multiply(br, br, kr['rlk'])
    same(sq_g, sq_r, "batched multiply relin")
    assert np.allclose(eg.decrypt(sq_g, kg['sk']), vb * vb, atol=slot_tol)
    same(eg.multiply(bg, cg, kg['rlk']), er.multiply(br, cr, kr['rlk']), "batch x single multiply")
    same(eg.multiply(bg, pt_g), er.multiply(br, pt_r), "batch multiply plaintext")
    same(eg.add(bg, pt_g), er.add(br, pt_r), "batch add plaintext")
    same(eg.rotate(bg, kg['rot'], -2), er.rotate(br, kr['rot'], -2), "batch rotate")
    same(eg.conjugate(bg, kg['cj']), er.conjugate(br, kr['cj']), "batch conjugate")
    # hoisted rotations (one ModUp shared) are bit-identical to separate rotations and to the oracle
    fk_g = [eg.create_fixed_rotation_key(kg['sk'], d) for d in (1, -3, 7)]
    fk_r = [er.create_fixed_rotation_key(kr['sk'], d) for d in (1, -3, 7)]
    for hg, k_g, k_r in zip(eg.rotate_hoisted(bg, fk_g), fk_g, fk_r):
        assert np.array_equal(eg.backend.to_numpy(hg.polys), eg.backend.to_numpy(eg.rotate(bg, k_g).polys)), "hoisted == separate rotation"
        same(hg, er.rotate(br, k_r), "hoisted rotation vs oracle")
    # rotate-mask-add in one fused pass (fhe_mul_plain_sum) against the oracle's mul + add sequence
    pts_g = [eg.encode(rng.random(sc) - 0.5) for _ in range(3)]
    cts_g = [bg, eg.rotate(bg, kg['rot'], 1), sq_g]
    cts_r = [br, er.rotate(br, kr['rot'], 1), sq_r]
    pts_r = [er.encode(p.values) for p in pts_g]
    same(eg.multiply_plain_sum(cts_g, pts_g), er.multiply_plain_sum(cts_r, pts_r), "multiply_plain_sum")
    pg, pr = eg.make_power_basis(cg, 5, kg['rlk']), er.make_power_basis(cr, 5, kr['rlk'])
    for k, (x, y) in enumerate(zip(pg, pr), 1):
        same(x, y, f"power {k}")
        assert np.allclose(eg.decrypt(x, kg['sk']), v ** k, atol=slot_tol)


def check_fused_services(P, gpu_backend, ref_backend, batch=2):
    """BSGS S-box and fused 4-bit XOR through fhe_lincomb / fhe_tensor_acc: residues identical
    on both backends, decoded values equal to the AES S-box / a ^ b."""
    from aes_fhe_b200.services.engine_context import EngineContext
    from aes_fhe_b200.services.sbox_service import SBoxService, AES_SBOX
    from aes_fhe_b200.services.xor_service import (XORService, EngineWrapper, XORConfig, CoefficientCache,
                                                   ZetaEncoder)
    outs = []
    for be in (gpu_backend, ref_backend):
        ctx = EngineContext(signature=2, max_level=P.max_level,
                            _engine_kwargs=dict(_params=P, _backend=be, seed=5), rotation_steps=[])
        svc = SBoxService(ctx)
        sc = ctx.engine.slot_count
        rng = np.random.default_rng(3)
        x = rng.integers(0, 256, (batch, sc), dtype=np.uint8)
        x[0, :256] = np.arange(256)
        out = svc.sub_bytes_array_bsgs(ctx.encrypt(ZetaEncoder.to_zeta(x, 256)))
        dec = ZetaEncoder.from_zeta(np.atleast_2d(ctx.decrypt(out)), 256)
        assert np.array_equal(dec, np.array(AES_SBOX, dtype=np.uint8)[x]), "S-box bytes"
        w = EngineWrapper.__new__(EngineWrapper)
        w.ctx, w.engine = ctx, ctx.engine
        w.public_key, w.secret_key, w.relin_key = ctx.public_key, ctx.secret_key, ctx.relinearization_key
        w.conj_key, w.rot_key, w.boot_key = ctx.conjugation_key, ctx.rotation_key, ctx.bootstrap_key
        xs = XORService(w, CoefficientCache(XORConfig().coeffs_path))
        a = rng.integers(0, 16, (batch, sc), dtype=np.uint8)
        b = rng.integers(0, 16, (batch, sc), dtype=np.uint8)
        r = xs.xor_cipher_fused(w.encrypt(ZetaEncoder.to_zeta(a)), w.encrypt(ZetaEncoder.to_zeta(b)))
        assert np.array_equal(ZetaEncoder.from_zeta(np.atleast_2d(w.decrypt(r))), a ^ b), "xor nibbles"
        outs.append((be.to_numpy(out.polys), be.to_numpy(r.polys)))
    assert np.array_equal(outs[0][0], outs[1][0]), "BSGS S-box residues"
    assert np.array_equal(outs[0][1], outs[1][1]), "fused XOR residues"


def check_multiply_gather(P, gpu_backend, ref_backend):
    """Engine.multiply_gather (fhe_mul_relin_rescale_ptrs: the operands of a batched fused multiply gathered through a
    pointer table -- permutations, replications, slices and concatenations of two tensors at different levels) against
    the oracle, which materialises both sides and runs its own tensor / relinearise / rescale: identical residues,
    identical scale bookkeeping, decoded slots equal to the plain products."""
    from aes_fhe_b200.engine import Ciphertext, Engine
    res = []
    for be in (gpu_backend, ref_backend):
        eng = Engine(_params=P, _backend=be, seed=21)
        sk = eng.create_secret_key()
        pk = eng.create_public_key(sk)
        rlk = eng.create_relinearization_key(sk)
        n = eng.slot_count
        rng = np.random.default_rng(6)
        va = rng.uniform(-1, 1, (5, n))
        vb = rng.uniform(-1, 1, (3, n))
        vc = rng.uniform(-1, 1, (2, n))
        L = P.max_level
        a, b, c = eng.encrypt(va, pk, level=L), eng.encrypt(vb, pk, level=L - 1), eng.encrypt(vc, pk, level=L - 1)
        ia, ib, ic = [4, 0, 0, 3, 1, 2, 2], [2, 2, 0, 1, 1], [1, 0]
        out, dev = eng.multiply_gather([(a, ia)], [(b, ib), (c, ic)], rlk)          # a is one level above b and c: used in place
        assert out.level == L - 2 and out.batch == 7
        assert dev == P.delta[L] / P.delta[L - 1]
        want = va[ia] * np.concatenate([vb[ib], vc[ic]])
        got = eng.decrypt(out, sk).real * float(dev)
        assert np.abs(got - want).max() < 5e-5, np.abs(got - want).max()       # CKKS noise of one product at the 40-bit scale
        # the same through the materialised operands and the fused call of the library
        ref = eng._mul_ct_dropped(Ciphertext(eng, be.permute_batch(a.polys, ia), L),
                                  Ciphertext(eng, be.concat_batch([be.permute_batch(b.polys, ib), be.permute_batch(c.polys, ic)]), L - 1), rlk)[0]
        assert np.array_equal(be.to_numpy(ref.polys), be.to_numpy(out.polys))
        res.append(be.to_numpy(out.polys))
    assert np.array_equal(res[0], res[1]), "multiply_gather residues"


def check_double_hoisted_transform(P, gpu_backend, ref_backend, batch=3, half_width=20, stride=2):
    """A BSGS linear transform (the shape of the bootstrap's CoeffToSlot / SlotToCoeff factors: diagonals at the
    rotations stride * k, -half_width <= k < half_width) with double hoisting: the fused kernel fhe_bsgs_inner on
    the GPU side against the same sum assembled from the oracle's primitives (ModUp, key inner product with the
    permuted key, lift, automorphism, extended-basis plaintext products) -- identical residues; slots equal to the
    matrix applied in NumPy; and identical to the single-hoisted evaluation within CKKS noise."""
    from aes_fhe_b200 import bootstrap as B
    from aes_fhe_b200.engine import Engine
    outs, slots = [], []
    for be in (gpu_backend, ref_backend):
        eng = Engine(_params=P, _backend=be, seed=11)
        sk = eng.create_secret_key()
        pk = eng.create_public_key(sk)
        n = eng.slot_count
        rng = np.random.default_rng(4)
        mat = {(stride * k) % n: (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.2
               for k in range(-half_width, half_width)}
        v = rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))
        ct = eng.encrypt(v, pk)
        keys = {}
        entry = B.prepare_matrix(eng, sk, mat, keys, None, double_hoist=True)
        nb = len({b for items in entry["giants"].values() for b, _ in items})
        assert entry["dh"] and nb <= B.BSGS_MAX_BABY and len(entry["giants"]) <= 4
        out = B._linear_transform(eng, ct, entry)
        assert out.level == ct.level - 1
        want = sum(d_vec * np.roll(v, -d, axis=-1) for d, d_vec in mat.items())
        got = eng.decrypt(out, sk)
        assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max()), np.abs(got - want).max()
        single = B._linear_transform(eng, ct, B.prepare_matrix(eng, sk, mat, keys, None, double_hoist=False))
        assert np.abs(eng.decrypt(single, sk) - got).max() < 1e-5 * max(1.0, np.abs(want).max())
        outs.append(be.to_numpy(out.polys))
        slots.append(got)
    assert np.array_equal(outs[0], outs[1]), "double-hoisted transform residues"
