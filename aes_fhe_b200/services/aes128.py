"""Ten chained homomorphic AES rounds (BASELINE config 5): the FIPS-197-correct round of
``aes_round.AESRoundService`` plus the refresh step that makes chaining possible.

Why a refresh after EVERY LUT layer.  A round is four layers of bivariate 16x16 look-up
polynomials (S-box planes, two XOR layers of MixColumns, AddRoundKey), 5 levels each (+1 for the
rotate-mask-add gather), and every layer amplifies a slot error by up to ~18x (rms 12.5x; measured
on the reference's xor_mono_coeffs.json and on the nibble S-box tables).  Bootstrapping
(``aes_fhe_b200/bootstrap.py``, the reference's ``engine.bootstrap``: xor_service.py:120-129)
restores the levels but adds a slot error of ~3e-3 (max, N = 2^16, 44-bit scale); two chained
layers without correction would leave the decode margin of zeta_16 (0.196).  So each refresh is

    bootstrap  ->  clean(x) = (17 x - x^17) / 16

The clean-up polynomial is the unique degree-17 polynomial with f(z) = z and f'(z) = 0 on the
sixteen 16th roots of unity: an input z (1 + e) comes out as z (1 - 8.5 e^2).  It costs five
levels (x^2, x^4, x^8, x^16, then x^16 * (-x / 16) + 17 x / 16).  A bootstrap (CoeffToSlot and
SlotToCoeff in 3 transforms each) uses 18 of the 30 levels: 12 left, 7 after the clean-up,
enough for one layer (gather + LUT = 6) with one level to spare for the next bootstrap.  A
simulation of the error dynamics (Gaussian refresh noise of 7e-4 rms, 131 072 slots, ten rounds)
stays at 3e-3 before and 6e-5 after each clean-up; without the clean-up the first round fails.

The state stays a nibble pair (ct_hi, ct_lo) of zeta_16-valued ciphertexts, 2048 blocks per
ciphertext, any batch of ciphertexts; all ciphertexts of a layer are stacked on the batch axis so
a layer costs ONE bootstrap call.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from ..engine import Ciphertext
from .aes_round import AESRoundService
from .key_expansion import expand_key


class AES128Service(AESRoundService):
    LUT_LEVELS = 5              # power bases (3) + product + two rescales folded into one more
    CLEAN_LEVELS = 5

    def __init__(self, eng_wrap, xor_svc, boot_groups=(3, 3)):
        super().__init__(eng_wrap, xor_svc)
        self.boot_key = self.engine.create_bootstrap_key(eng_wrap.secret_key, boot_groups[0], boot_groups[1])
        self.refreshes = 0          # ciphertexts refreshed (bootstrap + clean-up)

    # ------------------------------------------------------------------ refresh
    def clean(self, ct: Ciphertext) -> Ciphertext:
        """(17 x - x^17) / 16: pulls every slot back onto the nearest 16th root of unity
        (error e -> 8.5 e^2).  Five levels, five key switches."""
        e, rlk = self.engine, self.eng.relin_key
        x2 = e.multiply(ct, ct, rlk)
        x4 = e.multiply(x2, x2, rlk)
        x8 = e.multiply(x4, x4, rlk)
        x16 = e.multiply(x8, x8, rlk)
        # -x/16 straight onto the level of x^16 and 17x/16 onto the level of the product: one constant multiply +
        # rescale each, on the few limbs those levels have (not a multiply at the top followed by a level adjustment)
        y = e.level_down(ct, x16.level, factor=-1.0 / 16.0)
        z = e.level_down(ct, x16.level - 1, factor=17.0 / 16.0)
        return e.add(e.multiply(x16, y, rlk), z)

    def refresh(self, cts: Sequence[Ciphertext]) -> List[Ciphertext]:
        """bootstrap + clean-up of several ciphertexts in one batched call."""
        e, be = self.engine, self.engine.backend
        lo = min(c.level for c in cts)
        if lo < 1:
            raise RuntimeError("refresh: a ciphertext reached level 0 before it could be bootstrapped")
        cts = [e.level_down(c, lo) for c in cts]
        sizes = [c.batch for c in cts]
        stacked = Ciphertext(e, be.concat_batch([c.polys for c in cts]), lo) if len(cts) > 1 else cts[0]
        fresh = self.clean(e.bootstrap(stacked, self.eng.relin_key, self.eng.conj_key, self.boot_key))
        self.refreshes += sum(sizes)
        if len(cts) == 1:
            return [fresh]
        return [Ciphertext(e, p, fresh.level) for p in be.split_batch(fresh.polys, sizes)]

    def ensure(self, cts: Sequence[Ciphertext], need: int) -> List[Ciphertext]:
        """make sure every ciphertext has `need` levels left (refreshing all of them if not)"""
        if min(c.level for c in cts) >= need:
            return list(cts)
        out = self.refresh(cts)
        if min(c.level for c in out) < need:
            raise RuntimeError(f"refresh leaves {out[0].level} levels, {need} needed: use fewer bootstrap groups")
        return out

    # ------------------------------------------------------------------ rounds with level management
    def round_refreshed(self, state, round_key, last: bool = False, final: bool = False):
        L = self.LUT_LEVELS
        hi, lo = self.ensure(list(state), L + 1)
        if last:
            p = self.sbox_planes((hi, lo), ("S_hi", "S_lo"))
            s_hi, s_lo = self.ensure([p["S_hi"], p["S_lo"]], 1 + L + (0 if final else 1))
            s_hi, s_lo = self.shift_rows((s_hi, s_lo))
        else:
            names = self.PLANES
            planes = self.sbox_planes((hi, lo), names)
            fresh = self.ensure([planes[n] for n in names], 1 + L + 1)
            planes = dict(zip(names, fresh))
            layer1 = []
            for nib in ("hi", "lo"):
                t0 = self._gather(planes["2S_" + nib], 0)
                t1 = self._gather(planes["3S_" + nib], 1)
                t2, t3 = self._gather_many(planes["S_" + nib], [2, 3])
                layer1 += [self._xor(t0, t1), self._xor(t2, t3)]
            a_hi, b_hi, a_lo, b_lo = self.ensure(layer1, L + 1)
            s_hi, s_lo = self.ensure([self._xor(a_hi, b_hi), self._xor(a_lo, b_lo)], L + (0 if final else 1))
        return self.add_round_key((s_hi, s_lo), round_key)

    def encrypt_blocks(self, state, key16, rounds: int = 10):
        """AES-128 (or its first `rounds` rounds followed by nothing) on an encrypted nibble-pair
        state; round keys come from the clear FIPS-197 key schedule (key_expansion.py) and are
        encrypted once each.  With rounds = 10 the last round has no MixColumns."""
        rks = expand_key(bytes(key16))
        st = self.add_round_key(state, self.encrypt_round_key(rks[0]))
        for r in range(1, rounds + 1):
            last = (r == 10)
            st = self.round_refreshed(st, self.encrypt_round_key(rks[r]), last=last, final=(r == rounds))
        return st
