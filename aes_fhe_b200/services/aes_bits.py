"""Bit-sliced homomorphic AES-128 (BASELINE configs 4 and 5): every state bit is a slot value
s = (-1)^bit, so that

  * XOR is a product (MixColumns + AddRoundKey: three levels, 140 ciphertext products per 32 state bits),
  * the S-box is eight multilinear polynomials  out_k = sum_{A,B} w_k[A,B] m_A(hi) m_B(lo)  over the 16 x 16
    monomials of the high / low four input bits (its Walsh-Hadamard spectrum) -- the same bivariate
    baby-step/giant-step shape as the reference's 16 x 16 zeta_16 XOR table
    (/root/reference/xor_service.py:271-286), evaluated by the same fused schedule (aes_fhe_b200/fused.py:
    constant-only inner sums in one pass, lazily relinearised outer sums): 22 + 8 key switches, four levels,
    first-order error gain <= 8 (every partial derivative of a multilinear +-1 function is in {-1, 0, 1}),
  * the refresh -- wherever the levels run out: six times per AES-128 on the default 26-level chain, before rounds
    4..9 (plan_levels) -- is the *bit bootstrap* of aes_fhe_b200/bootstrap.py (``bootstrap_bits``):
    SlotToCoeff at the bottom of the chain, ModRaise with the bits at +- q_0 / 4, CoeffToSlot, and
    EvalMod = sin(2 pi x) whose derivative vanishes exactly there -- the refresh squares the incoming
    error, so no separate clean-up polynomial is needed, and two real ciphertexts share one bootstrap
    (real and imaginary part),
  * every batched product gathers its operands (row rolls, bit-plane slices, replications, concatenations) through
    a pointer table (Engine.multiply_gather) instead of materialising them, the eight S-box sums of a round are
    relinearised in one key switch, and the last round is polished on the way out (final_round_key).

This replaces the zeta_16 nibble-pair pipeline of services/aes128.py (116 refreshed ciphertexts per
2048 blocks, 5-level LUT layers with error gain 18) for the AES-128 throughput path; the reference's
round functions (new.py:186-227, test_all_process.py:21-48) define *what* is computed -- ARK, SubBytes,
ShiftRows, MixColumns on 16-byte states in FIPS order -- and FIPS-197 is the ground truth.

Layout.  One *state* is 32 ciphertexts (bit k = 0..7 of the byte in state row r = 0..3), each packing the four
columns of slot_count / 4 blocks: slot = c * Bs + b for column c of block b (Bs = 8192 at N = 2^16), i.e.
FIPS byte i = 4 c + r of block b sits in ciphertext (k, r), slot c * Bs + b.  ShiftRows then is a pure
rotation of ciphertext (k, r) by r columns -- no masks, no level -- and MixColumns a slot-wise product of
the four row ciphertexts.  G states ride together on the batch axis: batch index = (k * 4 + r) * G + g.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..engine import Ciphertext
from ..fused import _outer_sum
from . import lut
from .key_expansion import expand_key

# xtime on bit planes: (2 t)_k = t_(k-1) ^ (t_7 if k in {0, 1, 3, 4}), FIPS-197 section 4.2.1 (0x1b = bits 0, 1, 3, 4)
_XT_WITH_T7 = (1, 3, 4)
import os as _os
LUT_VIEWS = _os.environ.get("FHE_LUT_VIEWS", "1") != "0"     # A/B switch: LUT inputs as strided views (default) or slice copies


# --------------------------------------------------------------------------- plain bit algebra
def bits_pm(x) -> np.ndarray:
    """integer array [...] -> float64 [8, ...] with (-1)^(bit k of x)"""
    x = np.asarray(x).astype(np.int64)
    return np.stack([1.0 - 2.0 * ((x >> k) & 1) for k in range(8)])


def from_pm(s: np.ndarray) -> np.ndarray:
    """[8, ...] (approximately +-1) -> uint8 [...]"""
    out = np.zeros(s.shape[1:], dtype=np.uint8)
    for k in range(8):
        out |= ((np.real(s[k]) < 0).astype(np.uint8) << k)
    return out


def monomials(s4: np.ndarray) -> np.ndarray:
    """[4, ...] -> [16, ...]: entry A = prod_{j in A} s4[j] (A a 4-bit subset mask; A = 0 is the constant 1)"""
    out = [np.ones_like(s4[0])]
    for a in range(1, 16):
        j = a.bit_length() - 1
        out.append(out[a ^ (1 << j)] * s4[j])
    return np.stack(out)


_WALSH = None


def sbox_walsh() -> np.ndarray:
    """W[k, A, B]: (-1)^(bit k of S(x)) = sum_{A,B} W[k,A,B] m_A(bits 4..7 of x) m_B(bits 0..3 of x).
    The Walsh-Hadamard spectrum of the AES S-box (FIPS-197 5.1.1), derived from the definition."""
    global _WALSH
    if _WALSH is None:
        x = np.arange(256)
        s = bits_pm(x)
        chi = np.einsum("ax,bx->abx", monomials(s[4:]), monomials(s[:4]))          # [16,16,256] characters
        f = bits_pm(lut.AES_SBOX[x])                                                 # [8,256]
        _WALSH = np.einsum("kx,abx->kab", f, chi) / 256.0
    return _WALSH


class PlainBits:
    """The schedule on plain float arrays [8, 4, 4 * nb] (bit, row, column-major slots): the model the
    ciphertext code is tested against (tests/test_aes_bits.py) and the noise study (tools/bits_study.py)."""

    @staticmethod
    def from_blocks(blocks: np.ndarray) -> np.ndarray:
        b = np.asarray(blocks, dtype=np.uint8).reshape(-1, 4, 4)            # [nb, c, r]
        return bits_pm(b.transpose(2, 1, 0).reshape(4, -1))                 # [8, r, c * nb + b]

    @staticmethod
    def from_key(rk16, nb: int) -> np.ndarray:
        return PlainBits.from_blocks(np.tile(np.asarray(rk16, dtype=np.uint8).reshape(1, 16), (nb, 1)))

    @staticmethod
    def to_blocks(st: np.ndarray) -> np.ndarray:
        by = from_pm(st)                                                    # [4(r), 4 nb]
        nb = by.shape[1] // 4
        return by.reshape(4, 4, nb).transpose(2, 1, 0).reshape(nb, 16)

    @staticmethod
    def xor(a, b):
        return a * b

    @staticmethod
    def shift_rows(st):
        nb = st.shape[2] // 4
        return np.stack([np.roll(st[:, r], -r * nb, axis=-1) for r in range(4)], axis=1)

    @staticmethod
    def sub_bytes(st, W=None):
        W = sbox_walsh() if W is None else W
        return np.einsum("kab,a...,b...->k...", W, monomials(st[4:]), monomials(st[:4]))

    @staticmethod
    def mix_ark(a, key, last: bool = False):
        """MixColumns + AddRoundKey (AddRoundKey only when `last`): products in the order the ciphertext code uses"""
        if last:
            return a * key
        rr = lambda v, j: np.roll(v, -j, axis=1)                            # noqa: E731  row r <- row r + j
        t = a * rr(a, 1)
        u = rr(t, 1) * (rr(a, 3) * key)
        xt = np.empty_like(t)
        xt[0] = t[7]
        for k in range(1, 8):
            xt[k] = t[k - 1] * t[7] if k in _XT_WITH_T7 else t[k - 1]
        return xt * u


def len_groups(boot_key) -> int:
    """CoeffToSlot factors of a bootstrap key"""
    return int(getattr(boot_key, "_groups", 3))


# --------------------------------------------------------------------------- ciphertext service
class AESBitService:
    SBOX_LEVELS = 4             # monomials (2) + constant inner sums and outer products (2)
    MIX_LEVELS = 3
    ARK_LEVELS = 1
    FINAL_LEVELS = 2            # last round: AddRoundKey with the half-amplitude key and the sign polish (final_round_key)

    def __init__(self, eng_wrap, boot_groups: Tuple[int, int] = (3, 3), boot_key=None):
        self.eng = eng_wrap
        self.engine = eng_wrap.engine
        self.sc = self.engine.slot_count
        self.Bs = self.sc // 4                      # blocks per state
        self.W = sbox_walsh()
        self._rot_keys: Dict[int, object] = {}
        self.boot_key = boot_key if boot_key is not None else \
            self.engine.create_bootstrap_key(eng_wrap.secret_key, boot_groups[0], boot_groups[1])
        self.boot_in_levels = boot_groups[1] + 1    # SlotToCoeff transforms + the step down to level 0
        self.refreshes = 0                          # complex ciphertexts bootstrapped (two state bits each)
        self.stage_ms: Optional[Dict[str, float]] = None      # filled when `timer` is set
        self.timer = None

    def prepare_keys(self):
        """create (owner) or look up (other ranks) every Galois key the service will use: the three ShiftRows
        rotations and the rotations of the bootstrap's linear transforms.  sharding.distribute_keys ships what the
        owner's engine has issued, so the owner calls this first."""
        from ..bootstrap import _materialise
        for r in (1, 2, 3):
            self._rot_key(r)
        _materialise(self.engine, self.boot_key)

    # ------------------------------------------------------------------ batch-axis plumbing
    def _G(self, ct: Ciphertext) -> int:
        if ct.batch % 32:
            raise ValueError("state ciphertexts carry 32 * G batch elements")
        return ct.batch // 32

    def _take(self, ct: Ciphertext, idx: Sequence[int]) -> Ciphertext:
        return Ciphertext(self.engine, self.engine.backend.permute_batch(ct.polys, list(idx)), ct.level)

    def _slice(self, ct: Ciphertext, lo: int, hi: int) -> Ciphertext:
        return Ciphertext(self.engine, self.engine.backend.slice_batch(ct.polys, lo, hi), ct.level)

    def _view(self, ct: Ciphertext, lo: int, hi: int) -> Ciphertext:
        """batch elements lo..hi-1 for the LUT kernels WITHOUT a copy (the monomials are only ever read by fhe_lincomb
        / fhe_tensor_acc, which take a polynomial stride): 60 GB of slice copies less per AES-128 pass of two states,
        SubBytes 316 -> 304 ms in the steady state of bench.py.  (A single timed pass right after ONE warm-up pass
        -- tools/aes_bits_probe.py -- showed 406 ms instead: the views keep the product tensors alive longer, the
        caching allocator was still growing; tools/lut_layout_bench.py shows the kernels themselves do not care how
        their inputs are laid out.)  LUT_VIEWS = False (FHE_LUT_VIEWS=0) restores the copies."""
        be = self.engine.backend
        if LUT_VIEWS and hasattr(be, "view_batch"):
            return Ciphertext(self.engine, be.view_batch(ct.polys, lo, hi), ct.level)
        return self._slice(ct, lo, hi)

    def _cat(self, cts: Sequence[Ciphertext]) -> Ciphertext:
        e = self.engine
        lvl = min(c.level for c in cts)
        cts = [e.level_down(c, lvl) for c in cts]
        return Ciphertext(e, e.backend.concat_batch([c.polys for c in cts]), lvl)

    @staticmethod
    def _row_roll_index(G: int, j: int, nk: int = 8) -> List[int]:
        """batch permutation: (k, r, g) <- (k, (r + j) % 4, g)"""
        return [(k * 4 + (r + j) % 4) * G + g for k in range(nk) for r in range(4) for g in range(G)]

    # ------------------------------------------------------------------ packing (client side)
    def pack_bits(self, blocks: np.ndarray) -> np.ndarray:
        """blocks [G * Bs (or fewer), 16] bytes -> uint8 bit planes [32 * G, slot_count], row (k * 4 + r) * G + g,
        slot c * Bs + b; unused block slots hold zero bytes"""
        blocks = np.asarray(blocks, dtype=np.uint8).reshape(-1, 16)
        G = max(1, -(-blocks.shape[0] // self.Bs))
        full = np.zeros((G * self.Bs, 16), dtype=np.uint8)
        full[:blocks.shape[0]] = blocks
        by = full.reshape(G, self.Bs, 4, 4).transpose(3, 0, 2, 1).reshape(4, G, self.sc)     # [r, g, c * Bs + b]
        planes = np.stack([(by >> k) & 1 for k in range(8)])                                  # [k, r, g, slot]
        return planes.reshape(32 * G, self.sc)

    def unpack_bits(self, planes: np.ndarray, nb: Optional[int] = None) -> np.ndarray:
        planes = np.asarray(planes, dtype=np.uint8)
        G = planes.shape[0] // 32
        p = planes.reshape(8, 4, G, 4, self.Bs)
        by = np.zeros((4, G, 4, self.Bs), dtype=np.uint8)
        for k in range(8):
            by |= p[k] << k
        out = by.transpose(1, 3, 2, 0).reshape(G * self.Bs, 16)                               # [g, b, c, r]
        return out if nb is None else out[:nb]

    def _encrypt_planes(self, planes: np.ndarray, level: Optional[int], amplitude: float = 1.0) -> Ciphertext:
        """bit planes -> ciphertexts of amplitude * (-1)^bit.  With the engine's device codec the bytes go to the GPU
        and encoding, sampling and encryption happen there (Engine.encrypt_zeta, modulus 2); without it the host
        path keeps the ciphertexts reproducible from the seed (parity tests against the oracle)."""
        if self.engine.device_codec:
            return self.engine.encrypt_zeta(planes, self.eng.public_key, 2, level=level, amplitude=amplitude)
        return self.engine.encrypt(amplitude * (1.0 - 2.0 * planes.astype(np.float64)), self.eng.public_key, level=level)

    def encrypt_state(self, blocks: np.ndarray, level: Optional[int] = None) -> Ciphertext:
        return self._encrypt_planes(self.pack_bits(blocks), level)

    def encrypt_round_key(self, rk16, G: int = 1, level: Optional[int] = None, half: bool = False) -> Ciphertext:
        """one 16-byte round key replicated over every block slot and over the G states of a batch.  half: bit planes
        of amplitude 1/2 -- the form the LAST round key takes so that final_round_key can polish the output."""
        rk = np.tile(np.asarray(rk16, dtype=np.uint8).reshape(1, 16), (self.Bs, 1))
        ct = self._encrypt_planes(self.pack_bits(rk), level, 0.5 if half else 1.0)
        if G > 1:
            ct = self._take(ct, [i for i in range(32) for _ in range(G)])
        ct.amplitude = 0.5 if half else 1.0
        return ct

    def encrypt_round_keys(self, key16, G: int = 1, plan: Optional[Dict[str, object]] = None, rounds: int = 10):
        """the FIPS-197 key schedule of `key16`, every round key encrypted at the level the plan multiplies it in
        (plan_levels) and the key of round 10 at half amplitude"""
        rks = expand_key(bytes(key16))
        lv = (lambda r: None) if plan is None else (lambda r: plan["key_levels"][r])
        return [self.encrypt_round_key(rks[r], G, level=lv(r), half=(r == 10)) for r in range(rounds + 1)]

    # ---- throughput path: raw block bytes cross PCIe (16 B per block each way), bit planes are made on the GPU
    def _plane_index(self, G: int):
        """(byte index into the [G * Bs * 16] block buffer, bit) of every slot of every plane, on the device"""
        import torch
        key = ("plane_index", G)
        cache = self.__dict__.setdefault("_dev_cache", {})
        if key not in cache:
            dev = self.engine.backend.device
            r = torch.arange(4, device=dev).view(1, 4, 1, 1, 1)
            g = torch.arange(G, device=dev).view(1, 1, G, 1, 1)
            c = torch.arange(4, device=dev).view(1, 1, 1, 4, 1)
            b = torch.arange(self.Bs, device=dev).view(1, 1, 1, 1, self.Bs)
            byte = ((g * self.Bs + b) * 16 + 4 * c + r).expand(8, 4, G, 4, self.Bs).reshape(32 * G, self.sc)
            bit = torch.arange(8, device=dev, dtype=torch.uint8).view(8, 1, 1).expand(8, 4 * G, self.sc).reshape(32 * G, self.sc)
            cache[key] = (byte.contiguous(), bit.contiguous())
        return cache[key]

    def encrypt_state_device(self, blocks, level: Optional[int] = None) -> Ciphertext:
        """blocks: uint8 [G * Bs, 16] as a (pinned) torch tensor or array.  One H2D copy of the raw bytes; bit
        extraction, encoding, sampling and encryption on the GPU."""
        import torch
        if not self.engine.device_codec:
            raise RuntimeError("encrypt_state_device needs Engine(device_codec=True)")
        x = blocks if isinstance(blocks, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(blocks, dtype=np.uint8))
        if x.shape[0] % self.Bs:
            raise ValueError(f"a whole number of states ({self.Bs} blocks each) is required; pad with pack_bits")
        G = x.shape[0] // self.Bs
        flat = x.reshape(-1).to(self.engine.backend.device, non_blocking=True)
        byte, bit = self._plane_index(G)
        planes = (flat[byte] >> bit) & 1
        return self.engine.encrypt_zeta(planes, self.eng.public_key, 2, level=level)

    def decrypt_state_device(self, ct: Ciphertext):
        """decrypt + decode + bit packing on the GPU; one D2H copy of the block bytes [G * Bs, 16] (torch uint8, host)"""
        import torch
        G = self._G(ct)
        z = self.engine.decrypt_device(ct, self.eng.secret_key)                  # [32 G, sc] complex
        neg = (z.real < 0).to(torch.uint8).view(8, 4, G, 4, self.Bs)
        by = torch.zeros((4, G, 4, self.Bs), dtype=torch.uint8, device=z.device)
        for k in range(8):
            by |= neg[k] << k
        return by.permute(1, 3, 2, 0).reshape(G * self.Bs, 16).cpu()

    def decrypt_state(self, ct: Ciphertext, nb: Optional[int] = None) -> np.ndarray:
        if self.engine.device_codec:
            planes = np.atleast_2d(self.engine.decrypt_zeta(ct, self.eng.secret_key, 2))
        else:
            planes = (np.atleast_2d(self.eng.decrypt(ct)).real < 0).astype(np.uint8)
        return self.unpack_bits(planes, nb)

    def decrypt_slots(self, ct: Ciphertext) -> np.ndarray:
        return np.atleast_2d(self.eng.decrypt(ct))

    # ------------------------------------------------------------------ stages
    def _tick(self, name: str):
        if self.timer is not None:
            self.timer(name)

    def add_round_key(self, state: Ciphertext, key: Ciphertext) -> Ciphertext:
        out = self.engine.multiply(state, key, self.eng.relin_key)
        self._tick("add_round_key")
        return out

    def final_round_key(self, o: Ciphertext, key: Ciphertext) -> Ciphertext:
        """AddRoundKey of the last round.  With the round key at half amplitude (encrypt_round_key(half=True)) and two
        levels left the output is polished on the way out: for o = s (1 + e), s = +-1 (the SubBytes output) and
        k' = k / 2,   (o k') (3 - o^2) = s k (1 - 1.5 e^2 + ...):  the error the last round has accumulated since the
        previous refresh comes out squared, for two products at the bottom of the chain.  A full-amplitude key (or a
        single level left) gives the plain product."""
        e, rlk = self.engine, self.eng.relin_key
        if getattr(key, "amplitude", 1.0) != 0.5:
            return self.add_round_key(o, key)
        if o.level < 2:
            raise RuntimeError("final_round_key: a half-amplitude key needs two levels (plan_levels provides them)")
        key = e.level_down(key, o.level) if key.level > o.level else key
        n = o.batch
        ident = list(range(n))
        both, _ = e.multiply_gather([(o, ident + ident)], [(key, ident), (o, ident)], rlk)      # o k' | o^2
        z, sq = self._slice(both, 0, n), self._slice(both, n, 2 * n)
        out = e.multiply(z, e.add_plain(e.negate(sq), 3.0, inplace=True), rlk)
        self._tick("add_round_key")
        return out

    def _rot_key(self, r: int):
        key = self._rot_keys.get(r)
        if key is None:
            key = self.engine.create_fixed_rotation_key(self.eng.secret_key, -r * self.Bs)
            self._rot_keys[r] = key
        return key

    def shift_rows(self, state: Ciphertext) -> Ciphertext:
        """row r of every bit plane rotated left by r columns (three batched rotations, no level)"""
        G = self._G(state)
        rows = [[(k * 4 + r) * G + g for k in range(8) for g in range(G)] for r in range(4)]
        parts = [self._take(state, rows[0])]
        for r in (1, 2, 3):
            parts.append(self.engine.rotate(self._take(state, rows[r]), self._rot_key(r)))
        inv = [0] * (32 * G)
        for r in range(4):
            for pos, i in enumerate(rows[r]):
                inv[i] = r * 8 * G + pos
        out = self._take(self._cat(parts), inv)
        self._tick("shift_rows")
        return out

    def _monomial_bases(self, src: Ciphertext, groups: Sequence[Sequence[int]], bt: int):
        """For every group of four +-1 bit planes (bit j of a group = batch elements offs[j] .. offs[j] + bt - 1 of
        `src`): {A: prod_{j in A} bit_j} for the 15 non-empty subsets A.  The 6 pair products of ALL groups in one batched
        multiply, then the 4 triples (pair x single) of all groups in another and the quadruples (pair x pair) in a
        third -- 11 key switches per batch element and group, depth 2.  The products gather their operands from `src`
        and from the pair products (Engine.multiply_gather: no copies of slices or concatenations).  In a triple the
        single is one level above the pair: it is used in place (upper limb ignored, an exact modulus switch) and the
        product's scale leaves the table by delta[l] / delta[l-1]; returns [(monomials, scale factors), ...] for the
        LUT constants to absorb."""
        from fractions import Fraction
        e, rlk = self.engine, self.eng.relin_key
        ng = len(groups)
        R = lambda g, j: list(range(groups[g][j], groups[g][j] + bt))                # noqa: E731
        pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        PR = lambda g, n: list(range((g * 6 + n) * bt, (g * 6 + n + 1) * bt))        # noqa: E731  pair n of group g inside pp
        pp, dp = e.multiply_gather([(src, [i for g in range(ng) for a, _ in pairs for i in R(g, a)])],
                                   [(src, [i for g in range(ng) for _, b in pairs for i in R(g, b)])], rlk)
        assert dp == 1
        p01, p23 = pairs.index((0, 1)), pairs.index((2, 3))
        triples = [(0b0111, p01, 2), (0b1011, p01, 3), (0b1101, p23, 0), (0b1110, p23, 1)]
        tt, d = e.multiply_gather([(pp, [i for g in range(ng) for _, n, _ in triples for i in PR(g, n)])],
                                  [(src, [i for g in range(ng) for _, _, j in triples for i in R(g, j)])], rlk)
        qd, dq = e.multiply_gather([(pp, [i for g in range(ng) for i in PR(g, p01)])],
                                   [(pp, [i for g in range(ng) for i in PR(g, p23)])], rlk)
        assert dq == 1
        out = []
        for g in range(ng):
            mono = {1 << j: self._view(src, groups[g][j], groups[g][j] + bt) for j in range(4)}
            dev = {m: Fraction(1) for m in range(1, 16)}
            for n, (i, j) in enumerate(pairs):
                mono[(1 << i) | (1 << j)] = self._view(pp, (g * 6 + n) * bt, (g * 6 + n + 1) * bt)
            for n, (m, _, _) in enumerate(triples):
                mono[m] = self._view(tt, (g * 4 + n) * bt, (g * 4 + n + 1) * bt)
                dev[m] = d
            mono[0b1111] = self._view(qd, g * bt, (g + 1) * bt)
            out.append((mono, dev))
        return out

    def sub_bytes(self, state: Ciphertext) -> Ciphertext:
        """the S-box on every byte of the state: eight multilinear polynomials over the monomials of the high and
        the low four bits (22 + 8 key switches per (row, state), four levels)"""
        G = self._G(state)
        bt = 4 * G
        (lo, dlo), (hi, dhi) = self._monomial_bases(state, [[k * bt for k in range(4)], [k * bt for k in range(4, 8)]], bt)
        out = _outer_sum(self.engine, self.eng.relin_key, hi, lo, [self.W[k] for k in range(8)], ("sbox-bits",), dhi, dlo,
                         batched=True)                  # bit planes 0..7 one after the other: the state layout
        self._tick("sub_bytes")
        return out

    def mix_columns_ark(self, a: Ciphertext, key: Ciphertext) -> Ciphertext:
        """MixColumns + AddRoundKey on a state that has been through ShiftRows and SubBytes:
            out_r = 2 a_r ^ 3 a_(r+1) ^ a_(r+2) ^ a_(r+3) ^ k_r = xtime(t_r) ^ t_(r+1) ^ a_(r+3) ^ k_r,   t_r = a_r ^ a_(r+1)
        as +-1 products: t (32), a_(r+3) k (32), t_(r+1) (a_(r+3) k) (32), the three xtime bits that take t_7 (12),
        and the final product (32): 140 key switches per state, three levels.  Every product gathers its operands
        (row rolls, bit slices, replications) from the tensors that hold them: no copies."""
        e, rlk = self.engine, self.eng.relin_key
        G = self._G(a)
        S, Q = 32 * G, 4 * G                                                         # a state, one bit plane of it
        key = e.level_down(key, a.level) if key.level > a.level else key
        ident = list(range(S))
        r1, r3 = self._row_roll_index(G, 1), self._row_roll_index(G, 3)
        both, _ = e.multiply_gather([(a, ident + r3)], [(a, r1), (key, ident)], rlk)             # t | a_(r+3) k
        tk = lambda k: list(range(k * Q, (k + 1) * Q))                               # noqa: E731  bit plane k of t inside `both`
        nx = len(_XT_WITH_T7)
        second, _ = e.multiply_gather([(both, r1 + [i for k in _XT_WITH_T7 for i in tk(k - 1)])],
                                      [(both, [S + i for i in ident] + tk(7) * nx)], rlk)  # u | xtime bits 1, 3, 4
        plain_bits = [k for k in range(8) if k not in _XT_WITH_T7]                   # xtime bits that are a plain shift
        src = self._take(both, [i for k in plain_bits for i in (tk(7) if k == 0 else tk(k - 1))])
        src = e.level_down(src, second.level)
        parts = []
        for k in range(8):
            if k in _XT_WITH_T7:
                n = _XT_WITH_T7.index(k)
                parts.append((second, list(range(S + n * Q, S + (n + 1) * Q))))
            else:
                n = plain_bits.index(k)
                parts.append((src, list(range(n * Q, (n + 1) * Q))))
        out, _ = e.multiply_gather(parts, [(second, ident)], rlk)
        self._tick("mix_columns_ark")
        return out

    # ------------------------------------------------------------------ refresh and chained rounds
    def refresh(self, state: Ciphertext, top_level: Optional[int] = None) -> Ciphertext:
        """bit bootstrap of a whole state: bit planes k and k + 4 travel as real and imaginary part.  top_level: the
        level the refresh is raised to (plan_levels: only as high as the rounds up to the next refresh need)"""
        e = self.engine
        G = self._G(state)
        half = 16 * G
        re, im = self._slice(state, 0, half), self._slice(state, half, 2 * half)
        packed = e.add(re, e.multiply_by_i(im, 1))
        out = e.bootstrap_bits(packed, self.eng.relin_key, self.eng.conj_key, self.boot_key, top_level)
        self.refreshes += half
        self._tick("refresh")
        return out

    def round_levels(self, last: bool) -> int:
        return self.SBOX_LEVELS + (self.FINAL_LEVELS if last else self.MIX_LEVELS)

    def plan_levels(self, fresh_level: int, rounds: int = 10) -> Dict[str, object]:
        """Walk the level schedule of encrypt_blocks for an input encrypted at `fresh_level`: which rounds start with
        a refresh and at which level each round key is multiplied in (so the client can encrypt key r at exactly
        that level).  A refresh leaves max_level - depth_bits levels; a round needs 7 (the last: 5) and, unless it
        is the final one, must leave boot_in_levels for the next refresh."""
        from ..bootstrap import DOUBLE_ANGLES_BITS, POLY_DEGREE_BITS, _ps_depth
        depth = len_groups(self.boot_key) + _ps_depth(POLY_DEGREE_BITS) + DOUBLE_ANGLES_BITS
        after_boot = self.engine.max_level - depth
        # pass 1: which rounds start with a refresh (greedy, refreshes at full height)
        lvl = fresh_level - self.ARK_LEVELS
        boots = []
        for r in range(1, rounds + 1):
            last = (r == 10)
            need = self.round_levels(last) + (0 if r == rounds else self.boot_in_levels)
            if lvl < need:
                if lvl < self.boot_in_levels:
                    raise RuntimeError(f"round {r}: {lvl} levels left, the bit bootstrap needs {self.boot_in_levels}")
                boots.append(r)
                lvl = after_boot
                if lvl < need:
                    raise RuntimeError(f"refresh leaves {lvl} levels, a round needs {need}: raise max_level")
            lvl -= self.round_levels(last)
        # pass 2: a refresh is raised only as high as the rounds up to the next refresh (and its entry) need -- the
        # whole bootstrap and those rounds then run on fewer limbs
        tops = {}
        for i, r in enumerate(boots):
            nxt = boots[i + 1] if i + 1 < len(boots) else rounds + 1
            used = sum(self.round_levels(q == 10) for q in range(r, nxt)) + (self.boot_in_levels if nxt <= rounds else 0)
            tops[r] = min(self.engine.max_level, depth + used)
        lvl = fresh_level
        key_levels = [lvl]
        lvl -= self.ARK_LEVELS
        for r in range(1, rounds + 1):
            last = (r == 10)
            if r in tops:
                lvl = tops[r] - depth
            lvl -= self.SBOX_LEVELS
            key_levels.append(lvl)
            lvl -= self.FINAL_LEVELS if last else self.MIX_LEVELS
        return {"key_levels": key_levels, "refresh_before_rounds": boots, "refresh_top_levels": tops, "out_level": lvl}

    def best_fresh_level(self, rounds: int = 10) -> int:
        """the lowest input level with the fewest refreshes (a higher level only makes the first rounds dearer)"""
        best = None
        for f in range(1 + self.boot_in_levels, self.engine.max_level + 1):
            try:
                n = len(self.plan_levels(f, rounds)["refresh_before_rounds"])
            except RuntimeError:
                continue
            if best is None or n < best[0]:
                best = (n, f)
        return best[1]

    def encrypt_blocks(self, state: Ciphertext, key16, rounds: int = 10, round_keys: Optional[Sequence[Ciphertext]] = None):
        """AES-128 (or its first `rounds` rounds) on an encrypted state; the round keys come from the clear
        FIPS-197 key schedule (key_expansion.py) and are encrypted once each unless `round_keys` holds them.
        A round is ShiftRows (moved in front of SubBytes: they commute), [refresh], SubBytes,
        MixColumns + AddRoundKey; the refresh happens only when the levels left would not carry the round and
        the entry of the next bootstrap."""
        G = self._G(state)
        plan = self.plan_levels(state.level, rounds)
        if round_keys is None:
            round_keys = self.encrypt_round_keys(key16, G, plan, rounds=rounds)
        st = self.add_round_key(state, round_keys[0])
        for r in range(1, rounds + 1):
            last = (r == 10)
            need = self.round_levels(last) + (0 if r == rounds else self.boot_in_levels)
            st = self.shift_rows(st)
            if st.level < need:
                if st.level < self.boot_in_levels:
                    raise RuntimeError(f"round {r}: {st.level} levels left, the bit bootstrap needs {self.boot_in_levels}")
                st = self.refresh(st, plan["refresh_top_levels"].get(r))
                if st.level < need:
                    raise RuntimeError(f"refresh leaves {st.level} levels, a round needs {need}: raise max_level")
            st = self.sub_bytes(st)
            st = self.final_round_key(st, round_keys[r]) if last else self.mix_columns_ark(st, round_keys[r])
        return st


# --------------------------------------------------------------------------- bytes in, bytes out (row f-3)
def _run_blocks(svc: AESBitService, blocks: np.ndarray, key16: bytes) -> np.ndarray:
    """AES-128 of [nb, 16] plaintext blocks under `key16`, state by state (Bs blocks per state)"""
    nb = blocks.shape[0]
    st = svc.encrypt_state(blocks, level=1 + svc.boot_in_levels)
    return svc.decrypt_state(svc.encrypt_blocks(st, key16), nb)


def encrypt_ecb(svc: AESBitService, data: bytes, key16: bytes) -> bytes:
    """AES-128-ECB of `data` (PKCS#7-padded, chunked into 16-byte blocks: the reference's utils.pkcs7_pad /
    chunk_bytes, /root/reference/utils.py:62-91) evaluated homomorphically: the blocks are bit-sliced and encrypted
    under the service's public key, run through ten rounds with the encrypted round keys of `key16`, decrypted and
    re-assembled.  What the server sees is ciphertext only; this helper plays client and server."""
    from .utils import chunk_bytes, pkcs7_pad
    chunks = chunk_bytes(pkcs7_pad(bytes(data)))
    blocks = np.frombuffer(b"".join(chunks), dtype=np.uint8).reshape(-1, 16)
    return _run_blocks(svc, blocks, key16).tobytes()


def encrypt_ctr(svc: AESBitService, data: bytes, key16: bytes, nonce12: bytes, counter0: int = 1) -> bytes:
    """AES-128-CTR keystream (nonce || 32-bit big-endian counter, NIST SP 800-38A) evaluated homomorphically and
    XORed onto `data` in the clear -- the transciphering use of this pipeline (the keystream blocks are public
    counters, only the key is secret)."""
    n = -(-len(data) // 16)
    ctr = np.zeros((n, 16), dtype=np.uint8)
    ctr[:, :12] = np.frombuffer(bytes(nonce12), dtype=np.uint8)
    c = (counter0 + np.arange(n, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
    for j in range(4):
        ctr[:, 12 + j] = ((c >> np.uint64(8 * (3 - j))) & np.uint64(0xFF)).astype(np.uint8)
    ks = _run_blocks(svc, ctr, key16).reshape(-1)[:len(data)]
    return (np.frombuffer(bytes(data), dtype=np.uint8) ^ ks).tobytes()
