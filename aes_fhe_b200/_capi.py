"""ctypes binding of libaesfhe_b200.so (the C ABI in include/aesfhe_b200.h).

The product library is built in-tree by ``__graft_entry__.build()`` /
``aes_fhe_b200.build.build_cuda()``; importing this module never falls back to anything
else: a missing library raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libaesfhe_b200.so"

_P = C.c_void_p
_I = C.c_int
_U64 = C.c_uint64
_LL = C.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "fhe_ctx_create": [C.POINTER(_P), _I, _I, _I, _I, _P, _P, _I],
    "fhe_ctx_destroy": [_P],
    "fhe_last_error": [],
    "fhe_launch_count": [],
    "fhe_ntt_row_count": [],
    "fhe_set_ntt_fused": [_P, _I],
    "fhe_ntt_fused_status": [_P],
    "fhe_ntt_fwd": [_P, _P, _P, _I, _I, _I],
    "fhe_ntt_inv": [_P, _P, _P, _I, _I, _I],
    "fhe_add": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I],
    "fhe_sub": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I],
    "fhe_mul": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I],
    "fhe_neg": [_P, _P, _P, _P, _I, _I, _I],
    "fhe_tensor": [_P, _P, _P, _P, _P, _I, _I],
    "fhe_mul_const": [_P, _P, _P, _P, _P, _P, _I, _I, _I],
    "fhe_add_const": [_P, _P, _P, _P, _P, _P, _I, _I, _I],
    "fhe_lincomb": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "fhe_mul_plain_sum": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I],
    "fhe_mul_plain_multi": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I],
    "fhe_mul_relin_rescale_ptrs": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I],
    "fhe_bsgs_inner": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I],
    "fhe_tensor_acc": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _LL, _P],
    "fhe_rescale": [_P, _P, _P, _P, _I, _I],
    "fhe_mod_raise": [_P, _P, _P, _P, _I, _I],
    "fhe_automorphism": [_P, _P, _P, _P, _U64, _I],
    "fhe_keyswitch": [_P, _P, _P, _P, _P, _I, _I],
    "fhe_modup": [_P, _P, _P, _P, _I, _I],
    "fhe_ks_inner": [_P, _P, _P, _P, _P, _P, _I, _I],
    "fhe_moddown": [_P, _P, _P, _P, _I, _I],
    "fhe_relin_rescale": [_P, _P, _P, _P, _P, _I, _I],
    "fhe_mul_relin_rescale": [_P, _P, _P, _P, _I, _P, _I, _P, _I, _I],
    "fhe_ks_accum": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I],
    "fhe_moddown_rescale": [_P, _P, _P, _P, _I, _I],
    "fhe_from_i64": [_P, _P, _P, _P, _I, _I, _I],
    "fhe_crt_centered": [_P, _P, _P, _P, _I, _I],
}
_RESTYPES = {"fhe_ctx_destroy": None, "fhe_last_error": C.c_char_p, "fhe_launch_count": _U64, "fhe_ntt_row_count": _U64}


class FheError(RuntimeError):
    pass


def load(path: Path | str | None = None) -> C.CDLL:
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise FheError(
            f"{p} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, _I)
    return lib


def check(lib: C.CDLL, rc: int, what: str):
    if rc != 0:
        msg = lib.fhe_last_error()
        raise FheError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
