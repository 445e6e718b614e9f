"""Static SASS instruction histogram of the hot kernels (cuobjdump -sass of the in-tree library): the tracked evidence of
which pipes the kernels use (DFMA / DADD / DMUL / DMMA on the FP64 pipe, LDG / STG / LDS / STS, LDCU = constant-bank loads
into uniform registers, UBLKCP / LDGSTS = bulk / async copies).
    python tools/sass_histogram.py > profiles/r02_sass_histogram.md"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "aes_fhe_b200" / "csrc" / "libaesfhe_b200.so"
KERNELS = OrderedDict([
    ("ntt_fwd_chained<8, LoadPlain, StorePlain>", r"ntt_fwd_chainedILi8E9LoadPlain10StorePlain"),
    ("ntt_fwd_chained<8, LoadPlain, StoreSubMul>", r"ntt_fwd_chainedILi8E9LoadPlain11StoreSubMul"),
    ("ntt_inv_chained<8, LoadPlain, StorePlain>", r"ntt_inv_chainedILi8E9LoadPlain10StorePlain"),
    ("ntt_inv_chained<8, LoadMulPtr, StorePlain>", r"ntt_inv_chainedILi8E10LoadMulPtr10StorePlain"),
    ("k_bconv_param<7, 2>", r"k_bconv_paramILi7ELi2E"),
    ("k_bconv_param<8, 1>", r"k_bconv_paramILi8ELi1E"),
    ("k_bconv<7>", r"_Z7k_bconvILi7E"),
    ("k_ks_inner<4, 2, false>", r"k_ks_innerILi4ELi2ELb0E"),
    ("k_ks_inner<4, 2, AB>", r"k_ks_innerILi4ELi2ELb1E"),
    ("k_ks_inner_ptr<4, 2>", r"k_ks_inner_ptrILi4ELi2E"),
    ("k_bsgs_inner<4, 2, 4, smem>", r"k_bsgs_innerILi4ELi2ELi4ELb1E"),
    ("k_lincomb_mma<4>", r"k_lincomb_mmaILi4E"),
    ("k_tensor_acc", r"k_tensor_acc"),
    ("k_automorphism", r"k_automorphism"),
])
COLS = ["DFMA", "DADD", "DMUL", "DMMA", "IMAD", "LEA", "LDG", "STG", "LDS", "STS", "LDC", "LDCU", "BRA", "BAR", "UBLKCP", "LDGSTS", "SHFL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["total"] += 1
    print("# SASS instruction histograms of the hot kernels (cuobjdump -sass libaesfhe_b200.so, sm_100a)\n")
    print("Static counts per kernel body (`python tools/sass_histogram.py`).  The arithmetic is FP64 (DFMA / DADD / DMUL; DMMA =\n"
          "`mma.sync.m8n8k4.f64` in the LUT inner sums); no tcgen05 / TMA mnemonics appear: the path is FP64-exact integer\n"
          "arithmetic, for which tcgen05 has no operand kind, and the loads are plain coalesced LDG (the NTT keeps its tile in\n"
          "registers and shared memory).  LDCU / LDC: the base-conversion table and the operand pointer table of the gathered\n"
          "multiply are kernel parameters read from the constant bank.\n")
    print("| kernel | " + " | ".join(COLS) + " | total |")
    print("|---|" + "---|" * (len(COLS) + 1))
    for name, pat in KERNELS.items():
        hits = [k for k in per if re.search(pat, k)]
        if not hits:
            print(f"| `{name}` | not in this build |" + " |" * len(COLS))
            continue
        c = per[hits[0]]
        print(f"| `{name}` | " + " | ".join(str(c.get(x, 0)) for x in COLS) + f" | {c['total']} |")


if __name__ == "__main__":
    main()
