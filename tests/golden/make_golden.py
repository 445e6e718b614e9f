"""Generates tests/golden/coeff_digest.json from the reference's committed coefficient JSONs
and S-box table (run in the build container where /root/reference is mounted)."""
import json
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "coeff_digest.json"


def load1d(p):
    d = json.loads(p.read_text())
    out = np.zeros(d["n"], dtype=np.complex128)
    for k, re_, im_ in d["entries"]:
        out[int(k)] = re_ + 1j * im_
    return out


def main():
    xor = json.loads((REF / "xor_mono_coeffs.json").read_text())["entries"]
    hi = load1d(REF / "sbox/coeffs/sbox_hi_coeffs.json")
    lo = load1d(REF / "sbox/coeffs/sbox_lo_coeffs.json")
    idx = [0, 1, 2, 3, 17, 64, 127, 128, 129, 200, 254, 255]
    src = (REF / "sbox/sbox_service.py").read_text()
    tab = [int(x, 16) for x in re.findall(r"0x[0-9a-f]{2}", src[src.index("AES_SBOX = ["):src.index("]", src.index("AES_SBOX = ["))])]
    OUT.write_text(json.dumps({
        "source": "xor_mono_coeffs.json, sbox/coeffs/sbox_{hi,lo}_coeffs.json, sbox/sbox_service.py:31-49",
        "xor_nonzero": len(xor), "xor_entries": xor,
        "sbox_probe_idx": idx,
        "sbox_hi_probe_re": [hi[i].real for i in idx], "sbox_hi_probe_im": [hi[i].imag for i in idx],
        "sbox_lo_probe_re": [lo[i].real for i in idx], "sbox_lo_probe_im": [lo[i].imag for i in idx],
        "sbox_hi_l1": float(np.abs(hi).sum()), "sbox_lo_l1": float(np.abs(lo).sum()),
        "aes_sbox": tab}, indent=0))


if __name__ == "__main__":
    main()
