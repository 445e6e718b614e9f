"""Short, fixed kernel sequence for ncu (one GPU, a few launches of each hot kernel at the
BASELINE size N = 2^16, 31 + 6 limbs).  Not a benchmark: numbers printed under ncu are never
reported."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aes_fhe_b200.backend_cuda import CudaBackend
from aes_fhe_b200.params import make_params


def main():
    P = make_params(16, 30)
    gb = CudaBackend(P)
    n, K = P.n, P.n_p
    tot = P.n_q + K
    x = torch.randint(0, 2 ** 39, (8, tot, n), dtype=torch.int64, device="cuda")
    for _ in range(3):
        gb._call("fhe_ntt_fwd", gb._ptr(x), 8, 31, K)
        gb._call("fhe_ntt_inv", gb._ptr(x), 8, 31, K)
    nq, B = 31, 4
    ksk = torch.randint(0, 2 ** 39, (P.dnum, 2, tot, n), dtype=torch.int64, device="cuda")
    d = torch.randint(0, 2 ** 39, (B, nq, n), dtype=torch.int64, device="cuda")
    out = torch.empty(2, B, nq, n, dtype=torch.int64, device="cuda")
    for _ in range(2):
        gb._call("fhe_keyswitch", gb._ptr(out), gb._ptr(d), gb._ptr(ksk), nq, B)
    c2 = torch.randint(0, 2 ** 39, (2, B, nq, n), dtype=torch.int64, device="cuda")
    r = torch.empty(2, B, nq - 1, n, dtype=torch.int64, device="cuda")
    gb._call("fhe_rescale", gb._ptr(r), gb._ptr(c2), 2 * B, nq)
    torch.cuda.synchronize()
    print("profile target done")


if __name__ == "__main__":
    main()
