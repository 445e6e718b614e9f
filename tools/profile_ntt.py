"""ncu target: the fused single-launch NTT / iNTT on 32 x 39 limbs of N = 2^16 (numbers printed
under ncu are never reported)."""
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
    x = torch.randint(0, 2 ** 39, (32, tot, n), dtype=torch.int64, device="cuda")
    for _ in range(2):
        gb._call("fhe_ntt_fwd", gb._ptr(x), 32, 31, K)
        gb._call("fhe_ntt_inv", gb._ptr(x), 32, 31, K)
    torch.cuda.synchronize()
    assert gb.lib.fhe_ntt_fused_status(gb.ctx) == 0
    print("profile target done")


if __name__ == "__main__":
    main()
