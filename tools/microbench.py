"""Per-kernel timings on the B200 (CUDA events on the launching stream, L2 flushed between
timed iterations).  Prints one JSON line per measurement; `achieved_gbs` uses the
ALGORITHMIC bytes of SURVEY.md section 8(d)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aes_fhe_b200.backend_cuda import CudaBackend
from aes_fhe_b200.params import make_params


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)       # us
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    peak = 6455.9
    try:
        peak = json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    P = make_params(16, 30)
    import os
    if os.environ.get('FHE_LIB'):          # older builds of the library (A/B runs) lack the newest entry points
        import ctypes
        from aes_fhe_b200 import _capi
        probe = ctypes.CDLL(os.environ['FHE_LIB'])
        for name in list(_capi.SIGNATURES):
            if not hasattr(probe, name):
                _capi.SIGNATURES.pop(name)
    gb = CudaBackend(P, _lib_path=os.environ.get('FHE_LIB') or None)
    has_fused = hasattr(gb.lib, "fhe_set_ntt_fused")
    only_ntt = bool(os.environ.get('FHE_ONLY_NTT'))
    n, K = P.n, P.n_p
    limb = n * 8
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    rng = np.random.default_rng(0)

    def rnd(*shape):
        return torch.randint(0, 2 ** 39, shape, dtype=torch.int64, device="cuda")   # valid residues of every modulus

    out = []

    def rec(name, us_med, us_min, alg_bytes, **kw):
        r = dict(kernel=name, us=round(us_med, 2), us_min=round(us_min, 2), algorithmic_MB=round(alg_bytes / 1e6, 2),
                 achieved_gbs=round(alg_bytes / us_med / 1e3, 1), frac_of_measured_peak=round(alg_bytes / us_med / 1e3 / peak, 3), **kw)
        out.append(r)
        print(json.dumps(r), flush=True)

    # copy roofline reference (torch's own copy kernel) on the same buffers
    tot = P.n_q + K
    a = rnd(8, tot, n); b = torch.empty_like(a)
    m, mn = timeit(lambda: b.copy_(a), flush=flush)
    rec("torch_copy_ref", m, mn, 2 * a.numel() * 8)

    modes = [int(m) for m in os.environ.get("FHE_NTT_MODES", "0,2,1").split(",")] if has_fused else [0]
    for fused in modes:
        if has_fused:
            gb.lib.fhe_set_ntt_fused(gb.ctx, fused)
        tag = {0: "two_pass", 1: "fused", 2: "chained"}[fused]
        for rows in (tot, 2 * tot, 8 * tot, 32 * tot):
            x = rnd(rows // tot, tot, n)
            m, mn = timeit(lambda: gb._call("fhe_ntt_fwd", gb._ptr(x), x.shape[0], 31, K), flush=flush)
            rec("ntt_fwd." + tag, m, mn, 2 * rows * limb, rows=rows)
            m, mn = timeit(lambda: gb._call("fhe_ntt_inv", gb._ptr(x), x.shape[0], 31, K), flush=flush)
            rec("ntt_inv." + tag, m, mn, 2 * rows * limb, rows=rows)
    if has_fused:
        assert gb.lib.fhe_ntt_fused_status(gb.ctx) == 0, "fused NTT barrier timed out"
        gb.lib.fhe_set_ntt_fused(gb.ctx, int(os.environ.get("FHE_NTT_FUSED", "0")))

    # streaming kernels at the batch the AES-128 step runs them with (32 ciphertexts of 2 polynomials): the 2-polynomial
    # cases below are launch-latency sized (12-22 us)
    import ctypes as C
    for nq in (() if only_ntt else (27, 14)):
        B = 32
        x = rnd(2, B, nq, n); y = rnd(2, B, nq, n); o = torch.empty_like(x)
        m, mn = timeit(lambda: gb._call("fhe_add", gb._ptr(o), gb._ptr(x), gb._ptr(y), 2, B, 2, B, nq, 0), flush=flush)
        rec("add_ct.batch32", m, mn, 6 * B * nq * limb, nq=nq)
        m, mn = timeit(lambda: gb._call("fhe_mul", gb._ptr(o), gb._ptr(x), gb._ptr(y), 2, B, 1, 1, nq, 0), flush=flush)
        rec("mul_plain_broadcast.batch32", m, mn, (4 * B + 1) * nq * limb, nq=nq)
        m, mn = timeit(lambda: gb._call("fhe_automorphism", gb._ptr(o), gb._ptr(x), C.c_uint64(5), 2 * B * nq), flush=flush)
        rec("automorphism.batch32", m, mn, 4 * B * nq * limb, nq=nq)
    for nq in (() if only_ntt else (31, 21, 11)):
        x = rnd(2, nq, n); y = rnd(2, nq, n); o = torch.empty_like(x)
        m, mn = timeit(lambda: gb._call("fhe_add", gb._ptr(o), gb._ptr(x), gb._ptr(y), 2, 1, 2, 1, nq, 0), flush=flush)
        rec("add_ct", m, mn, 6 * nq * limb, nq=nq)
        m, mn = timeit(lambda: gb._call("fhe_mul", gb._ptr(o), gb._ptr(x), gb._ptr(y), 2, 1, 2, 1, nq, 0), flush=flush)
        rec("mul_pointwise", m, mn, 6 * nq * limb, nq=nq)
        o3 = torch.empty(3, nq, n, dtype=torch.int64, device="cuda")
        m, mn = timeit(lambda: gb._call("fhe_tensor", gb._ptr(o3), gb._ptr(x), gb._ptr(y), nq, 1), flush=flush)
        rec("tensor", m, mn, 7 * nq * limb, nq=nq)
        r = torch.empty(2, nq - 1, n, dtype=torch.int64, device="cuda")
        m, mn = timeit(lambda: gb._call("fhe_rescale", gb._ptr(r), gb._ptr(x), 2, nq), flush=flush)
        rec("rescale", m, mn, (2 * nq + 2 * (nq - 1)) * limb, nq=nq)
        import ctypes as C
        m, mn = timeit(lambda: gb._call("fhe_automorphism", gb._ptr(o), gb._ptr(x), C.c_uint64(5), 2 * nq), flush=flush)
        rec("automorphism", m, mn, 4 * nq * limb, nq=nq)
        # key switch, batch of B ciphertexts sharing one pass over the key
        ksk = rnd(P.dnum, 2, tot, n)
        beta = P.digits_at(nq)
        for B in (1, 4, 8):
            d = rnd(B, nq, n); ko = torch.empty(2, B, nq, n, dtype=torch.int64, device="cuda")
            alg = (2 * beta * (nq + K) + 3 * nq * B) * limb
            m_, mn = timeit(lambda: gb._call("fhe_keyswitch", gb._ptr(ko), gb._ptr(d), gb._ptr(ksk), nq, B), flush=flush)
            rec("keyswitch", m_, mn, alg, nq=nq, beta=beta, batch=B, us_per_ct=round(m_ / B, 1))
            ext = torch.empty(B, beta, nq + K, n, dtype=torch.int64, device="cuda")
            acc = torch.empty(2, B, nq + K, n, dtype=torch.int64, device="cuda")
            m_, mn = timeit(lambda: gb._call("fhe_modup", gb._ptr(ext), gb._ptr(d), nq, B), flush=flush)
            rec("ks.modup", m_, mn, B * (nq + beta * (nq + K)) * limb, nq=nq, batch=B)
            m_, mn = timeit(lambda: gb._call("fhe_ks_inner", gb._ptr(acc), gb._ptr(ext), gb._ptr(d), gb._ptr(ksk), nq, B), flush=flush)
            rec("ks.inner", m_, mn, (2 * beta * (nq + K) + B * (beta * (nq + K) + 2 * (nq + K))) * limb, nq=nq, batch=B)
            m_, mn = timeit(lambda: gb._call("fhe_moddown", gb._ptr(ko), gb._ptr(acc), nq, 2 * B), flush=flush)
            rec("ks.moddown", m_, mn, B * (2 * (nq + K) + 2 * nq) * limb, nq=nq, batch=B)
    Path("gpurun_out").mkdir(exist_ok=True)
    json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
