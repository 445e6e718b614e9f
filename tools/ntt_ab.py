"""A/B timing of the NTT on the B200: every (library, environment) variant runs in its own process (the switches are read when
the context is created), CUDA events on the launching stream, L2 flushed between timed iterations, outputs hashed so that the
variants can be checked against each other.
    python tools/ntt_ab.py                      # driver: runs all variants listed in VARIANTS (edit per experiment)
The measurement-only libraries of the last experiment (profiles/r02_ntt_experiments.md) are built by hand into
aes_fhe_b200/csrc/variants/ (git-ignored) with the flags of aes_fhe_b200/build.py plus
    -DFHE_NTT_DIAG_NOMEM                          -> libnomem.so         (no global data traffic)
    -DFHE_NTT_DIAG_NOMEM -DFHE_NTT_DIAG_NOSMEM    -> libnomem_nosmem.so  (and no shared-memory exchange in the forward transform)
A variant whose library is missing is reported as FAILED and skipped.
    python tools/ntt_ab.py --one                # worker (FHE_LIB / FHE_CHAIN_PERSIST / ... from the environment)"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
V = ROOT / "aes_fhe_b200" / "csrc" / "variants"
VARIANTS = [
    ("shipped", {}),
    ("diagnostic: no global data traffic (arithmetic, twiddle loads and shared-memory exchanges only)", {"FHE_LIB": str(V / "libnomem.so")}),
    ("diagnostic: no global data traffic, no shared-memory exchange", {"FHE_LIB": str(V / "libnomem_nosmem.so")}),
]


def worker():
    import torch
    from aes_fhe_b200 import _capi
    from aes_fhe_b200.backend_cuda import CudaBackend
    from aes_fhe_b200.params import LOG_PQ_BUDGET_SPARSE, make_params
    if os.environ.get("FHE_LIB"):
        import ctypes
        probe = ctypes.CDLL(os.environ["FHE_LIB"])
        for name in list(_capi.SIGNATURES):
            if not hasattr(probe, name):
                _capi.SIGNATURES.pop(name)
    P = make_params(16, 26, scale_bits=44, log_pq_budget=LOG_PQ_BUDGET_SPARSE)      # the AES-128 set: 27 + 7 limbs
    gb = CudaBackend(P, _lib_path=os.environ.get("FHE_LIB") or None)
    n, K, nq = P.n, P.n_p, P.n_q
    tot = nq + K
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = {"limbs": [nq, K]}
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for polys in (128, 16, 2):
        x0 = torch.randint(0, 2 ** 39, (polys, tot, n), dtype=torch.int64, device="cuda", generator=g)
        for name in ("fhe_ntt_fwd", "fhe_ntt_inv"):
            x = x0.clone()
            gb._call(name, gb._ptr(x), polys, nq, K)
            torch.cuda.synchronize()
            w = torch.arange(1, 1 + x.numel(), dtype=torch.int64, device="cuda").view_as(x)
            h = int(((x * w).sum()).item())
            ts = []
            for it in range(13):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); gb._call(name, gb._ptr(x), polys, nq, K); b.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            rows = polys * tot
            res[f"{name}.{rows}"] = {"us": round(ts[len(ts) // 2], 1), "us_min": round(ts[0], 1),
                                     "us_per_row": round(ts[len(ts) // 2] / rows, 4), "hash": h}
    assert gb.lib.fhe_ntt_fused_status(gb.ctx) == 0 or "DIAG" in os.environ.get("FHE_LIB", "") or "nomem" in os.environ.get("FHE_LIB", "")
    print("RESULT " + json.dumps(res))


def main():
    if "--one" in sys.argv:
        return worker()
    out = []
    for tag, env in VARIANTS:
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "--one"], env=e, capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(tag, "FAILED", r.stderr[-800:]); continue
        d = json.loads(line[0][7:]); d["variant"] = tag
        out.append(d)
        print(json.dumps(d), flush=True)
    ref = out[0] if out else None
    for d in out[1:]:
        for k, v in d.items():
            if isinstance(v, dict) and v["hash"] != ref[k]["hash"]:
                print("HASH MISMATCH", d["variant"], k)
    Path("gpurun_out").mkdir(exist_ok=True)
    json.dump(out, open("gpurun_out/ntt_ab.json", "w"), indent=1)


if __name__ == "__main__":
    main()
