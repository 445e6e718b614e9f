"""Where the time of one refresh (bootstrap + clean-up) goes on the B200, N = 2^16, L = 30."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import aes_fhe_b200.bootstrap as B
from aes_fhe_b200.services.aes128 import AES128Service
from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache

cfg = XORConfig()
w = EngineWrapper(cfg, _engine_kwargs=dict(seed=3), rotation_steps=[])
svc = AES128Service(w, XORService(w, CoefficientCache(cfg.coeffs_path)))
e = w.engine
bt = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(0)
v = np.exp(-2j * np.pi * rng.integers(0, 16, (bt, e.slot_count)) / 16)
ct = e.encrypt(v, w.public_key, level=2)
svc.refresh([ct]); torch.cuda.synchronize()          # warm: keys, matrices
marks = []
orig_lt = B._linear_transform
def timed(name, fn):
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); c0 = dict(e.op_counts)
        r = fn(*a, **k)
        torch.cuda.synchronize(); marks.append((name, time.perf_counter() - t0, sum(v - c0.get(k, 0) for k, v in e.op_counts.items() if k.startswith("keyswitch"))))
        return r
    return wrap
B._linear_transform = timed("linear_transform", orig_lt)
B.chebyshev_eval_ps = timed("chebyshev_eval_ps", B.chebyshev_eval_ps)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = e.bootstrap(ct, w.relin_key, w.conj_key, svc.boot_key)
torch.cuda.synchronize(); t_boot = time.perf_counter() - t0
t0 = time.perf_counter(); cl = svc.clean(out); torch.cuda.synchronize(); t_clean = time.perf_counter() - t0
print(f"batch {bt}: bootstrap {t_boot*1e3:.0f} ms ({t_boot/bt*1e3:.1f} ms/ct), clean {t_clean*1e3:.0f} ms")
for name, t, ks in marks:
    print(f"  {name:18s} {t*1e3:7.1f} ms  key-switch ops {ks}")
print("  other (mod raise, conj, double angles, ...)", f"{(t_boot - sum(t for _, t, _ in marks))*1e3:.1f} ms")
print("err", np.abs(e.decrypt(cl, w.secret_key) - v).max())
