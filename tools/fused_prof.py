"""Where a CTA of the fused NTT spends its cycles (FHE_FUSED_PROFILE build, csrc/variants/libprof.so:
`python -c "from aes_fhe_b200.build import build_profile_variant as b; b()"` before `gpurun`)."""
import ctypes as C
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aes_fhe_b200.backend_cuda import CudaBackend
from aes_fhe_b200.params import make_params

P = make_params(16, 30)
gb = CudaBackend(P, _lib_path=str(Path(__file__).resolve().parent.parent / "aes_fhe_b200/csrc/variants/libprof.so"))
n, K = P.n, P.n_p
tot = P.n_q + K
buf = (C.c_ulonglong * (8 + 4 * 2048))()
for rows_ct in (8, 32):
    x = torch.randint(0, 2 ** 39, (rows_ct, tot, n), dtype=torch.int64, device="cuda")
    for _ in range(2):
        gb._call("fhe_ntt_fwd", gb._ptr(x), rows_ct, 31, K)
    gb.lib.fhe_fused_profile(gb.ctx, buf)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gb._call("fhe_ntt_fwd", gb._ptr(x), rows_ct, 31, K); e.record(); torch.cuda.synchronize()
    gb.lib.fhe_fused_profile(gb.ctx, buf)
    tot_c, wa, wb, pa, pb, nrows = [buf[i] for i in range(6)]
    print(f"rows={rows_ct*tot} us={s.elapsed_time(e)*1e3:.1f} per-CTA-row cycles: total={tot_c/nrows:.0f} phaseA={pa/nrows:.0f} (waitB {wb/nrows:.0f}) "
          f"phaseB={pb/nrows:.0f} (waitA {wa/nrows:.0f})  cta-rows={nrows}")

import numpy as np
a = np.array(buf[8:8 + 4 * 864], dtype=np.uint64).reshape(-1, 4).astype(np.int64)
t0 = a[:, 1].min()
dur = (a[:, 2] - a[:, 1]) / 1e3
print("CTA duration us: min %.0f mean %.0f max %.0f; start spread %.1f us; end spread %.1f us" % (
    dur.min(), dur.mean(), dur.max(), (a[:, 1].max() - t0) / 1e3, (a[:, 2].max() - a[:, 2].min()) / 1e3))
import os
packed = not (int(os.environ.get("FHE_FUSED_FLAGS", "0")) & 1)
gi = (np.arange(864) // 32) if packed else (np.arange(864) % 27)
g = np.stack([dur[gi == i] for i in range(27)])
print("per-group duration us (max over CTAs):", np.round(g.max(axis=1)).astype(int).tolist())
print("per-group effective MHz:", [int(a[gi == i, 3].mean() / g[i].mean()) for i in range(27)])
print("per-group wait fraction:", np.round(np.array([a[gi == i, 3].mean() for i in range(27)]) / (g.mean(axis=1) * 1.9e3), 2).tolist())
cnt = np.bincount(a[:, 0], minlength=148)
print("CTAs per SM histogram:", np.bincount(cnt).tolist())
sm_of_group = [sorted(set(a[gi == i, 0].tolist())) for i in range(27)]
print("group 0 SMs:", sm_of_group[0])
slow = int(np.argmax(g.max(axis=1))); print("slowest group", slow, "SM load:", [int(cnt[s]) for s in sm_of_group[slow]])
fast = int(np.argmin(g.max(axis=1))); print("fastest group", fast, "SM load:", [int(cnt[s]) for s in sm_of_group[fast]])

import json
json.dump(a.tolist(), open("gpurun_out/fused_prof_records.json", "w"))
