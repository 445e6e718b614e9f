#!/usr/bin/env python
"""Benchmark of the homomorphic-AES hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

Workload (BASELINE.json configs[1]): SubBytes on 2K-block-packed ciphertexts -- the function of
the reference's ``SBoxService.sub_bytes_array`` (zeta_256^x -> zeta_256^S(x) through the
sbox_hi / sbox_lo degree-255 LUT polynomials and their product;
/root/reference/sbox/sbox_service.py:116-138) at N = 2^16, max_level = 22
(test/test_sbox_service.py:19), on a batch of B ciphertexts per GPU; one ciphertext packs
slot_count/16 = 2048 AES blocks.  Both arms evaluate it with the baby-step/giant-step
Paterson-Stockmeyer schedule of aes_fhe_b200/fused.py on the product polynomial hi * lo folded by
conjugate symmetry (SBoxService.sub_bytes_array_bsgs: 23 key switches, 10 levels); the reference's
own operation order (255 key switches) is timed with --reference-order.

A "step" is one pass of that path over one batch.  `value` = AES blocks per second with the
input ciphertexts resident in HBM; `e2e` = the same through the public API from host buffers
(encode + encrypt + H2D, SubBytes, decrypt + D2H + decode inside the timed region).
Ciphertext batches shard across ranks with no data-path collective (weak scaling); NCCL is
used once to broadcast the evaluation key and once to gather result checksums.

`--impl reference` times the same operation sequence on the CPU oracle (oracle/refmod.cpp,
OpenMP over all host cores): the reference's own arithmetic lives in the closed `desilofhe`
wheel, so the oracle port is the CPU arm (cpu_baseline.kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "homomorphic AES blocks/sec (SubBytes stage, 2048 blocks per ciphertext)"
UNIT = "blocks/s"
MAX_LEVEL = 22
LOG_N = 16
DNUM = None          # key-switch digit count; None = the parameter set's default (params.make_params)


def _params(level=MAX_LEVEL):
    from aes_fhe_b200.params import make_params
    return make_params(LOG_N, level) if DNUM is None else make_params(LOG_N, level, dnum=DNUM)


def _peak_gbs():
    try:
        return json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def _inputs(slot_count: int, batch: int, seed: int) -> np.ndarray:
    """test_sbox_array_simd's input (tile(arange(256))) for ciphertext 0, seeded uniform bytes
    for the others (test/test_sbox_service.py:55-66)."""
    rng = np.random.default_rng(seed)
    rows = [np.tile(np.arange(256, dtype=np.uint8), slot_count // 256 + 1)[:slot_count]]
    for _ in range(batch - 1):
        rows.append(rng.integers(0, 256, slot_count, dtype=np.uint8))
    return np.stack(rows)


def _make_service(backend=None, seed=1, device_id=0):
    from aes_fhe_b200.params import make_params
    from aes_fhe_b200.services.engine_context import EngineContext
    from aes_fhe_b200.services.sbox_service import SBoxService
    kw = dict(_params=_params(), seed=seed)
    if backend is not None:
        kw["_backend"] = backend
    ctx = EngineContext(signature=2, max_level=MAX_LEVEL, mode="parallel", device_id=device_id, _engine_kwargs=kw,
                        rotation_steps=[])
    return ctx, SBoxService(ctx)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (nvidia-smi's numbers via NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------- CPU arm
def cpu_arm(threads: int = 0):
    """One full pass of the same schedule (sub_bytes_array_bsgs, 23 key switches) on ONE
    ciphertext (2048 blocks) on the CPU oracle, N = 2^16, max_level 22, all host cores."""
    from oracle.refmod import RefBackend, build
    build()
    from aes_fhe_b200.params import make_params
    from aes_fhe_b200.services.xor_service import ZetaEncoder
    from aes_fhe_b200.services.sbox_service import AES_SBOX
    P = _params()
    be = RefBackend(P, threads=threads)
    ctx, svc = _make_service(backend=be)
    eng = ctx.engine
    x = _inputs(eng.slot_count, 1, 0)[0]
    ct = ctx.encrypt(ZetaEncoder.to_zeta(x, 256))
    t0 = time.perf_counter()
    out = svc.sub_bytes_array_bsgs(ct)
    full = time.perf_counter() - t0
    got = ZetaEncoder.from_zeta(ctx.decrypt(out), 256)
    assert np.array_equal(got, np.array(AES_SBOX, dtype=np.uint8)[x])
    blocks = eng.slot_count // 16
    return {"value": blocks / full, "unit": UNIT, "cores": be.threads, "kind": "port",
            "sample": f"one ciphertext ({blocks} blocks), full sub_bytes_array_bsgs at N=2^16, L=22: {full:.2f}s "
                      f"on {be.threads} threads (oracle/refmod.cpp)"}, full


def full_round_probe(batch: int):
    """configs[3]: AddRoundKey_0 + one full AES round (SubBytes, ShiftRows, MixColumns,
    AddRoundKey) on `batch` ciphertexts of 2048 blocks at N = 2^16, L = 30; FIPS-197 App. B is
    block 0 and the decoded bytes of every block are checked against plain AES."""
    import torch
    from aes_fhe_b200.params import make_params
    from aes_fhe_b200.services.aes_round import AESRoundService
    from aes_fhe_b200.services.key_expansion import expand_key
    from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache
    from oracle import aes_plain as A
    cfg = XORConfig()
    w = EngineWrapper(cfg, _engine_kwargs=dict(_params=make_params(LOG_N, 30), seed=2), rotation_steps=[])
    svc = AESRoundService(w, XORService(w, CoefficientCache(cfg.coeffs_path)))
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    rks = expand_key(key)
    rng = np.random.default_rng(7)
    blocks = [rng.integers(0, 256, (svc.B, 16), dtype=np.uint8) for _ in range(batch)]
    blocks[0][0] = np.frombuffer(bytes.fromhex("3243f6a8885a308d313198a2e0370734"), np.uint8)
    st = svc.encrypt_state(blocks)
    k0, k1 = svc.encrypt_round_key(rks[0]), svc.encrypt_round_key(rks[1])

    def one():
        return svc.round(svc.add_round_key(st, k0), k1)

    r1 = one()                                     # warm-up: builds rotation keys, LUT tables
    torch.cuda.synchronize()
    c0 = dict(w.engine.op_counts)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r1 = one(); b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    ks = sum(v - c0.get(k, 0) for k, v in w.engine.op_counts.items() if k.startswith("keyswitch"))
    got = svc.decrypt_state(r1)
    ok = np.array_equal(got, A.round_fn(np.stack(blocks) ^ rks[0], rks[1]))
    fips = got.reshape(-1, 16)[0].tobytes().hex() == "a49c7ff2689f352b6b5bea43026a5049"
    nblk = batch * svc.B
    return {"workload": "configs[3]: AddRoundKey_0 + full round 1, N=2^16, L=30", "batch": batch, "ms": ms,
            "blocks_per_s": nblk / (ms * 1e-3), "keyswitches_per_ciphertext_pair": int(ks),
            "levels_used": 30 - r1[0].level, "bytes_equal_plain_aes": bool(ok), "fips197_appendix_b_round2_input": bool(fips)}


def aes128_probe(batch: int):
    """configs[4]: full AES-128 (ten rounds, clear key schedule, refresh = bootstrap + clean-up after
    every LUT layer; aes_fhe_b200/services/aes128.py) on `batch` ciphertexts of 2048 blocks, N = 2^16,
    L = 30.  Timed on the second run (keys, matrices and LUT tables exist); every decoded block is
    checked against plain AES, block 0 is FIPS-197 Appendix B."""
    import torch
    from aes_fhe_b200.services.aes128 import AES128Service
    from aes_fhe_b200.services.xor_service import XORService, EngineWrapper, XORConfig, CoefficientCache
    from oracle import aes_plain as A
    cfg = XORConfig()
    w = EngineWrapper(cfg, _engine_kwargs=dict(seed=3), rotation_steps=[])
    svc = AES128Service(w, XORService(w, CoefficientCache(cfg.coeffs_path)))
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    rng = np.random.default_rng(9)
    blocks = [rng.integers(0, 256, (svc.B, 16), dtype=np.uint8) for _ in range(batch)]
    blocks[0][0] = np.frombuffer(bytes.fromhex("3243f6a8885a308d313198a2e0370734"), np.uint8)
    st = svc.encrypt_state(blocks)
    svc.encrypt_blocks(st, key)                               # warm-up: builds every key / matrix / table
    torch.cuda.synchronize()
    c0, r0 = dict(w.engine.op_counts), svc.refreshes
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = svc.encrypt_blocks(st, key); b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    got = svc.decrypt_state(out).reshape(batch, svc.B, 16)
    want = np.stack([A.encrypt_blocks(bl, key) for bl in blocks])
    ks = sum(v - c0.get(k, 0) for k, v in w.engine.op_counts.items() if k.startswith("keyswitch"))
    nblk = batch * svc.B
    return {"workload": "configs[4]: AES-128, 10 rounds, N=2^16, L=30, refresh after every LUT layer", "batch": batch,
            "ms": ms, "ms_per_round": ms / 10, "blocks_per_s": nblk / (ms * 1e-3),
            "bootstrap_calls": int(w.engine.op_counts["bootstrap"] - c0.get("bootstrap", 0)),
            "ciphertexts_refreshed_per_batch_element": int((svc.refreshes - r0) // batch),
            "keyswitches": int(ks), "bytes_equal_plain_aes": bool(np.array_equal(got, want)),
            "fips197_appendix_b_ciphertext": bool(got[0, 0].tobytes().hex() == "3925841d02dc09fbdc118597196a0b32")}


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="ciphertexts per GPU per step (16: 727k blocks/s, 32: 736k)")
    ap.add_argument("--dnum", type=int, default=None, help="key-switch digit count of the SubBytes parameter set")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-round", action="store_true",
                    help="skip the extra measurement of one full AES round (configs[3])")
    ap.add_argument("--no-aes128", action="store_true",
                    help="skip the extra measurement of full AES-128 (configs[4], ten rounds with bootstrapping)")
    ap.add_argument("--reference-order", action="store_true",
                    help="also time the reference's own 255-key-switch operation order")
    args = ap.parse_args()
    global DNUM
    DNUM = args.dnum

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "configs[1]: SubBytes (sbox_hi x sbox_lo degree-255 zeta_256 LUT polynomials, BSGS schedule) on "
                          "2048-block ciphertexts, N=2^16, max_level=22",
              "batch_ciphertexts_per_gpu": args.batch, "blocks_per_ciphertext": (1 << (LOG_N - 1)) // 16,
              "keyswitch_digits": _params().dnum, "special_primes": _params().n_p,
              "l2": "working set per step (GBs of power-basis ciphertexts) exceeds the 126 MB L2",
              "sharding": "independent ciphertext batches per rank, no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        times = []
        for _ in range(max(1, min(args.steps, 3))):
            cb, full = cpu_arm()
            times.append(full)
        full = float(np.median(times))
        blocks = (1 << (LOG_N - 1)) // 16
        v = blocks / full
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64-exact-int",
                          "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from aes_fhe_b200.services.xor_service import ZetaEncoder
    from aes_fhe_b200.services.sbox_service import AES_SBOX
    import aes_fhe_b200.backend_cuda as bc

    ctx, svc = _make_service(seed=1, device_id=local)
    eng = ctx.engine
    be = eng.backend
    if world > 1:
        # evaluation key comes from rank 0 over NCCL/NVLink (every rank derived the same key from
        # the shared seed; the broadcast is the deployment path and must be a no-op on the bits)
        from aes_fhe_b200.sharding import broadcast_handle
        before = svc.rlk.data.clone() if rank else None
        svc.rlk.data = broadcast_handle(be, svc.rlk.data, src=0)
        if rank:
            assert torch.equal(before, svc.rlk.data)
    sc = eng.slot_count
    sbox = np.array(AES_SBOX, dtype=np.uint8)
    data = _inputs(sc, args.batch, seed=rank)
    zeta = ZetaEncoder.to_zeta(data, 256)
    ct_in = eng.encrypt(zeta, ctx.public_key)

    def step():
        return svc.sub_bytes_array_bsgs(ct_in)

    pinned = torch.from_numpy(data).pin_memory()             # host input buffer of the user: the bytes

    def e2e_step():
        # bytes in, bytes out (the reference's XORService.xor / SBoxService tests do zeta-encode + encrypt
        # ... decrypt + decode around the service call): zeta codec, sampling and decode on the GPU
        out = svc.sub_bytes_array_bsgs(eng.encrypt_zeta(pinned, ctx.public_key, 256))     # H2D inside
        return eng.decrypt_zeta(out, ctx.secret_key, 256)                                 # D2H inside

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # correctness of what is being timed (also the first warm-up)
    out = step()
    got = ZetaEncoder.from_zeta(np.atleast_2d(eng.decrypt(out, ctx.secret_key)), 256)
    assert np.array_equal(got, sbox[data]), "SubBytes output differs from the AES S-box"
    for _ in range(max(0, args.warmup - 1)):
        step()

    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    l0 = be.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.nvtx.range_push("timed_region")
    ev[0].record()
    for _ in range(args.steps):
        step()
    ev[1].record()
    torch.cuda.nvtx.range_pop()
    sync_all()
    launches = be.launch_count() - l0
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    clocks = sampler.stop()

    # end-to-end through the public API from host buffers
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / max(1, args.steps // 2) * 1e3
    assert np.array_equal(np.atleast_2d(res), sbox[data])

    # dominant kernel: the forward NTT (both passes, one chained launch), timed alone on this stream at
    # the row count one key-switch ModUp of this batch launches
    nq = MAX_LEVEL + 1
    P = eng.params
    rows = args.batch * P.digits_at(nq) * (nq + P.n_p)
    x = torch.randint(0, 2 ** 39, (rows // (nq + P.n_p), nq + P.n_p, P.n), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        be._call("fhe_ntt_fwd", be._ptr(x), x.shape[0], nq, P.n_p)
    kt = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); be._call("fhe_ntt_fwd", be._ptr(x), x.shape[0], nq, P.n_p); b.record()
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b))
    k_ms = float(np.median(kt))
    alg_bytes = 2 * rows * P.n * 8                      # read once + write once per limb
    peak, peak_kind = _peak_gbs()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(ROOT / "profiles" / "ntt_traffic.json"))["dram_bytes_per_row"] * rows
    except Exception:
        pass

    # second kernel of the key switch, timed the same way: the key inner product (HBM-bound).  Algorithmic bytes per
    # launch: the key 2 beta (n+K) limbs once per batch + per ciphertext beta (n+K) extended rows in
    # (own-digit rows come from the n input limbs) and 2 (n+K) accumulator rows out
    beta, ne = P.digits_at(nq), nq + P.n_p
    ext = torch.randint(0, 2 ** 39, (args.batch, beta, ne, P.n), dtype=torch.int64, device="cuda")
    dd = torch.randint(0, 2 ** 39, (1, args.batch, nq, P.n), dtype=torch.int64, device="cuda")
    acc = torch.empty(2, args.batch, ne, P.n, dtype=torch.int64, device="cuda")
    ksk = svc.rlk.data
    kt2 = []
    for it in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); be._call("fhe_ks_inner", be._ptr(acc), be._ptr(ext), be._ptr(dd), be._ptr(ksk), nq, args.batch); b.record()
        torch.cuda.synchronize()
        if it >= 3:
            kt2.append(a.elapsed_time(b))
    ks_ms = float(np.median(kt2))
    ks_bytes = (2 * beta * ne + args.batch * (beta * ne + 2 * ne)) * P.n * 8
    ks_inner = {"kernel": "k_ks_inner", "bound": "hbm", "achieved": ks_bytes / (ks_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": ks_bytes / (ks_ms * 1e-3) / 1e9 / peak, "us": ks_ms * 1e3,
                "algorithmic_bytes": int(ks_bytes)}
    del ext, dd, acc

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        chk = torch.tensor([int(got.astype(np.uint64).sum())], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
    ms, e2e_ms = float(t[0]), float(t[1])
    blocks = world * args.batch * (sc // 16)
    value = blocks / (ms * 1e-3)
    e2e_value = blocks / (e2e_ms * 1e-3)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64-exact-int", "data": "synthetic", "config": config,
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(data.nbytes),
                        "d2h_bytes_per_step": int(data.nbytes)},
                "roofline": {"bound": "hbm", "kernel": "ntt_fwd_chained (both radix passes in one launch)", "rows_per_launch": rows,
                             "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic,
                             "note": "FP64-pipe bound: 8 FP64 ops per butterfly, pipe at 70 % (profiles/r01_ncu_chained.md); "
                                     "traffic = ncu dram bytes of the launch; the NTT launches are 47 % of the step "
                                     "(profiles/r01_kernel_breakdown_final.md)"},
                "roofline_keyswitch_inner": ks_inner,
                "ms_per_ciphertext": ms / args.batch,
                "keyswitches_per_ciphertext": 23}
        if not args.no_full_round and world == 1:
            try:
                line["full_round"] = full_round_probe(min(args.batch, 4))
            except Exception as e:  # pragma: no cover
                line["full_round"] = {"unavailable": repr(e)}
        if not args.no_aes128 and world == 1:
            try:
                line["aes128"] = aes128_probe(min(args.batch, 4))
            except Exception as e:  # pragma: no cover
                line["aes128"] = {"unavailable": repr(e)}
        if args.reference_order:
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            svc.sub_bytes_array(ct_in)
            a.record(); o2 = svc.sub_bytes_array(ct_in); b.record()
            torch.cuda.synchronize()
            ok = np.array_equal(ZetaEncoder.from_zeta(np.atleast_2d(eng.decrypt(o2, ctx.secret_key)), 256), sbox[data])
            line["reference_order"] = {"ms_per_step": a.elapsed_time(b), "keyswitches_per_ciphertext": 255,
                                       "blocks_per_s": blocks / (a.elapsed_time(b) * 1e-3), "bytes_ok": bool(ok)}
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"], _ = cpu_arm()
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"unavailable": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
