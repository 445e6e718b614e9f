#!/usr/bin/env python
"""Benchmark of the homomorphic-AES hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--states G]
                    [--scaling weak|strong --total-states T] [--no-extras]

Workload (BASELINE.json configs[4]; the metric "homomorphic AES-128 blocks/sec"): full AES-128 -- AddRoundKey_0 and
ten rounds (SubBytes, ShiftRows, MixColumns, AddRoundKey; the functions of the reference's round drivers
/root/reference/test_all_process.py:21-48 and /root/reference/new.py:186-227,248-261), round keys from the clear
FIPS-197 key schedule encrypted bit by bit -- at N = 2^16 on G *states* per GPU.  One state = 8192 AES blocks in 32
bit-plane ciphertexts (aes_fhe_b200/services/aes_bits.py); the state is refreshed by a bit bootstrap wherever the
levels run out (aes_fhe_b200/bootstrap.py::bootstrap_bits; six times per AES-128: `config.refresh_before_rounds`).  Parameter set: the default bootstrappable engine, 26 levels at a
44-bit scale, 27 + 7 limbs, log2(PQ) = 1504 (inside the 1553-bit bound for the sparse secret, params.py).

A "step" is one AES-128 pass over the G states of a rank.  `value` = AES blocks per second with the input
ciphertexts and the encrypted round keys resident in HBM (CUDA events, max over ranks); `e2e` = the same from host
buffers: the raw block bytes go to the GPU (H2D inside the timed region), are bit-sliced, encoded and encrypted
there, run through AES-128, come back to the key owner as ciphertexts (NCCL gather when N > 1), are decrypted and
decoded on the GPU and returned as bytes (D2H).  Every decoded block is compared with plain AES (FIPS-197
Appendix B is block 0).

Multi-GPU: states are sharded over the ranks with no data-path collective ("weak": G states per rank; "strong":
--total-states split over the ranks).  Rank 0 alone generates the keys and keeps the secret; public,
relinearisation, conjugation and all Galois keys are broadcast once over NCCL (`startup`), result ciphertexts are
gathered to rank 0.

`--impl reference` / `cpu_baseline`: the reference's arithmetic lives in the closed `desilofhe` wheel (absent), so
the CPU arm is the oracle port (oracle/refmod.cpp, OpenMP, all host cores).  A whole AES-128 at N = 2^16 takes hours
there, so each step times a bounded SAMPLE of the same arithmetic at N = 2^16 on the same parameter set (relinearised
products and rotations at the top, middle and bottom of the chain), measures seconds per length-N transform row, and
scales by the transform rows of one AES-128 state (tests/golden/aes_bits_work.json, counted by the library's
fhe_ntt_row_count at N = 2^16).  The extrapolation is stated in `cpu_baseline.sample`.
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

METRIC = "homomorphic AES-128 blocks/sec"
UNIT = "blocks/s"
LOG_N = 16
AES_LEVEL = 26
KEY_HEX = "2b7e151628aed2a6abf7158809cf4f3c"         # FIPS-197 Appendix B
PT_HEX = "3243f6a8885a308d313198a2e0370734"
CT_HEX = "3925841d02dc09fbdc118597196a0b32"


def aes_params():
    from aes_fhe_b200.params import LOG_PQ_BUDGET_SPARSE, make_params
    return make_params(LOG_N, AES_LEVEL, scale_bits=44, log_pq_budget=LOG_PQ_BUDGET_SPARSE)


def _peak_gbs():
    try:
        return json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def _golden_work():
    try:
        return json.load(open(ROOT / "tests" / "golden" / "aes_bits_work.json"))
    except Exception:
        return None


def host_threads() -> int:
    """cores this process may use -- never omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def rank_blocks(rank: int, n_blocks: int) -> np.ndarray:
    """synthetic plaintext of a rank: seeded uniform bytes; block 0 of rank 0 is FIPS-197 Appendix B"""
    b = np.random.default_rng(1000 + rank).integers(0, 256, (n_blocks, 16), dtype=np.uint8)
    if rank == 0:
        b[0] = np.frombuffer(bytes.fromhex(PT_HEX), np.uint8)
    return b


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


class StageTimer:
    """CUDA events at the stage boundaries of AESBitService (its `timer` hook), on the launching stream"""

    def __init__(self):
        import torch
        self.torch = torch
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks = [("start", e)]

    def __call__(self, name):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks.append((name, e))

    def ms(self):
        self.torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


# --------------------------------------------------------------------------- CPU arm
class CpuSample:
    """The bounded CPU sample: on ONE ciphertext at N = 2^16, same parameter set, oracle backend with all host
    cores -- two relinearised products and two rotations each at levels 26, 17 and 8.  `run()` returns
    (seconds, transform rows)."""

    LEVELS = (26, 17, 8)

    def __init__(self, threads: int):
        from oracle.refmod import RefBackend, build
        build()
        from aes_fhe_b200.engine import Engine
        P = aes_params()
        self.be = RefBackend(P, threads=threads)
        eng = Engine(_params=P, _backend=self.be, use_bootstrap=True)
        sk = eng.create_secret_key()
        pk = eng.create_public_key(sk)
        self.rlk = eng.create_relinearization_key(sk)
        self.rot = eng.create_fixed_rotation_key(sk, -eng.slot_count // 4)          # the ShiftRows rotation by one column
        self.eng, self.sk = eng, sk
        rng = np.random.default_rng(0)
        self.v = rng.choice([-1.0, 1.0], eng.slot_count)
        self.cts = {lvl: eng.encrypt(self.v, pk, level=lvl) for lvl in self.LEVELS}
        self.threads = self.be.threads

    def run(self):
        eng = self.eng
        r0 = self.be.ntt_row_count()
        t0 = time.perf_counter()
        outs = []
        for lvl in self.LEVELS:
            ct = self.cts[lvl]
            a = eng.multiply(ct, ct, self.rlk)
            b = eng.multiply(a, a, self.rlk)
            c = eng.rotate(eng.rotate(ct, self.rot), self.rot)
            outs.append((b, c))
        dt = time.perf_counter() - t0
        rows = self.be.ntt_row_count() - r0
        return dt, rows, outs

    def check(self, outs):
        for b, c in outs:
            assert np.abs(self.eng.decrypt(b, self.sk) - 1.0).max() < 1e-4          # (+-1)^4
            assert np.abs(self.eng.decrypt(c, self.sk) - np.roll(self.v, -self.eng.slot_count // 2)).max() < 1e-4


def cpu_baseline(samples: int, rows_per_state: int, blocks_per_state: int):
    cs = CpuSample(host_threads())
    times, rows = [], 0
    outs = None
    for _ in range(samples):
        dt, rows, outs = cs.run()
        times.append(dt)
    cs.check(outs)
    s_per_row = float(np.median(times)) / rows
    t_state = s_per_row * rows_per_state
    return {"value": blocks_per_state / t_state, "unit": UNIT, "cores": cs.threads, "kind": "port", "extrapolated": True,
            "sample": f"{samples} x [2 relinearised products + 2 rotations at each of levels 26/17/8, one ciphertext, N=2^16, "
                      f"27+7 limbs] = {rows} transform rows in {float(np.median(times)):.2f}s (median) on {cs.threads} threads "
                      f"(oracle/refmod.cpp) -> {s_per_row * 1e6:.1f} us per row; one AES-128 state (8192 blocks) = "
                      f"{rows_per_state} rows -> {t_state:.0f}s extrapolated",
            "seconds_per_transform_row": s_per_row, "rows_per_state_aes128": rows_per_state}, times


def reference_arm(args):
    work = _golden_work()
    if work is None:
        print(json.dumps({"impl": "reference", "unavailable": "tests/golden/aes_bits_work.json missing"}))
        return
    rows_state, bps = int(work["ntt_rows_per_state_aes128"]), int(work["blocks_per_state_at_2_16"])
    cs = CpuSample(host_threads())
    for _ in range(args.warmup):
        cs.run()
    times, rows, outs = [], 0, None
    for _ in range(args.steps):
        dt, rows, outs = cs.run()
        times.append(dt)
    cs.check(outs)
    ms_step = float(np.mean(times)) * 1e3
    s_per_row = float(np.mean(times)) / rows
    v = bps / (s_per_row * rows_state)
    cb = {"value": v, "unit": UNIT, "cores": cs.threads, "kind": "port", "extrapolated": True,
          "sample": f"each step = 2 relinearised products + 2 rotations at each of levels 26/17/8 on one ciphertext at N=2^16 "
                    f"({rows} transform rows, {ms_step:.0f} ms mean on {cs.threads} threads, oracle/refmod.cpp); value = 8192 blocks / "
                    f"({rows_state} rows of one AES-128 state x {s_per_row * 1e6:.1f} us per row)",
          "seconds_per_transform_row": s_per_row, "rows_per_state_aes128": rows_state}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64-exact-int",
                      "data": "synthetic", "config": workload_config(args, 1), "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------- extras (configs[1])
def subbytes_extra(batch: int = 16, steps: int = 2):
    """BASELINE configs[1] (last round's headline): the reference's zeta_256 SubBytes polynomial pair
    (sbox/sbox_service.py:116-138) in the 23-key-switch BSGS schedule, N = 2^16, max_level 22"""
    import torch
    from aes_fhe_b200.params import make_params
    from aes_fhe_b200.services.engine_context import EngineContext
    from aes_fhe_b200.services.sbox_service import SBoxService, AES_SBOX
    from aes_fhe_b200.services.xor_service import ZetaEncoder
    ctx = EngineContext(signature=2, max_level=22, mode="parallel",
                        _engine_kwargs=dict(_params=make_params(LOG_N, 22)), rotation_steps=[])
    svc = SBoxService(ctx)
    sc = ctx.engine.slot_count
    rng = np.random.default_rng(0)
    data = np.stack([np.tile(np.arange(256, dtype=np.uint8), sc // 256 + 1)[:sc]] +
                    [rng.integers(0, 256, sc, dtype=np.uint8) for _ in range(batch - 1)])
    ct = ctx.engine.encrypt(ZetaEncoder.to_zeta(data, 256), ctx.public_key)
    out = svc.sub_bytes_array_bsgs(ct)
    ok = np.array_equal(ZetaEncoder.from_zeta(np.atleast_2d(ctx.engine.decrypt(out, ctx.secret_key)), 256),
                        np.array(AES_SBOX, dtype=np.uint8)[data])
    svc.sub_bytes_array_bsgs(ct)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        svc.sub_bytes_array_bsgs(ct)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"workload": "configs[1]: SubBytes (sbox_hi x sbox_lo zeta_256 polynomials, BSGS, 23 key switches), N=2^16, L=22",
            "batch": batch, "ms_per_step": ms, "blocks_per_s": batch * (sc // 16) / (ms * 1e-3), "bytes_equal_sbox": bool(ok)}


def workload_config(args, world):
    P = aes_params()
    G = states_per_rank(args, world)
    return {"workload": "configs[4]: full AES-128 (AddRoundKey_0 + 10 rounds, encrypted FIPS-197 round keys), bit-sliced, "
                        "N=2^16, 26 levels, a bit bootstrap of the state wherever the levels run out (six per AES-128)",
            "states_per_gpu": G, "states_per_call": min(G, args.states_per_call), "blocks_per_state": (1 << (LOG_N - 1)) // 4, "ciphertexts_per_state": 32,
            "limbs_q_p_alpha_dnum": [P.n_q, P.n_p, P.alpha, P.dnum], "log2_pq": round(P.log_pq, 1), "scale_bits": 44,
            "l2": "working set per step (tens of GB of ciphertexts and keys) exceeds the 126 MB L2",
            "sharding": "independent states per rank, no data-path collective; keys broadcast once from rank 0, "
                        "result ciphertexts gathered to rank 0",
            "scaling_mode": args.scaling}


def states_per_rank(args, world):
    if args.scaling == "strong":
        if args.total_states % world:
            raise SystemExit("--total-states must be a multiple of the number of GPUs")
        return args.total_states // world
    return args.states


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--states", type=int, default=2, help="states (8192 blocks each) per GPU per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-states", type=int, default=8, help="strong scaling: states in the whole job")
    ap.add_argument("--states-per-call", type=int, default=2,
                    help="a rank's states go through the pipeline in calls of at most this many (HBM: ~50 GB of "
                         "temporaries per two states); the step is the sum of the calls")
    ap.add_argument("--fresh-level", type=int, default=None,
                    help="level the input state is encrypted at (default: the lowest level with the fewest refreshes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1] SubBytes extra and the kernel rooflines")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from aes_fhe_b200.engine import Engine
    from aes_fhe_b200.services.aes_bits import AESBitService
    from aes_fhe_b200.services.key_expansion import expand_key
    from aes_fhe_b200.services.xor_service import EngineWrapper, XORConfig
    from aes_fhe_b200.sharding import distribute_keys, gather_ciphertexts
    from oracle import aes_plain as A          # the byte checker, outside every timed region

    P = aes_params()
    G = states_per_rank(args, world)
    config = workload_config(args, world)
    t_setup = time.perf_counter()
    startup = None
    if rank == 0:
        # the key owner: OS-entropy randomness (no seed), all keys of the service issued before they are shipped
        w = EngineWrapper(XORConfig(device_id=local), _engine_kwargs=dict(_params=P, device_codec=True), rotation_steps=[])
        engine = w.engine
        svc = AESBitService(w)
        svc.prepare_keys()
    else:
        w, svc = None, None
        engine = Engine(_params=P, use_bootstrap=True, device_codec=True, device_id=local)
    if world > 1:
        ctx, startup = distribute_keys(engine, w, src=0)
        if rank:
            svc = AESBitService(ctx)
            svc.prepare_keys()
        startup = {"broadcast_keys": startup["keys"], "broadcast_bytes": startup["bytes"],
                   "broadcast_seconds": startup["seconds"], "broadcast_gbs": startup["gbs"],
                   "secret_key": "rank 0 only", "transport": "ncclBroadcast per key over NVLink/NVSwitch"}
    be = engine.backend
    Bs = svc.Bs
    key = bytes.fromhex(KEY_HEX)
    rks = expand_key(key)
    blocks = rank_blocks(rank, G * Bs)
    # the input is encrypted at the lowest level that gives the fewest refreshes (26 of 26: three rounds run on the fresh
    # levels, six bit bootstraps instead of ten), each round key at exactly the level where it is multiplied in
    fresh = args.fresh_level if args.fresh_level is not None else svc.best_fresh_level()
    plan = svc.plan_levels(fresh)
    config["fresh_level"] = fresh
    config["refresh_before_rounds"] = plan["refresh_before_rounds"]
    # a rank's G states go through the pipeline in calls of at most --states-per-call states
    calls = [(lo, min(G, lo + args.states_per_call)) for lo in range(0, G, args.states_per_call)]
    sts = [svc.encrypt_state(blocks[lo * Bs:hi * Bs], level=fresh) for lo, hi in calls]
    rkeys = {n: svc.encrypt_round_keys(key, n, plan) for n in {hi - lo for lo, hi in calls}}     # round 10 at half amplitude
    pinned = [torch.from_numpy(blocks[lo * Bs:hi * Bs]).pin_memory() for lo, hi in calls]

    def step():
        return [svc.encrypt_blocks(st, key, round_keys=rkeys[hi - lo]) for st, (lo, hi) in zip(sts, calls)]

    def collect(outs):
        """result ciphertexts on the key owner -> decoded block bytes per rank (None on other ranks)"""
        per_call = []
        for out in outs:
            parts = gather_ciphertexts(engine, out, dst=0) if world > 1 else [out]
            if not rank:
                per_call.append([svc.decrypt_state_device(p).numpy() for p in parts])
        if rank:
            return None
        return [np.concatenate([c[r] for c in per_call]) for r in range(world)]

    def e2e_step():
        outs = []
        for pin, (lo, hi) in zip(pinned, calls):
            ct = svc.encrypt_state_device(pin, level=fresh)                       # H2D of the raw block bytes inside
            outs.append(svc.encrypt_blocks(ct, key, round_keys=rkeys[hi - lo]))
        return collect(outs)                                                      # gather + decrypt + D2H of the bytes inside

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def verify(decoded):
        if rank:
            return True
        ok = True
        for r, got in enumerate(decoded):
            ok &= bool(np.array_equal(got, A.encrypt_blocks(rank_blocks(r, G * Bs), key)))
        return ok and decoded[0][0].tobytes().hex() == CT_HEX

    # correctness of what is being timed (also the first warm-up)
    out = step()
    ok_resident = verify(collect(out))
    if not ok_resident:
        raise SystemExit("AES-128 output differs from plain AES / FIPS-197")
    slot_err = None
    if rank == 0:                             # north_star check 2: decoded slots against +-1 (rank 0's first call)
        lo, hi = calls[0]
        want = A.encrypt_blocks(blocks[lo * Bs:hi * Bs], key)
        slot_err = float(np.abs(svc.decrypt_slots(out[0]) - (1.0 - 2.0 * svc.pack_bits(want))).max())
    setup_s = time.perf_counter() - t_setup
    for _ in range(max(0, args.warmup - 1)):
        step()

    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    l0, r0 = be.launch_count(), be.ntt_row_count()
    c0, b0 = dict(engine.op_counts), svc.refreshes
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.nvtx.range_push("timed_region")
    ev[0].record()
    for _ in range(args.steps):
        out = step()
    ev[1].record()
    torch.cuda.nvtx.range_pop()
    sync_all()
    launches = be.launch_count() - l0
    rows_step = (be.ntt_row_count() - r0) // args.steps
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    clocks = sampler.stop()
    counts = {k: (v - c0.get(k, 0)) // args.steps for k, v in engine.op_counts.items() if v - c0.get(k, 0)}
    boots = (svc.refreshes - b0) // args.steps

    # per-stage device time of one more pass (events at the stage boundaries)
    svc.timer = StageTimer()
    out = step()
    stages = svc.timer.ms()
    svc.timer = None

    # end to end from host buffers through the public API
    n_e2e = max(1, args.steps // 4)
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        dec = e2e_step()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
    ok_e2e = verify(dec)

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    n_blocks = world * G * Bs
    value = n_blocks / (ms * 1e-3)

    if rank == 0:
        peak, peak_kind = _peak_gbs()
        work = _golden_work()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64-exact-int", "data": "synthetic", "config": config,
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": n_blocks / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(world * blocks.nbytes), "d2h_bytes_per_step": int(world * blocks.nbytes),
                        "steps_timed": n_e2e, "bytes_equal_plain_aes": bool(ok_e2e),
                        "round_keys": "encrypted once, resident (the client's key does not change per step)"},
                "ms_per_round": ms / 10, "ms_per_state": ms / G,
                "stage_ms": {**stages, "note": "device time of one pass on rank 0, CUDA events at stage boundaries; refresh = "
                                               "bit bootstraps of the whole state (config.refresh_before_rounds)"},
                "bytes_equal_plain_aes": bool(ok_resident), "fips197_appendix_b": True,
                "bootstrapped_ciphertexts_per_step": int(boots), "bootstrapped_ciphertexts_per_2048_blocks": boots / G / 4,
                "max_slot_error": slot_err, "batched_op_calls_per_step": counts, "ntt_rows_per_state": int(rows_step // G),
                "ntt_rows_per_state_golden": None if work is None else work["ntt_rows_per_state_aes128"],
                "setup_seconds": setup_s, "security": engine.security,
                "hbm_peak_allocated_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        if startup is not None:
            line["startup"] = startup
        if not args.no_extras:
            line.update(kernel_rooflines(be, P, G, peak, peak_kind, clocks))
        else:
            line["roofline"] = None
        print_line = line
    del sts, out
    if rank == 0 and world == 1 and not args.no_extras:
        del rkeys, svc
        torch.cuda.empty_cache()
        try:
            print_line["subbytes"] = subbytes_extra()
        except Exception as e:  # pragma: no cover
            print_line["subbytes"] = {"unavailable": repr(e)}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            work = _golden_work()
            try:
                rows_state = int(work["ntt_rows_per_state_aes128"]) if work else int(print_line["ntt_rows_per_state"])
                print_line["cpu_baseline"], _ = cpu_baseline(3, rows_state, Bs)
            except Exception as e:  # pragma: no cover
                print_line["cpu_baseline"] = {"unavailable": repr(e)}
        print(json.dumps(print_line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_rooflines(be, P, G, peak, peak_kind, clocks):
    """The dominant kernel (forward NTT, both radix passes in one chained launch) and the key inner product, each
    timed alone with CUDA events on the launching stream, L2 flushed between launches, at the row counts of one
    key switch of a bootstrap batch at the top of the chain."""
    import torch
    nq, K = P.n_q, P.n_p
    batch = 16 * G
    beta, ne = P.digits_at(nq), nq + K
    rows = batch * beta * ne
    x = torch.randint(0, 2 ** 39, (batch * beta, ne, P.n), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        be._call("fhe_ntt_fwd", be._ptr(x), x.shape[0], nq, K)
    kt = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); be._call("fhe_ntt_fwd", be._ptr(x), x.shape[0], nq, K); b.record()
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b))
    k_ms = float(np.median(kt))
    alg_bytes = 2 * rows * P.n * 8                      # read once + write once per limb
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        traffic = json.load(open(ROOT / "profiles" / "ntt_traffic.json"))["dram_bytes_per_row"] * rows
        traffic_src = "static: ncu dram__bytes of the same kernel (profiles/ntt_traffic.json) x rows"
    except Exception:
        pass
    # FP64 work of the transform: N/2 log2 N butterflies x 8 FP64 instructions + 1.1 N (conversions, final reduction)
    fp64_ops = rows * (P.n // 2 * P.log_n * 8 + 1.1 * P.n)
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
    fp64_peak = 64 * 148 * sm_mhz * 1e6 / 1e9           # G instr/s: 64 DFMA-class per clock per SM (tools/ubench/pipes.cu)
    del x
    ext = torch.randint(0, 2 ** 39, (batch, beta, ne, P.n), dtype=torch.int64, device="cuda")
    dd = torch.randint(0, 2 ** 39, (1, batch, nq, P.n), dtype=torch.int64, device="cuda")
    acc = torch.empty(2, batch, ne, P.n, dtype=torch.int64, device="cuda")
    ksk = torch.randint(0, 2 ** 39, (P.dnum, 2, 1, ne, P.n), dtype=torch.int64, device="cuda")
    kt2 = []
    for it in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); be._call("fhe_ks_inner", be._ptr(acc), be._ptr(ext), be._ptr(dd), be._ptr(ksk), nq, batch); b.record()
        torch.cuda.synchronize()
        if it >= 3:
            kt2.append(a.elapsed_time(b))
    ks_ms = float(np.median(kt2))
    ks_bytes = (2 * beta * ne + batch * (beta * ne + 2 * ne)) * P.n * 8
    return {
        "roofline": {"bound": "hbm", "limiter": "fp64 pipe (see roofline_fp64)", "kernel": "ntt_fwd_chained (both radix passes in one launch)",
                     "rows_per_launch": rows, "us_per_launch": k_ms * 1e3, "us_per_row": k_ms * 1e3 / rows,
                     "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": int(alg_bytes)},
        "roofline_fp64": {"kernel": "ntt_fwd_chained", "achieved": fp64_ops / (k_ms * 1e-3) / 1e9, "peak": fp64_peak,
                          "unit": "G FP64 instr/s", "frac": fp64_ops / (k_ms * 1e-3) / 1e9 / fp64_peak,
                          "peak_kind": f"64 per clk per SM x 148 SMs x {sm_mhz:.0f} MHz (median SM clock of the timed region)",
                          "ops_per_row": fp64_ops / rows,
                          # static count of the compiled kernel (profiles/r02_sass_histogram.md): 1180 FP64 instructions per
                          # thread of 16 coefficients, twiddle products and reductions included
                          "frac_counting_all_fp64_instructions": rows * (1180 / 16) * P.n / (k_ms * 1e-3) / 1e9 / fp64_peak,
                          "diagnostic_builds": "profiles/r02_ntt_experiments.md: data movement costs 17 % of the kernel"},
        "roofline_keyswitch_inner": {"kernel": "k_ks_inner", "bound": "hbm", "achieved": ks_bytes / (ks_ms * 1e-3) / 1e9,
                                     "peak": peak, "unit": "GB/s", "frac": ks_bytes / (ks_ms * 1e-3) / 1e9 / peak,
                                     "us": ks_ms * 1e3, "algorithmic_bytes": int(ks_bytes)}}


if __name__ == "__main__":
    main()
