#!/usr/bin/env python3
"""bench.py -- the Sonic prover hot path on N B200s (one process per GPU).

A step is one prove() at BASELINE.json's headline configuration (config 4: synthetic circuit,
n = 2^16 multiplication constraints, Q = 8 linear constraints, d = 7n, SRS resident in HBM).
metric = G1 MSM Mpoints/s = MSM terms of one proof (4Q+7 = 39 commitments/openings) / step time;
`prove_ms` (= ms_per_step) is reported alongside.

  value     : inputs (assignment, draws) already resident in HBM when the timed region starts
  e2e       : the same proofs through the public C ABI with HOST (pinned) buffers: host->device
              copies of the assignment and device->host of the proof inside the timed region
  roofline  : the bucket-accumulation kernel against the measured integer-multiply peak
  N > 1     : one proof sharded over the GPUs (equal runs of its MSM terms), one NCCL all-gather of
              the ~4 KB exchange records per proof; strong scaling.  Two modes are measured:
              "library"   rank 0 binds all N GPUs (sonic_init(devices, N)) and calls plain sonic_prove --
                          the drop-in path, `prove` stays one call in one process; the other torchrun
                          ranks wait.  This mode supplies the headline value and e2e.
              "processes" one process per GPU: sonic_prove_shard + torch.distributed all-gather + fold
  config5   : SRS.new at d = 2^22 (sharded over the GPUs + all-gather) and 64 proofs at n = 2^14 in one
              sonic_prove_batch call (8 per GPU at N = 8)

`--impl reference` times the reference's CPU algorithm (one double-and-add scalar
multiplication per term, CommitmentScheme.hs:26-29) as restated in oracle/csrc/sonic_ref.c
-- the Haskell original cannot be built here (no GHC) -- on all host threads.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CANON_WINDOW = 16                      # canonical c = 16 => W = 16 bucket insertions / point
LMAC_PER_MADD = 3000                   # canonical: 10 Fq multiplications x 300 LMAC
EXEC_LMAC_PER_MADD = 6 * 300 + 2 * 234 + 444   # executed by the XYZZ kernel: 6 mul, 2 half-product squarings, 1 fused a*b-c*d (2 products, 1 reduction)
EXEC_LMAC_PER_AFFINE_ADD = 5 * 300 + 234 + 60  # executed by the affine stage: prefix product, 1/d, running inverse, lambda, y3; lambda^2 as a half-product squaring; ~0.2 products of block trees
CANON_LMAC_PER_POINT = 16 * LMAC_PER_MADD


def msm_terms_per_proof(n: int, Q: int) -> int:
    """Terms of the 4Q+7 MSMs of one proof (A3/A4 sizes)."""
    r, t, s, c = 3 * n + 5, 7 * n + 9, 3 * n + 1, 2 * n + Q + 1
    main = r + t + 2 * (r - 1) + (t - 1)
    hsc = Q * (s + (s - 1)) + Q * ((s - 1) + (c - 1)) + (c - 1) + c
    return main + hsc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


_TABLES = {}


def cpu_sample(sample_terms: int, threads: int, seed: int):
    """Reference-algorithm MSM on `sample_terms` terms of the workload's distribution (uniform Fr
    coefficients against SRS bases g^{x^k} of the same trapdoor).  Returns (points_per_s, seconds)."""
    from oracle import cref
    from sonic_b200 import synth

    x, alpha = synth.trapdoor()
    d = sample_terms // 2
    if d not in _TABLES:                                                 # setup, not timed
        _TABLES[d] = cref.srs_new(d, x, alpha, threads=host_threads())
    pts = _TABLES[d][:96 * sample_terms]                                 # plain family, exponents -d ..
    scal = synth.fr_bytes_fast(seed, sample_terms).tobytes()
    t0 = time.perf_counter()
    cref.msm_naive(pts, scal, sample_terms, threads)
    dt = time.perf_counter() - t0
    return sample_terms / dt, dt


def run_reference(args, rank: int, world: int) -> None:
    """`--impl reference`: rank 0 alone times the CPU restatement; other ranks exit 0."""
    if rank != 0:
        return
    n, Q = 1 << args.log_n, args.Q
    threads = host_threads()
    sample = threads * 2048
    for _ in range(args.warmup):
        cpu_sample(sample, threads, 99)
    times = []
    for s in range(args.steps):
        _, dt = cpu_sample(sample, threads, 100 + s)
        times.append(dt)
    value = sample * len(times) / sum(times) / 1e6
    terms = msm_terms_per_proof(n, Q)
    # the all-threads figure moves with the box (16 or 32 host threads between two drivers' runs): fixed-thread
    # figures beside it make rounds comparable
    fixed = {}
    for th in (1, 16):
        if th > threads:
            continue
        cnt = th * 1024
        rate, _ = cpu_sample(cnt, th, 7)
        fixed[str(th)] = rate / 1e6
    line = {
        "impl": "reference", "metric": "G1 MSM Mpoints/s inside prove() at n=2^%d" % args.log_n, "value": value,
        "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "prove() n=2^%d Q=%d d=7n: MSM terms of one proof = %d; each step = reference-algorithm MSM "
                               "(double-and-add per term, CommitmentScheme.hs:26-29) on a %d-term sample" % (args.log_n, Q, terms, sample),
                   "extrapolated_prove_ms": 1e3 * terms / (value * 1e6)},
        "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": threads, "kind": "port", "fixed_threads_mpoints_per_s": fixed,
                         "sample": "%d uniform-Fr terms against SRS bases of the bench trapdoor, %d threads; C restatement "
                                   "(oracle/csrc/sonic_ref.c) -- the Haskell reference cannot be built here (no GHC)" % (sample, threads)},
        "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


GOLDEN_CONFIG4 = os.path.join(ROOT, "tests", "golden", "prove_config4.json")


def golden_sha(log_n: int, Q: int):
    """SHA-256 of the config-4 proof as the C restatement of the reference computed it (tools/gen_golden_large.py)."""
    if (log_n, Q) != (16, 8) or not os.path.exists(GOLDEN_CONFIG4):
        return None
    g = json.load(open(GOLDEN_CONFIG4))
    return g["proof_sha256"] if (g["n"], g["Q"], g["circuit_seed"], g["rnd_seed"]) == (1 << 16, 8, 4, 40) else None


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=16, help="log2 of the number of multiplication constraints")
    ap.add_argument("--Q", type=int, default=8)
    ap.add_argument("--no-sweep", action="store_true", help="skip the standalone MSM sweep (config 3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sweep-max", type=int, default=24)
    ap.add_argument("--sweep-tables", type=int, default=1, help="1: precomputed 20-bit window tables for the large sweep SRS (42 GB at d=2^23); 0: none")
    ap.add_argument("--no-config5", action="store_true", help="skip BASELINE config 5 (SRS.new d=2^22, then 64 proofs at n=2^14 over all GPUs)")
    ap.add_argument("--mode", default="both", choices=["both", "library", "processes"],
                    help="N > 1: 'library' = rank 0 drives all N GPUs through plain sonic_prove (sonic_init with N devices; the other torchrun "
                         "ranks wait), 'processes' = one process per GPU with sonic_prove_shard + NCCL all-gather, 'both' = measure both")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import hashlib

    import torch
    import torch.distributed as dist

    import sonic_b200 as sb
    from sonic_b200 import capi, synth
    from sonic_b200 import dist as sdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    idle_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idle_group = dist.new_group(backend="gloo")   # ranks that only wait must not spin a kernel on their GPU
    L = capi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload (host side, the same on every rank) ------------------------------------------------------
    n, Q = 1 << args.log_n, args.Q
    d = 7 * n
    x, alpha = synth.trapdoor()
    circ = synth.synthetic_circuit_bytes(n, Q, seed=4)
    nr = 2 * Q + 8
    rnd_ints = [v or 1 for v in synth.fr_ints(40, nr)]
    # pinned host buffers (what the Haskell shim would hand over)
    host_in = torch.empty(3 * n * 32, dtype=torch.uint8).pin_memory()
    host_in.numpy()[:] = np.concatenate([circ["aL"], circ["aR"], circ["aO"]])
    host_rnd = torch.empty(nr * 32, dtype=torch.uint8).pin_memory()
    host_rnd.numpy()[:] = np.frombuffer(synth.ints_to_bytes(rnd_ints), dtype=np.uint8)
    hin, hrnd = host_in.data_ptr(), host_rnd.data_ptr()
    proof_size = int(L.sonic_proof_size(Q))
    blob_size = int(L.sonic_shard_blob_size(Q))
    rec_size = int(L.sonic_shard_exchange_size(Q))
    terms = msm_terms_per_proof(n, Q)
    nm = 4 * Q + 7
    want_sha = golden_sha(args.log_n, Q)
    written = ctypes.c_uint64(0)
    STAGES = ("msm", "msm.sort", "msm.accumulate", "msm.accumulate_kernel", "msm.reduce", "poly", "total",
              "msm.window_bits", "msm.windows", "msm.terms", "msm.entries", "msm.chunk", "msm.buckets", "msm.jobs", "msm.affine")

    def check_proof(proof: bytes, what: str):
        if want_sha is not None and hashlib.sha256(proof).hexdigest() != want_sha:
            raise SystemExit("%s: the proof differs from tests/golden/prove_config4.json (the C restatement's proof of this "
                             "exact input): refusing to report a number for wrong bytes" % what)

    def setup(devices):
        """binds the library, builds the resident SRS and circuit, uploads the device-resident copies of the inputs"""
        sb.init(devices)
        t0 = time.perf_counter()
        srs = sb.SRS.new(d, x, alpha)
        wall = 1e3 * (time.perf_counter() - t0)
        srs_t = {"wall": wall, "device": sb.last_timing_ms("total"), "generate": sb.last_timing_ms("srs.generate"),
                 "allgather": sb.last_timing_ms("srs.allgather"), "points": 4 * d + 1}
        ch = ctypes.c_void_p()
        capi.check(L.sonic_circuit_load(n, Q, circ["wL"].ctypes.data, circ["wR"].ctypes.data, circ["wO"].ctypes.data,
                                        circ["cs"].ctypes.data, ctypes.byref(ch)))
        d_in, d_rnd = ctypes.c_void_p(), ctypes.c_void_p()
        capi.check(L.sonic_dev_alloc(3 * n * 32, ctypes.byref(d_in)))
        capi.check(L.sonic_dev_alloc(nr * 32, ctypes.byref(d_rnd)))
        capi.check(L.sonic_dev_upload(d_in, hin, 3 * n * 32))
        capi.check(L.sonic_dev_upload(d_rnd, hrnd, nr * 32))
        return srs, srs_t, ch, d_in, d_rnd

    def teardown(srs, ch, d_in, d_rnd):
        L.sonic_dev_free(d_in)
        L.sonic_dev_free(d_rnd)
        L.sonic_circuit_free(ch)
        srs.free()
        sb.shutdown()

    def timed(step, steps, sync_all):
        if sync_all:
            barrier()
        launches0 = sb.launch_count()
        capi.check(L.sonic_bench_mark(0))
        w0 = time.perf_counter()
        proof = None
        per_step = []
        for _ in range(steps):
            ts = time.perf_counter()
            proof = step()
            per_step.append(1e3 * (time.perf_counter() - ts))
        capi.check(L.sonic_bench_mark(1))
        ev_ms = L.sonic_bench_elapsed_ms(0, 1)
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - w0)
        if sync_all:
            barrier()
            ev_ms, wall_ms = max_over_ranks(ev_ms), max_over_ranks(wall_ms)
        return {"ev_ms": ev_ms, "wall_ms": wall_ms, "launches": sb.launch_count() - launches0, "proof": proof,
                "step_wall": {"min": min(per_step), "median": sorted(per_step)[len(per_step) // 2], "max": max(per_step)}}

    sampler = ClockSampler(local_rank)   # started before the warm-up so that NVML start-up is not in the timed region
    sampler.start()
    out = ctypes.create_string_buffer(max(proof_size, blob_size))
    proof_buf = ctypes.create_string_buffer(proof_size)

    # ========================================================================================================
    # Mode "processes" (and N = 1): one process per GPU; a proof is sharded with sonic_prove_shard_sink, the
    # exchange records meet in ONE NCCL all-gather on device buffers, every rank folds.
    # ========================================================================================================
    modes = {}
    imad_peak = None
    peaks = None
    sweep = []
    cpu = None
    run_processes = world == 1 or args.mode in ("both", "processes")
    run_library = world > 1 and args.mode in ("both", "library")
    if run_processes:
        srs, srs_t, ch, d_in, d_rnd = setup(local_rank)
        # roofline denominator: measured integer-multiply peak on this device
        peaks = [L.sonic_imad_peak_lmacs(v, 3000) for v in (0, 1, 2, 3)]
        imad_peak = max(peaks)
        if world > 1:
            part_t = torch.empty(rec_size, dtype=torch.uint8, device="cuda")
            gath_t = torch.empty(world * rec_size, dtype=torch.uint8, device="cuda")
        phase = {"shard": 0.0, "exchange": 0.0, "combine": 0.0, "calls": 0}

        def exchange_and_combine() -> bytes:
            t0 = time.perf_counter()
            dist.all_gather_into_tensor(gath_t, part_t)
            torch.cuda.current_stream().synchronize()
            t1 = time.perf_counter()
            capi.check(L.sonic_prove_combine_device(Q, world, gath_t.data_ptr(), out, proof_buf, proof_size, ctypes.byref(written)))
            t2 = time.perf_counter()
            phase["exchange"] += t1 - t0
            phase["combine"] += t2 - t1
            phase["calls"] += 1
            return proof_buf.raw

        def step_resident() -> bytes:
            if world == 1:
                capi.check(L.sonic_prove_device(srs._h, ch, d_in, d_rnd, hrnd, out, len(out), ctypes.byref(written)))
                return out.raw[:proof_size]
            t0 = time.perf_counter()
            capi.check(L.sonic_prove_shard_sink(srs._h, ch, d_in, 1, d_rnd, hrnd, rank, world, out, len(out), ctypes.byref(written), part_t.data_ptr()))
            phase["shard"] += time.perf_counter() - t0
            return exchange_and_combine()

        def step_e2e() -> bytes:
            if world == 1:
                capi.check(L.sonic_prove(srs._h, ch, hin, hin + n * 32, hin + 2 * n * 32, hrnd, out, len(out), ctypes.byref(written)))
                return out.raw[:proof_size]
            capi.check(L.sonic_prove_shard_sink(srs._h, ch, hin, 0, None, hrnd, rank, world, out, len(out), ctypes.byref(written), part_t.data_ptr()))
            return exchange_and_combine()

        if world > 1:
            for _ in range(20):               # NCCL sets its channels up lazily: keep that out of the timed steps
                dist.all_gather_into_tensor(gath_t, part_t)
            torch.cuda.synchronize()
        single = None
        if world > 1:
            # the sharded proof must be the single-GPU proof, byte for byte.  Computed BEFORE the warm-up:
            # the whole-proof call grows the workspace arena, and the call after it pays for resizing it
            capi.check(L.sonic_prove(srs._h, ch, hin, hin + n * 32, hin + 2 * n * 32, hrnd, out, len(out), ctypes.byref(written)))
            single = out.raw[:proof_size]
        for _ in range(args.warmup):
            p_res = step_resident()
        for _ in range(args.warmup):
            p_e2e = step_e2e()
        if p_res != p_e2e:
            raise SystemExit("resident and host-buffer proofs differ")
        if single is not None and single != p_res:
            raise SystemExit("sharded proof differs from the single-GPU proof")
        check_proof(p_res, "processes mode")
        for k in phase:
            phase[k] = 0.0 if k != "calls" else 0
        res = timed(step_resident, args.steps, True)
        stage = {k: sb.last_timing_ms(k) for k in STAGES}
        per_rank = None
        if world > 1:
            # every rank's view of the last timed step (device ms), to see the balance of the dealing
            keys = ("total", "poly", "msm.sort", "msm.accumulate", "msm.reduce", "msm.terms", "msm.jobs")
            mine = torch.tensor([stage[k] for k in keys], dtype=torch.float64, device="cuda")
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = {k: [round(float(t[i]), 3) for t in allr] for i, k in enumerate(keys)}
        e2e = timed(step_e2e, args.steps, True)
        if res["proof"] != e2e["proof"] or res["proof"] != p_res:
            raise SystemExit("proofs differ between steps")
        modes["processes"] = {
            "ms_per_step": res["ev_ms"] / args.steps, "e2e_ms_per_step": e2e["ev_ms"] / args.steps,
            "wall_ms_per_step": res["wall_ms"] / args.steps, "e2e_wall_ms_per_step": e2e["wall_ms"] / args.steps,
            "launches": res["launches"], "step_wall_ms_rank0": {"resident": res["step_wall"], "e2e": e2e["step_wall"]},
            "stage": stage, "per_rank": per_rank, "srs_new_ms": srs_t,
            "phases_ms_rank0": None if world == 1 or not phase["calls"] else {
                "shard_call": 1e3 * phase["shard"] / max(1, args.steps), "exchange": 1e3 * phase["exchange"] / phase["calls"],
                "combine": 1e3 * phase["combine"] / phase["calls"], "note": "host wall clock on rank 0 over the timed steps (exchange/combine: both timed regions)"},
            "d2h_bytes_per_step": int(ProveOutBytes(Q)) if world == 1 else blob_size + int(ProveOutBytes(Q)),
            "proof_sha": hashlib.sha256(p_res).hexdigest(),
        }

        # ---- standalone MSM sweep (config 3), scalars resident in HBM ----------------------------------
        if not args.no_sweep:
            top = max(16, min(args.sweep_max, 24))
            # sizes up to 2^20 run against an SRS small enough for the precomputed window tables
            # (d = 2^19: 3.2 GB of tables); larger sizes against a d = 2^(top-1) SRS
            small_d = 1 << 19
            small = srs if small_d <= d else sb.SRS.new(small_d, x, alpha)
            dd = 1 << (top - 1)
            # the large SRS gets 20-bit window tables (13 levels, 42 GB at d = 2^23, one-time cost in SRS.new):
            # 13 insertions per point into ONE set of 2^19 buckets instead of 13-15 sets and a Horner tail
            if dd > small_d:
                sb.set_option("precompute", 20 if args.sweep_tables else 0)
            big = small if dd <= small_d else sb.SRS.new(dd, x, alpha)
            sb.set_option("precompute", -1)
            for logn in range(16, top + 1, 2):
                N = 1 << logn
                for kind in ("uniform", "skewed"):
                    if kind == "skewed" and logn not in (20, top):
                        continue
                    use = small if N <= 2 * small_d else big
                    sc = synth.fr_bytes_fast(logn, N) if kind == "uniform" else synth.skewed_fr_bytes(logn, N)
                    lo, hi = sdist.slice_bounds(-(N // 2), N // 2, rank, world)
                    part = np.ascontiguousarray(sc[lo + N // 2:hi + N // 2])
                    dsc = ctypes.c_void_p()
                    capi.check(L.sonic_dev_alloc(part.nbytes, ctypes.byref(dsc)))
                    capi.check(L.sonic_dev_upload(dsc, part.ctypes.data, part.nbytes))
                    o48 = ctypes.create_string_buffer(96)

                    def one():
                        if world == 1:
                            capi.check(L.sonic_msm_g1_device(use._h, 0, lo, hi - lo, dsc, o48))
                            return o48.raw[:48]
                        capi.check(L.sonic_msm_g1_device_partial(use._h, 0, lo, hi - lo, dsc, o48))
                        parts = sdist.all_gather_bytes(o48.raw)
                        return sb.g1_sum(parts)
                    for _ in range(2):
                        one()
                    reps = 3
                    barrier()
                    capi.check(L.sonic_bench_mark(2))
                    for _ in range(reps):
                        one()
                    capi.check(L.sonic_bench_mark(3))
                    ms = max_over_ranks(L.sonic_bench_elapsed_ms(2, 3)) / reps
                    barrier()
                    # canonical work is quoted at c = 16 whatever window ran; executed work = the entries
                    # this rank's launch really inserted x the executed IMAD.WIDE per mixed addition; both against N x peak
                    entries = sb.last_timing_ms("msm.entries")
                    acc_ms = sb.last_timing_ms("msm.accumulate_kernel")
                    row_affine = bool(sb.last_timing_ms("msm.affine"))
                    sweep.append({"log2_n": logn, "scalars": kind, "ms": ms, "mpoints_per_s": N / (ms * 1e-3) / 1e6,
                                  "frac_of_imad_peak": N * CANON_LMAC_PER_POINT / (ms * 1e-3) / (imad_peak * world),
                                  "executed_frac_of_imad_peak": entries * world * (EXEC_LMAC_PER_AFFINE_ADD if row_affine else EXEC_LMAC_PER_MADD) / (ms * 1e-3) / (imad_peak * world),
                                  "bucket_stage": "affine" if row_affine else "xyzz",
                                  "accumulate_kernel_ms_rank0": acc_ms, "entries_rank0": entries,
                                  "stages_ms_rank0": {k: sb.last_timing_ms(k) for k in ("msm.sort", "msm.accumulate", "msm.reduce")},
                                  "window_bits": sb.last_timing_ms("msm.window_bits"), "precomputed_tables": bool(sb.last_timing_ms("msm.precomputed"))})
                    capi.check(L.sonic_dev_free(dsc))
            if big is not small:
                big.free()
            if small is not srs:
                small.free()

        # ---- CPU baseline (rank 0, N = 1 only): the reference algorithm on one core --------------------
        if rank == 0 and world == 1 and not args.no_cpu:
            sample = 1 << 15
            rate, dt = cpu_sample(sample, 1, 7)
            # context: a CPU bucket-method MSM on all host threads (not the reference's algorithm)
            from oracle import cref
            psample, pc = 1 << 16, 12
            x_, a_ = synth.trapdoor()
            if psample // 2 not in _TABLES:
                _TABLES[psample // 2] = cref.srs_new(psample // 2, x_, a_, threads=host_threads())
            pscal = synth.fr_bytes_fast(8, psample).tobytes()
            tp = time.perf_counter()
            cref.msm_pippenger(_TABLES[psample // 2][:96 * psample], pscal, psample, pc, host_threads())
            pdt = time.perf_counter() - tp
            cpu = {"value": rate / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
                   "pippenger_context": {"value": psample / pdt / 1e6, "unit": "Mpoints/s", "cores": host_threads(), "window_bits": pc,
                                         "sample": "%d terms, %.1f s; plain windowed bucket method in C, NOT the reference's algorithm" % (psample, pdt)},
                   "sample": "%d uniform-Fr terms, reference algorithm (double-and-add per term + fold, CommitmentScheme.hs:26-29) "
                             "in C (oracle/csrc/sonic_ref.c), %.1f s; the Haskell reference is single-threaded and cannot be built here" % (sample, dt),
                   "extrapolated_prove_ms": 1e3 * terms / rate, "host_threads_available": host_threads()}
        if world > 1 or not run_library:
            pass
        if world > 1:
            teardown(srs, ch, d_in, d_rnd)

    # ========================================================================================================
    # Mode "library" (N > 1): ONE process (rank 0) binds all N GPUs with sonic_init(devices, N) and calls the
    # plain entry points of the reference surface: sonic_srs_new (sharded + all-gather), sonic_prove (sharded
    # inside the library, one ncclAllGather, fold on device 0), sonic_prove_batch.  The other torchrun ranks wait.
    # ========================================================================================================
    config5 = None
    if world > 1:
        dist.barrier(group=idle_group)
    drive = rank == 0 and (run_library or world == 1)
    if run_library and rank == 0:
        srs, srs_t, ch, d_in, d_rnd = setup(list(range(world)))
        if imad_peak is None:
            peaks = [L.sonic_imad_peak_lmacs(v, 3000) for v in (0, 1, 2, 3)]
            imad_peak = max(peaks)

        def lib_resident() -> bytes:
            capi.check(L.sonic_prove_device(srs._h, ch, d_in, d_rnd, hrnd, out, len(out), ctypes.byref(written)))
            return out.raw[:proof_size]

        def lib_e2e() -> bytes:
            capi.check(L.sonic_prove(srs._h, ch, hin, hin + n * 32, hin + 2 * n * 32, hrnd, out, len(out), ctypes.byref(written)))
            return out.raw[:proof_size]

        for _ in range(args.warmup):
            p_res = lib_resident()
        for _ in range(args.warmup):
            p_e2e = lib_e2e()
        if p_res != p_e2e:
            raise SystemExit("library mode: resident and host-buffer proofs differ")
        check_proof(p_res, "library mode")
        if "processes" in modes and modes["processes"].get("proof_sha") not in (None, hashlib.sha256(p_res).hexdigest()):
            raise SystemExit("library-mode proof differs from the processes-mode proof")
        res = timed(lib_resident, args.steps, False)
        per_dev = {k: [round(sb.last_timing_ms(k, r), 3) for r in range(world)]
                   for k in ("total", "poly", "msm.sort", "msm.accumulate", "msm.reduce", "msm.terms", "msm.jobs")}
        stage = {k: sb.last_timing_ms(k) for k in STAGES}
        e2e = timed(lib_e2e, args.steps, False)
        if res["proof"] != e2e["proof"] or res["proof"] != p_res:
            raise SystemExit("library mode: proofs differ between steps")
        modes["library"] = {
            "ms_per_step": res["ev_ms"] / args.steps, "e2e_ms_per_step": e2e["ev_ms"] / args.steps,
            "wall_ms_per_step": res["wall_ms"] / args.steps, "e2e_wall_ms_per_step": e2e["wall_ms"] / args.steps,
            "launches": res["launches"], "step_wall_ms_rank0": {"resident": res["step_wall"], "e2e": e2e["step_wall"]},
            "stage": stage, "per_rank": per_dev, "srs_new_ms": srs_t, "d2h_bytes_per_step": int(ProveOutBytes(Q)),
            "note": "one process, sonic_init(devices, %d): device-0 CUDA events around the K calls; the proof leaves device 0 only after the "
                    "NCCL all-gather of every device's record, so the events bracket all devices' work" % world,
        }
    if drive and not args.no_config5:
        # ---- BASELINE config 5: SRS.new at d = 2^22 (sharded over the devices + all-gather), then 64 independent
        # proofs at n = 2^14 in one sonic_prove_batch call (whole proofs dealt round-robin, no exchange) -----------
        if world == 1:
            pass   # the library is still bound to device 0 from the processes mode
        d5, n5, total_proofs = 1 << 22, 1 << 14, 64
        t0 = time.perf_counter()
        srs5 = sb.SRS.new(d5, x, alpha)
        srs5_wall = 1e3 * (time.perf_counter() - t0)
        srs5_t = {k: sb.last_timing_ms(k) for k in ("total", "srs.generate", "srs.allgather")}
        gen_by_dev = [sb.last_timing_ms("srs.generate", r) for r in range(world)]
        c5 = synth.synthetic_circuit_bytes(n5, Q, seed=5)
        ch5 = ctypes.c_void_p()
        capi.check(L.sonic_circuit_load(n5, Q, c5["wL"].ctypes.data, c5["wR"].ctypes.data, c5["wO"].ctypes.data,
                                        c5["cs"].ctypes.data, ctypes.byref(ch5)))
        a5 = np.ascontiguousarray(np.tile(np.concatenate([c5["aL"], c5["aR"], c5["aO"]]), total_proofs))
        r5 = np.concatenate([np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(500 + i, nr)]), dtype=np.uint8) for i in range(total_proofs)]).copy()
        o5 = ctypes.create_string_buffer(proof_size * total_proofs)

        def batch():
            capi.check(L.sonic_prove_batch(srs5._h, ch5, total_proofs, a5.ctypes.data, r5.ctypes.data, o5, len(o5), ctypes.byref(written)))
        batch()                                   # warm-up: builds the window tables restricted to n = 2^14 on every device
        pre5 = bool(sb.last_timing_ms("msm.precomputed"))
        tb = time.perf_counter()
        batch()
        b_ms = 1e3 * (time.perf_counter() - tb)
        dev_ms = [sb.last_timing_ms("batch", r) for r in range(world)]
        # one proof of the batch against the single-call path
        capi.check(L.sonic_prove(srs5._h, ch5, c5["aL"].ctypes.data, c5["aR"].ctypes.data, c5["aO"].ctypes.data, r5.ctypes.data + 7 * nr * 32,
                                 out, len(out), ctypes.byref(written)))
        if out.raw[:proof_size] != o5.raw[7 * proof_size:8 * proof_size]:
            raise SystemExit("config 5: batch proof 7 differs from sonic_prove on the same inputs")
        npts5 = 4 * d5 + 1
        gen_ms = max(gen_by_dev) if gen_by_dev and max(gen_by_dev) > 0 else srs5_t["total"]
        fb_lmac = npts5 * (16 * LMAC_PER_MADD + 1800)      # canonical w = 16 plus the batched inversion
        config5 = {"n_gpus": world,
                   "srs_new_d_2pow22": {"wall_ms": srs5_wall, "device_ms": srs5_t["total"], "generate_ms_per_device": gen_by_dev,
                                        "allgather_ms": srs5_t["srs.allgather"], "points": npts5,
                                        "mpoints_per_s": npts5 / (srs5_t["total"] * 1e-3) / 1e6, "full_range_tables": False,
                                        "roofline_k_fixed_base": {"bound": "imad", "achieved": fb_lmac / (gen_ms * 1e-3) / 1e12,
                                                                  "peak": imad_peak * world / 1e12, "unit": "TLMAC/s",
                                                                  "frac": fb_lmac / (gen_ms * 1e-3) / (imad_peak * world),
                                                                  "note": "canonical (16 x 3000 + 1800) LMAC per point / slowest device's generate time, against N x the measured IMAD peak; "
                                                                          "generate = scalars + fixed-base table + k_fixed_base + k_batch_affine"}},
                   "batch_64_proofs_n_2pow14": {"wall_ms": b_ms, "device_ms_per_device": dev_ms, "proofs_per_s": total_proofs / (b_ms * 1e-3),
                                                "ms_per_proof_per_device": [v / max(1, len(range(r, total_proofs, world))) for r, v in enumerate(dev_ms)],
                                                "window_tables": "restricted to the 17n+23 exponents n = 2^14 reads (built by the first proof)" if pre5 else "none",
                                                "scaling": "weak (independent proofs dealt round-robin, no exchange)"}}
        L.sonic_circuit_free(ch5)
        srs5.free()
    if run_library and rank == 0:
        teardown(srs, ch, d_in, d_rnd)
    if world > 1:
        dist.barrier(group=idle_group)
    clocks = sampler.stop()

    if rank == 0:
        # headline mode: the in-library path is the drop-in (`prove` stays one call in one process); the one-process-per-GPU
        # numbers are reported beside it
        head = "library" if "library" in modes else "processes"
        m = modes[head]
        ms_per_step, e2e_ms_per_step = m["ms_per_step"], m["e2e_ms_per_step"]
        value = terms / (ms_per_step * 1e-3) / 1e6
        e2e_value = terms / (e2e_ms_per_step * 1e-3) / 1e6
        stage = m["stage"]
        # ---- roofline of the dominant kernel (bucket accumulation), from the last timed step on device 0 ----------
        acc_ms = stage["msm.accumulate_kernel"]
        shard_terms = stage["msm.terms"]                       # terms device 0's launch processed
        canon_lmac = shard_terms * CANON_LMAC_PER_POINT
        affine = bool(stage.get("msm.affine"))
        exec_lmac = stage["msm.entries"] * (EXEC_LMAC_PER_AFFINE_ADD if affine else EXEC_LMAC_PER_MADD)
        achieved = canon_lmac / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0
        traffic = None
        if world == 1 and args.log_n == 16:
            # measured under ncu on this configuration only (profiles/): not repeated where it was not profiled.  Affine mode:
            # the DRAM bytes of every kernel of the bucket stage of one proof (tools/ncu_affine_stage.sh)
            suffix = "_affine_traffic.json" if affine else "_accumulate_traffic.json"
            tpath = sorted([os.path.join(ROOT, "profiles", f) for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith(suffix)] or [""])[-1]
            if os.path.exists(tpath):
                try:
                    traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        roofline = {
            "bound": "imad",
            "kernel": "bucket stage in affine coordinates: k_aff_prefix + k_aff_inverses + k_aff_add per round, serial tail" if affine else "k_msm_accumulate_compact",
            "achieved": achieved, "peak": imad_peak / 1e12, "unit": "TLMAC/s",
            "frac": achieved / (imad_peak / 1e12) if imad_peak else None, "traffic": traffic,
            "note": "integer-multiply roofline (SURVEY.md 8d) of ONE device's launch: achieved = canonical LMAC (terms of that launch x ceil(255/16) x 3000) "
                    "/ kernel time; peak = max of four register-only IMAD microbenchmarks measured in this run on one device (of measured)"
                    + ("; frac can pass 1 here: the affine stage with batched inversions executes ~1 794 LMAC per insertion, fewer than the 3 000 "
                       "of the canonical mixed addition the algorithmic figure is defined on -- `executed.frac` is the multiplier-pipe utilisation" if affine else ""),
            "executed": {"tlmac_per_s": exec_lmac / (acc_ms * 1e-3) / 1e12 if acc_ms > 0 else 0.0,
                         "frac": (exec_lmac / (acc_ms * 1e-3)) / imad_peak if acc_ms > 0 and imad_peak else None,
                         "lmac_per_insertion": EXEC_LMAC_PER_AFFINE_ADD if affine else EXEC_LMAC_PER_MADD,
                         "window_bits": stage["msm.window_bits"], "windows": stage["msm.windows"], "entries": stage["msm.entries"]},
            "kernel_ms": acc_ms, "msm_ms": stage["msm"], "kernel_share_of_step": acc_ms / ms_per_step if ms_per_step else None,
            "whole_msm_frac": (canon_lmac / (stage["msm"] * 1e-3)) / imad_peak if stage["msm"] > 0 and imad_peak else None,
            "whole_step_frac_of_n_gpus": (terms * CANON_LMAC_PER_POINT / (ms_per_step * 1e-3)) / (imad_peak * world) if imad_peak else None,
            "imad_microbench_lmacs": {"mad.lo.cc+madc.hi": peaks[0], "mad.wide.u32": peaks[1], "mad.lo+mad.hi": peaks[2],
                                      "carry-chained rows (madc.lo.cc/madc.hi.cc x4)": peaks[3]},
            "hbm": {"algorithmic_gb_per_s": stage["msm.entries"] * 100.0 / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else None,
                    "peak_gb_per_s": hbm_peak, "note": "96 B base + 4 B entry per insertion; not the bound"},
        }
        par = ("1 GPU" if world == 1 else
               "%d GPUs, ONE process: sonic_init(devices, %d); sonic_prove deals equal runs of the proof's MSM terms to the devices (Fr side by ownership), "
               "1 ncclAllGather of %d B per device per proof, fold on device 0" % (world, world, rec_size) if head == "library" else
               "%d GPUs, one process each: sonic_prove_shard + 1 NCCL all-gather of %d B per rank per proof (device buffers) + fold on every rank" % (world, rec_size))
        line = {
            "metric": "G1 MSM Mpoints/s inside prove() at n=2^%d" % args.log_n, "value": value, "unit": "Mpoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "prove_ms": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "prove() on a synthetic circuit, n=2^%d mult constraints, Q=%d, d=7n=%d, SRS resident in HBM "
                                   "(BASELINE.json config 4); %d MSM terms in %d MSMs per proof" % (args.log_n, Q, d, terms, 4 * Q + 7),
                       "n": n, "Q": Q, "d": d, "msm_terms_per_proof": terms, "ntt_len": 1 << (7 * n + 9 - 1).bit_length(),
                       "parallelism": par, "mode": head,
                       "l2": "inputs larger than L2: the resident SRS is %.0f MB and is gathered at random every step" % ((4 * d + 2) * 96 / 1e6),
                       "srs_new_ms": m["srs_new_ms"],
                       "proof_pinned_to": "tests/golden/prove_config4.json (sha256 checked before any number is printed)" if want_sha else None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpoints/s", "ms_per_step": e2e_ms_per_step, "wall_ms_per_step": m["e2e_wall_ms_per_step"],
                    "h2d_bytes_per_step": (3 * n * 32 + nr * 32) * (world if head == "library" else 1), "d2h_bytes_per_step": m["d2h_bytes_per_step"],
                    "note": "host (pinned) buffers through sonic_prove; with N devices every device pulls its own copy of the assignment"},
            "gpu_launches": m["launches"],
            "wall_ms_per_step": m["wall_ms_per_step"],
            "step_wall_ms_rank0": m["step_wall_ms_rank0"],
            "stages_ms_last_step": {k: stage[k] for k in ("poly", "msm.sort", "msm.accumulate", "msm.reduce", "msm", "total")},
            "shard_stages_ms_per_rank": m["per_rank"],
            "modes": {k: {kk: vv for kk, vv in v.items() if kk not in ("stage",)} | {"value": terms / (v["ms_per_step"] * 1e-3) / 1e6,
                                                                                   "e2e_value": terms / (v["e2e_ms_per_step"] * 1e-3) / 1e6} for k, v in modes.items()},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "msm_sweep": sweep,
            "config5": config5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ProveOutBytes(Q: int) -> int:
    """bytes of the result buffer one proof brings back: (4Q+7) G1 + (2Q+3) Fr + status words (csrc/internal.h: ProveLayout)"""
    nm, nv = 4 * Q + 7, 2 * Q + 3
    return ((nm * 48 + nv * 32 + (3 * nm + 3) * 4) + 31) // 32 * 32


if __name__ == "__main__":
    main()
