#!/usr/bin/env python3
"""Latency of dependent point operations for a lone warp (the regime of the MSM's tail stages) and at full
occupancy, generic vs quad-cooperative formulas; then the standalone MSM and prove() timings per reduce_mode."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

sb.init(0)
L = capi.lib()
names = ["fq_mul", "g1_add", "g1_dbl", "g1_madd", "g1_add_quad", "g1_dbl_quad"]
out = {"latency_ns": {}}
for blocks, threads, tag in ((148, 32, "1 warp/SM"), (148 * 4, 32, "4 warps/SM (1 per scheduler)"), (148 * 8, 64, "16 warps/SM"), (148 * 8, 128, "32 warps/SM")):
    row = {}
    for op, nm in enumerate(names):
        row[nm] = round(L.sonic_selftest_latency_ns(op, 64 if op else 2048, blocks, threads), 1)
    out["latency_ns"][tag] = row
    print(tag, row, flush=True)
x, alpha = synth.trapdoor()
srs = sb.SRS.new(1 << 19, x, alpha)
out["msm"] = {}
for logn in (16, 18, 20):
    N = 1 << logn
    sc = np.ascontiguousarray(synth.fr_bytes_fast(logn, N))
    dsc = ctypes.c_void_p()
    capi.check(L.sonic_dev_alloc(sc.nbytes, ctypes.byref(dsc)))
    capi.check(L.sonic_dev_upload(dsc, sc.ctypes.data, sc.nbytes))
    o48 = ctypes.create_string_buffer(48)
    res = {}
    for mode in (2, 3):
        sb.set_option("reduce_mode", mode)
        best = None
        for rep in range(4):
            capi.check(L.sonic_msm_g1_device(srs._h, 0, -(N // 2), N, dsc, o48))
            tm = {k: round(sb.last_timing_ms(k), 4) for k in ("msm", "msm.sort", "msm.accumulate", "msm.reduce")}
            if best is None or tm["msm"] < best["msm"]:
                best = tm
        best["point"] = o48.raw.hex()[:16]
        res["reduce_mode=%d" % mode] = best
    assert res["reduce_mode=3"]["point"] == res["reduce_mode=2"]["point"]
    out["msm"]["2^%d" % logn] = res
    print(logn, res, flush=True)
    capi.check(L.sonic_dev_free(dsc))
srs.free()
# prove at n = 2^16 and n = 2^13 (the per-rank size of an 8-way shard is in between)
out["prove"] = {}
for log_n in (16, 13):
    n, Q = 1 << log_n, 8
    srs = sb.SRS.new(7 * n, x, alpha)
    c = synth.synthetic_circuit_bytes(n, Q, seed=4)
    ch = ctypes.c_void_p()
    capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
    rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
    buf = ctypes.create_string_buffer(int(L.sonic_proof_size(Q)))
    w = ctypes.c_uint64(0)
    res = {}
    for mode in (2, 3):
        sb.set_option("reduce_mode", mode)
        best = None
        for rep in range(4):
            capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, buf, len(buf), ctypes.byref(w)))
            tm = {k: round(sb.last_timing_ms(k), 4) for k in ("total", "poly", "msm.sort", "msm.accumulate", "msm.reduce")}
            if best is None or tm["total"] < best["total"]:
                best = tm
        best["proof"] = buf.raw.hex()[:16]
        res["reduce_mode=%d" % mode] = best
    assert res["reduce_mode=3"]["proof"] == res["reduce_mode=2"]["proof"]
    out["prove"]["n=2^%d" % log_n] = res
    print(log_n, res, flush=True)
    L.sonic_circuit_free(ch)
    srs.free()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "latency_probe.json"), "w"), indent=1)
