#!/usr/bin/env python3
"""Device timings of prove() at n = 2^16 on ONE GPU: the whole proof, and every rank's share of a `world`-way
sharded proof run one after the other (what each GPU of `world` would execute)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
opts = [a.split("=") for a in sys.argv[3:]]
sb.init(0)
for k, v in opts:
    sb.set_option(k, int(v))
L = capi.lib()
x, alpha = synth.trapdoor()
n, Q = 1 << log_n, 8
srs = sb.SRS.new(7 * n, x, alpha)
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
buf = ctypes.create_string_buffer(max(int(L.sonic_proof_size(Q)), int(L.sonic_shard_blob_size(Q))))
w = ctypes.c_uint64(0)
a = np.concatenate([c["aL"], c["aR"], c["aO"]])
KEYS = ("total", "poly", "msm.sort", "msm.accumulate", "msm.accumulate_kernel", "msm.reduce", "msm.terms", "msm.jobs")


def best_of(call, reps=4):
    best = None
    for _ in range(reps):
        call()
        tm = {k: round(sb.last_timing_ms(k), 3) for k in KEYS}
        if best is None or tm["total"] < best["total"]:
            best = tm
    return best


print("whole", best_of(lambda: capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, buf, len(buf), ctypes.byref(w)))), flush=True)
worst = 0
for rank in range(world):
    tm = best_of(lambda: capi.check(L.sonic_prove_shard_sink(srs._h, ch, a.ctypes.data, 0, None, rnd.ctypes.data, rank, world, buf, len(buf), ctypes.byref(w), None)))
    worst = max(worst, tm["total"])
    print("rank %d/%d" % (rank, world), tm, flush=True)
print("slowest rank %.3f ms" % worst)
