#!/usr/bin/env python3
"""prove() at n = 2^16 (and a 1/8 shard of it) for several upper bounds of the accumulate chunk length."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
n, Q = 1 << 16, 8
srs = sb.SRS.new(7 * n, x, alpha)
print("srs_new ms", sb.last_timing_ms("total"), flush=True)
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
buf = ctypes.create_string_buffer(max(int(L.sonic_proof_size(Q)), int(L.sonic_shard_blob_size(Q))))
w = ctypes.c_uint64(0)
a = np.concatenate([c["aL"], c["aR"], c["aO"]])
KEYS = ("total", "poly", "msm.sort", "msm.accumulate", "msm.accumulate_kernel", "msm.reduce", "msm.chunk")
for cm in (64, 96, 128, 192, 256, 512):
    sb.set_option("chunk_max", cm)
    best = None
    for rep in range(3):
        capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, buf, len(buf), ctypes.byref(w)))
        tm = {k: round(sb.last_timing_ms(k), 3) for k in KEYS}
        if best is None or tm["total"] < best["total"]:
            best = tm
    print("whole proof chunk_max", cm, best, flush=True)
    for rank in (0, 3):
        best = None
        for rep in range(3):
            capi.check(L.sonic_prove_shard_sink(srs._h, ch, a.ctypes.data, 0, None, rnd.ctypes.data, rank, 8, buf, len(buf), ctypes.byref(w), None))
            tm = {k: round(sb.last_timing_ms(k), 3) for k in KEYS}
            if best is None or tm["total"] < best["total"]:
                best = tm
        print("  shard %d/8 chunk_max" % rank, cm, best, flush=True)
