#!/usr/bin/env python3
"""prove() at n = 2^16 for a list of option settings: `python tools/time_prove_modes.py acc_mode=2 acc_mode=2,aff_fused=1 ...`"""
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
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
buf = ctypes.create_string_buffer(int(L.sonic_proof_size(Q)))
w = ctypes.c_uint64(0)
ref = None
for spec in sys.argv[1:] or ["acc_mode=1"]:
    for kv in spec.split(","):
        k, v = kv.split("=")
        sb.set_option(k, int(v))
    best = None
    for _ in range(3):
        capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, buf, len(buf), ctypes.byref(w)))
        tm = {k: round(sb.last_timing_ms(k), 3) for k in ("total", "poly", "msm.sort", "msm.accumulate", "msm.reduce")}
        if best is None or tm["total"] < best["total"]:
            best = tm
    ref = ref or buf.raw
    print(spec, best, "same proof" if buf.raw == ref else "DIFFERENT PROOF", flush=True)
