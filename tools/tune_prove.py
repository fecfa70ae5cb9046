#!/usr/bin/env python3
"""Option sweep for prove() at n = 2^16: python tools/tune_prove.py name=v1,v2 ..."""
import ctypes, os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sonic_b200 as sb
from sonic_b200 import capi, synth

log_n = 16
opts = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "log_n":
        log_n = int(v)
    else:
        opts[k] = [int(x) for x in v.split(",")]
n, Q = 1 << log_n, 8
sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
out = ctypes.create_string_buffer(int(L.sonic_proof_size(Q)))
w = ctypes.c_uint64(0)
srs_opts = {k: v for k, v in opts.items() if k.startswith("precompute")}
run_opts = {k: v for k, v in opts.items() if not k.startswith("precompute")}
ref = None
for sv in itertools.product(*srs_opts.values()) if srs_opts else [()]:
    for k, v in zip(srs_opts, sv):
        sb.set_option(k, v)
    srs = sb.SRS.new(7 * n, x, alpha)
    for rv in itertools.product(*run_opts.values()) if run_opts else [()]:
        for k, v in zip(run_opts, rv):
            sb.set_option(k, v)
        best = None
        for i in range(4):
            capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, out, len(out), ctypes.byref(w)))
            t = {k: sb.last_timing_ms(k) for k in ("total", "poly", "msm.sort", "msm.accumulate_kernel", "msm.accumulate", "msm.reduce")}
            if best is None or t["total"] < best["total"]:
                best = t
        if ref is None:
            ref = out.raw
        print(dict(zip(srs_opts, sv)), dict(zip(run_opts, rv)), {k: round(v, 2) for k, v in best.items()}, "same" if out.raw == ref else "DIFFERENT", flush=True)
    srs.free()
