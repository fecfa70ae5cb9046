#!/usr/bin/env python3
"""Workload for `ncu --set full` of the Fr-side and SRS kernels: SRS.new without window tables (one
k_fixed_base, two k_batch_affine), circuit load and ONE prove() at n = 2^log_n."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, Q = 1 << log_n, 8
sb.init(0)
sb.set_option("precompute", 0)
L = capi.lib()
x, alpha = synth.trapdoor()
srs = sb.SRS.new(7 * n, x, alpha)
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
out = ctypes.create_string_buffer(int(L.sonic_proof_size(Q)))
w = ctypes.c_uint64(0)
capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, out, len(out), ctypes.byref(w)))
print("proof ms", sb.last_timing_ms("total"), "poly", sb.last_timing_ms("poly"), flush=True)
