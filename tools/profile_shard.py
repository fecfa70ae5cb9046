#!/usr/bin/env python3
"""Workload for ncu: rank `r` of an 8-way sharded prove() at n = 2^16 (what one GPU of eight executes), three times."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
n, Q = 1 << 16, 8
srs = sb.SRS.new(7 * n, x, alpha)
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
buf = ctypes.create_string_buffer(int(L.sonic_shard_blob_size(Q)))
w = ctypes.c_uint64(0)
a = np.concatenate([c["aL"], c["aR"], c["aO"]])
for rep in range(3):
    capi.check(L.sonic_prove_shard_sink(srs._h, ch, a.ctypes.data, 0, None, rnd.ctypes.data, rank, world, buf, len(buf), ctypes.byref(w), None))
    print("shard", rank, world, {k: round(sb.last_timing_ms(k), 3) for k in ("total", "poly", "msm.sort", "msm.accumulate", "msm.reduce")}, "launches", sb.launch_count(), flush=True)
