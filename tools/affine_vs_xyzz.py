#!/usr/bin/env python3
"""The bucket stage in affine coordinates with batched inversions (acc_mode 2) against the XYZZ kernel (acc_mode 1):
prove() at n = 2^16 and the standalone MSM of 2^`top` points (20-bit window tables for the large SRS)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

top = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
out = {}
KEYS = ("total", "msm", "msm.sort", "msm.accumulate", "msm.accumulate_kernel", "msm.reduce", "msm.entries")
n, Q = 1 << 16, 8
srs = sb.SRS.new(7 * n, x, alpha)
c = synth.synthetic_circuit_bytes(n, Q, seed=4)
ch = ctypes.c_void_p()
capi.check(L.sonic_circuit_load(n, Q, c["wL"].ctypes.data, c["wR"].ctypes.data, c["wO"].ctypes.data, c["cs"].ctypes.data, ctypes.byref(ch)))
rnd = np.frombuffer(synth.ints_to_bytes([v or 1 for v in synth.fr_ints(40, 2 * Q + 8)]), dtype=np.uint8).copy()
buf = ctypes.create_string_buffer(int(L.sonic_proof_size(Q)))
w = ctypes.c_uint64(0)
proofs = {}
for mode, fused in ((1, 0), (2, 0), (2, 1)):
    sb.set_option("acc_mode", mode)
    sb.set_option("aff_fused", fused)
    best = None
    for _ in range(3):
        capi.check(L.sonic_prove(srs._h, ch, c["aL"].ctypes.data, c["aR"].ctypes.data, c["aO"].ctypes.data, rnd.ctypes.data, buf, len(buf), ctypes.byref(w)))
        tm = {k: round(sb.last_timing_ms(k), 3) for k in KEYS}
        if best is None or tm["total"] < best["total"]:
            best = tm
    proofs[(mode, fused)] = buf.raw
    out["prove n=2^16 acc_mode=%d aff_fused=%d" % (mode, fused)] = best
    print("prove", mode, fused, best, flush=True)
assert len(set(proofs.values())) == 1, "proofs differ between accumulation modes"
L.sonic_circuit_free(ch)
srs.free()
sb.set_option("precompute", 20)
big = sb.SRS.new(1 << (top - 1), x, alpha)
sb.set_option("precompute", -1)
N = 1 << top
sc = np.ascontiguousarray(synth.fr_bytes_fast(top, N))
dsc = ctypes.c_void_p()
capi.check(L.sonic_dev_alloc(sc.nbytes, ctypes.byref(dsc)))
capi.check(L.sonic_dev_upload(dsc, sc.ctypes.data, sc.nbytes))
o48 = ctypes.create_string_buffer(48)
pts = {}
for mode in (1, 2):
    sb.set_option("acc_mode", mode)
    sb.set_option("aff_fused", 0)
    best = None
    for _ in range(3):
        capi.check(L.sonic_msm_g1_device(big._h, 0, -(N // 2), N, dsc, o48))
        tm = {k: round(sb.last_timing_ms(k), 3) for k in KEYS}
        if best is None or tm["msm"] < best["msm"]:
            best = tm
    pts[mode] = o48.raw
    out["msm 2^%d acc_mode=%d" % (top, mode)] = best
    print("msm", top, mode, best, flush=True)
assert pts[1] == pts[2]
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02n_affine_vs_xyzz.json"), "w"), indent=1)
