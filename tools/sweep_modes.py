#!/usr/bin/env python3
"""Standalone MSMs 2^20..2^24 (uniform and skewed scalars) per accumulation mode: acc_mode 1 (XYZZ) against 3 (automatic) or the mode given as second argument."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

top = int(sys.argv[1]) if len(sys.argv) > 1 else 24
other = int(sys.argv[2]) if len(sys.argv) > 2 else 3   # the mode compared with XYZZ: 3 automatic, 2 affine always
sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
small = sb.SRS.new(1 << 19, x, alpha)
sb.set_option("precompute", 20)
big = sb.SRS.new(1 << (top - 1), x, alpha)
sb.set_option("precompute", -1)
for logn in (20, 22, top):
    N = 1 << logn
    for kind in ("uniform", "skewed"):
        use = small if N <= (1 << 20) else big
        sc = np.ascontiguousarray(synth.fr_bytes_fast(logn, N) if kind == "uniform" else synth.skewed_fr_bytes(logn, N))
        dsc = ctypes.c_void_p()
        capi.check(L.sonic_dev_alloc(sc.nbytes, ctypes.byref(dsc)))
        capi.check(L.sonic_dev_upload(dsc, sc.ctypes.data, sc.nbytes))
        o48 = ctypes.create_string_buffer(48)
        res, pts = {}, {}
        for mode in (1, other):
            sb.set_option("acc_mode", mode)
            best = None
            for _ in range(3):
                capi.check(L.sonic_msm_g1_device(use._h, 0, -(N // 2), N, dsc, o48))
                tm = {k: round(sb.last_timing_ms(k), 3) for k in ("msm", "msm.sort", "msm.accumulate", "msm.reduce", "msm.affine")}
                if best is None or tm["msm"] < best["msm"]:
                    best = tm
            res[mode], pts[mode] = best, o48.raw
        assert pts[1] == pts[other]
        print(logn, kind, "xyzz", res[1], "mode %d" % other, res[other], flush=True)
        capi.check(L.sonic_dev_free(dsc))
