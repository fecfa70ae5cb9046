#!/usr/bin/env python3
"""Standalone MSM timing by size under option settings: python tools/tune_msm.py name=v1,v2 ..."""
import ctypes, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sonic_b200 as sb
from sonic_b200 import capi, synth
opts = {}
sizes = [16, 18, 20]
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "sizes":
        sizes = [int(x) for x in v.split(",")]
    else:
        opts[k] = [int(x) for x in v.split(",")]
sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
srs = sb.SRS.new(1 << (max(sizes) - 1), x, alpha)
for rv in itertools.product(*opts.values()) if opts else [()]:
    for k, v in zip(opts, rv):
        sb.set_option(k, v)
    for logn in sizes:
        N = 1 << logn
        sc = synth.fr_bytes_fast(logn, N)
        dsc = ctypes.c_void_p()
        capi.check(L.sonic_dev_alloc(sc.nbytes, ctypes.byref(dsc)))
        capi.check(L.sonic_dev_upload(dsc, sc.ctypes.data, sc.nbytes))
        o = ctypes.create_string_buffer(48)
        best = None
        for _ in range(5):
            capi.check(L.sonic_msm_g1_device(srs._h, 0, -(N // 2), N, dsc, o))
            t = {k: sb.last_timing_ms(k) for k in ("total", "msm.sort", "msm.accumulate_kernel", "msm.accumulate", "msm.reduce", "msm.chunk", "msm.window_bits")}
            if best is None or t["total"] < best["total"]:
                best = t
        print(dict(zip(opts, rv)), "2^%d" % logn, {k: round(v, 3) for k, v in best.items()}, flush=True)
        capi.check(L.sonic_dev_free(dsc))
