#!/usr/bin/env python3
"""Workload for ncu: standalone MSMs (BASELINE config 3) at the small sizes, scalars resident in HBM.
`ncu ... python tools/profile_msm_small.py [log_sizes...]` -- prints the library's own stage timings too."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sonic_b200 as sb  # noqa: E402
from sonic_b200 import capi, synth  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [16, 18]
sb.init(0)
L = capi.lib()
x, alpha = synth.trapdoor()
srs = sb.SRS.new(1 << 19, x, alpha)
for logn in sizes:
    N = 1 << logn
    sc = np.ascontiguousarray(synth.fr_bytes_fast(logn, N))
    dsc = ctypes.c_void_p()
    capi.check(L.sonic_dev_alloc(sc.nbytes, ctypes.byref(dsc)))
    capi.check(L.sonic_dev_upload(dsc, sc.ctypes.data, sc.nbytes))
    o48 = ctypes.create_string_buffer(48)
    for rep in range(4):
        capi.check(L.sonic_msm_g1_device(srs._h, 0, -(N // 2), N, dsc, o48))
        tm = {k: round(sb.last_timing_ms(k), 4) for k in ("total", "msm", "msm.sort", "msm.accumulate", "msm.accumulate_kernel", "msm.reduce",
                                                          "msm.window_bits", "msm.chunk", "msm.buckets", "msm.entries")}
        print("msm 2^%d rep %d" % (logn, rep), tm, "launches", sb.launch_count(), flush=True)
    capi.check(L.sonic_dev_free(dsc))
