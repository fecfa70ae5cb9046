#!/usr/bin/env python3
"""Ad-hoc GPU exploration: IMAD peak, SRS generation time, MSM timing breakdown by size/window."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sonic_b200 as sb
from sonic_b200 import capi, synth

def main():
    logd = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    sb.init(0)
    L = capi.lib()
    for variant in (0, 1):
        print("imad_peak variant", variant, "%.3e LMAC/s" % L.sonic_imad_peak_lmacs(variant, 4000), flush=True)
    x, alpha = synth.trapdoor()
    d = 1 << logd
    t = time.time(); srs = sb.SRS.new(d, x, alpha); dt = time.time() - t
    print(f"srs_new d=2^{logd}: wall {dt*1e3:.1f} ms, device {sb.last_timing_ms('total'):.1f} ms, {(4*d+1)/dt/1e6:.2f} Mpoints/s", flush=True)
    for logn in range(12, logd + 2, 2):
        N = 1 << logn
        sc = synth.fr_bytes_fast(logn, N)
        for wb in (0,):
            sb.set_option("window_bits", wb)
            best = None
            for rep in range(3):
                t = time.time(); out = sb.msm(srs, 0, -(N // 2), sc); dt = time.time() - t
                tm = {k: sb.last_timing_ms(k) for k in ("total", "msm", "msm.sort", "msm.accumulate", "msm.reduce")}
                if best is None or tm["msm"] < best[1]["msm"]:
                    best = (dt, tm)
            dt, tm = best
            print(f"msm N=2^{logn} wb={wb}: wall {dt*1e3:.2f} ms  dev {tm}  -> {N/tm['msm']/1e3:.2f} Mpoints/s (kernels)", flush=True)
    sk = synth.skewed_fr_bytes(5, 1 << (logd))
    t = time.time(); out = sb.msm(srs, 0, -(1 << (logd - 1)), sk); dt = time.time() - t
    print(f"msm skewed N=2^{logd}: wall {dt*1e3:.2f} ms dev msm {sb.last_timing_ms('msm'):.2f} ms", flush=True)

main()
