#!/usr/bin/env python3
"""Integer-multiply roofline denominators measured on the device (register-only microbenchmarks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sonic_b200 as sb
from sonic_b200 import capi
sb.init(0)
names = ["mad.lo.cc+madc.hi (IMAD.WIDE.U32, carry out)", "mad.wide.u32 (IMAD.WIDE.U32)", "mad.lo + mad.hi (2 IMAD)",
         "madc.lo.cc/madc.hi.cc rows (IMAD.WIDE.U32.X, carry in+out)"]
for v, n in enumerate(names):
    print("variant %d  %-62s %.3f TLMAC/s" % (v, n, capi.lib().sonic_imad_peak_lmacs(v, 4000) / 1e12), flush=True)
