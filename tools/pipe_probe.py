"""Issue-rate probe of the multiplier pipes: IMAD.WIDE, carry-chain IMAD.WIDE, 32-bit IMAD, DFMA, and DFMA + IMAD.WIDE
issued together (do the FP64 and the integer multiplier pipes overlap?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import porla_b200 as pb

lib = pb.load()
lib.porla_device_init()
names = {0: "mad.wide.u32", 1: "mad.lo.cc/madc.hi.cc chain", 2: "mad.lo.u32", 3: "fma.rz.f64", 4: "fma.rz.f64 + mad.wide.u32 (each)"}
for v in (0, 1, 2, 3, 4):
    r = lib.porla_measure_pint(v, 0.5)
    print("variant %d %-36s %.3e /s  (%.1f lanes/clk/SM at 1.965 GHz, 148 SMs)" % (v, names[v], r, r / 148 / 1.965e9))
