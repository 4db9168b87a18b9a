"""Latency of dependent field / point operations on one SM (porla_debug_latency; development aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import porla_b200 as pb

lib = pb.load(); lib.porla_device_init()
names = {0: "mul chain", 1: "2 mul chains", 2: "4 mul chains", 3: "sqr chain", 4: "XYZZ add (inlined)", 5: "XYZZ add (outlined mul)",
         6: "mixed add", 7: "doubling", 10: "quad add (4 lanes/point)", 11: "quad doubling"}
cyc, ns = C.c_double(), C.c_double()
for curve, cn in ((pb.CURVE_BN254, "bn254"), (pb.CURVE_SECP256K1, "secp256k1")):
    for mode in names:
        row = []
        for warps in (1, 4, 8):
            lib.porla_debug_latency(curve, mode, warps, 2000, C.byref(cyc), C.byref(ns))
            row.append("%dw %7.0f cyc %6.2f us" % (warps, cyc.value, ns.value / 1e3))
        print("%-9s %-24s %s" % (cn, names[mode], " | ".join(row)), flush=True)
