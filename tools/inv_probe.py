"""Cost of one modular inversion on the device: binary GCD (fp_inv.cuh) against Fermat, a lone block and the whole device."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import porla_b200 as pb
lib = pb.load(); lib.porla_device_init()
cy, ns = C.c_double(0), C.c_double(0)
lib.porla_measure_pint(1, 0.2)
for curve in (0, 1):
    for mode, name, iters in ((0, "field product", 2000), (8, "GCD inversion", 40), (9, "Fermat inversion", 20)):
        for warps in (1, 4):
            lib.porla_debug_latency(curve, mode, warps, iters, C.byref(cy), C.byref(ns))
            lone = cy.value
            lib.porla_debug_latency(curve, mode + 100, warps, iters, C.byref(cy), C.byref(ns))
            print("curve %d %-17s %d warps/block: lone block %9.0f cycles/op; whole device (4 blocks/SM) %9.0f cycles/op, %8.2f us/op"
                  % (curve, name, warps, lone, cy.value, ns.value / 1e3), flush=True)
