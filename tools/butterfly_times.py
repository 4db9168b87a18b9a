"""Per-stage time of the FFT in the exponent on a resident table (development aid): field flavour (inlined /
compact) x scalar form (plain double-and-add / GLV joint double-and-add), CUDA-synchronised wall clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random
import porla_b200 as pb
from oracle import curves_py as O, loader
be = lambda v: v.to_bytes(32, "big")
try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    clock = lambda: pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:
    clock = lambda: "?"
rnd = random.Random(1)
G = O.bn254_marshal((1, 2)); step = O.bn254_marshal(O.mul(O.BN254, 0xABCDEF12345, (1, 2)))
for n in [int(x) for x in os.environ.get("NS", "1024,65536,1048576").split(",")]:
    macs = loader.bn254_point_chain(G, step, n)
    t = pb.Table.from_host(pb.CURVE_BN254, macs)
    tw = b"".join(be(rnd.randrange(O.BN254.n)) for _ in range(256))
    for mode in ("i", "c"):
        for glv in ("0", "1"):
            os.environ["PORLA_BUTTERFLY_FIELD"] = mode
            os.environ["PORLA_BUTTERFLY_GLV"] = glv
            pb.load().porla_measure_pint(1, 0.2)      # a cold GPU runs these launches several times slower
            ts = []
            for _ in range(25):
                t0 = time.perf_counter()
                t.butterfly_stage(512, tw)
                ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            print(n, "field", mode, "glv", glv, "min %.3f median %.3f max %.3f ms/stage, SM clock %s MHz" % (ts[0], ts[len(ts) // 2], ts[-1], clock()), flush=True)
    t.destroy()
