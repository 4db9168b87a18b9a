"""Per-stage time of the FFT in the exponent with one thread per butterfly against four lanes per butterfly (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random
import porla_b200 as pb
from oracle import curves_py as O, loader
be = lambda v: v.to_bytes(32, "big")
rnd = random.Random(1)
G = O.bn254_marshal((1, 2)); step = O.bn254_marshal(O.mul(O.BN254, 0xABCDEF12345, (1, 2)))
for n in [int(x) for x in os.environ.get("NS", "128,1024,4096,16384,65536").split(",")]:
    macs = loader.bn254_point_chain(G, step, n)
    tw = b"".join(be(rnd.randrange(O.BN254.n)) for _ in range(256))
    res = {}
    for quad in ("0", "1"):
        os.environ["PORLA_BUTTERFLY_QUAD"] = quad
        t = pb.Table.from_host(pb.CURVE_BN254, macs)
        pb.load().porla_measure_pint(1, 0.2)
        ts = []
        for _ in range(15):
            t0 = time.perf_counter()
            t.butterfly_stage(min(512, n), tw)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        res[quad] = t.export()
        print(n, "quad", quad, "min %.3f median %.3f ms/stage" % (ts[0], ts[len(ts) // 2]), flush=True)
        t.destroy()
    print(n, "same bytes after 15 stages:", res["0"] == res["1"], flush=True)
