import os, sys, time
sys.path.insert(0, "/root/repo")
import random
import porla_b200 as pb
from oracle import curves_py as O, loader
be = lambda v: v.to_bytes(32, "big")
rnd = random.Random(1)
G = O.bn254_marshal((1, 2)); step = O.bn254_marshal(O.mul(O.BN254, 0xABCDEF12345, (1, 2)))
for n in (1024, 1 << 16, 1 << 20):
    macs = loader.bn254_point_chain(G, step, n)
    t = pb.Table.from_host(pb.CURVE_BN254, macs)
    tw = b"".join(be(rnd.randrange(O.BN254.n)) for _ in range(4))
    for mode in ("i", "c"):
        os.environ["PORLA_BUTTERFLY_FIELD"] = mode
        t.butterfly_stage(8, tw)
        t0 = time.perf_counter()
        for _ in range(3): t.butterfly_stage(8, tw)
        print(n, mode, "%.3f ms/stage" % ((time.perf_counter() - t0) / 3 * 1e3), flush=True)
    t.destroy()
