"""Stage times of one MSM with the closed-form inputs of the strong-scaling block against uniform random inputs."""
import ctypes as C, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb
import bench
lib = pb.load(); lib.porla_device_init()
names = ["count", "scan", "scatter", "accum", "reduce", "final"]
buf = (C.c_float * 8)()
lib.porla_stage_timing_enable(1)
dev = torch.device("cuda", 0)
for lg in [int(x) for x in sys.argv[1:]] or [21, 24]:
    n = 1 << lg
    rnd = random.Random(1000 + lg)
    a = rnd.getrandbits(228) | (1 << 227) | 1
    b = rnd.getrandbits(255) | (1 << 254)
    ks, ss = bench.closed_form_inputs(torch, 0, n, a, b, dev)
    g = torch.Generator(device=dev); g.manual_seed(lg)
    ks_r = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device=dev, generator=g)
    ss_r = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device=dev, generator=g)
    for pname, kk in (("points (i+1)G", ks), ("random points", ks_r)):
        tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, kk.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
        for sname, sc in (("scalars a*i+b", ss), ("random scalars", ss_r)):
            for r in range(3):
                tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32)
            torch.cuda.synchronize(); lib.porla_stage_timing_read(buf)
            print("2^%d %-14s %-15s total %7.3f | " % (lg, pname, sname, sum(buf[j] for j in range(6))) + "  ".join("%s %.3f" % (nm, buf[j]) for j, nm in enumerate(names)), flush=True)
        tab.destroy()
