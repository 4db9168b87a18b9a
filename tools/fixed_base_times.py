"""Single MSM over a resident table with and without its fixed-base window expansion (development aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb

lib = pb.load(); lib.porla_device_init()
lg = int(os.environ.get("LG", "20")); n = 1 << lg
g = torch.Generator(device="cuda"); g.manual_seed(1)
ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
sc = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
st = torch.cuda.current_stream().cuda_stream
names = ["count", "scan", "scatter", "accum", "reduce", "final"]
buf = (C.c_float * 8)()


def run(label):
    lib.porla_stage_timing_enable(1)
    acc = [0.0] * 6
    for r in range(4):
        out = tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32, stream=st)
        torch.cuda.synchronize()
        lib.porla_stage_timing_read(buf)
        if r:
            acc = [a + buf[j] for j, a in enumerate(acc)]
    lib.porla_stage_timing_enable(0)
    acc = [a / 3 for a in acc]
    print("%-18s total %7.3f ms | " % (label, sum(acc)) + "  ".join("%s %.3f" % (nm, v) for nm, v in zip(names, acc)), flush=True)
    return out


ref = run("general")
for c in [int(x) for x in os.environ.get("CS", "16,18,19,20,21").split(",")]:
    tab.precompute(c, n, 1)
    got = run("fixed-base c=%d" % c)
    assert got == ref, c
