"""Per-stage CUDA-event times of the device MSM for a list of sizes / window widths (dev aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb

lib = pb.load(); lib.porla_device_init()
sizes = [int(x) for x in os.environ.get("SIZES", "16,20,24").split(",")]
windows = [int(x) for x in os.environ.get("WINDOWS", "0").split(",")]
nmax = 1 << max(sizes)
g = torch.Generator(device="cuda"); g.manual_seed(1)
ks = torch.randint(-2**31, 2**31 - 1, (nmax, 8), dtype=torch.int32, device="cuda", generator=g)
tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), nmax, pb.SCALAR_LE32, on_device=True)
sc = torch.randint(-2**31, 2**31 - 1, (nmax, 8), dtype=torch.int32, device="cuda", generator=g)
st = torch.cuda.current_stream().cuda_stream
names = ["count", "scan", "scatter", "accum", "reduce", "final"]
lib.porla_stage_timing_enable(1)
buf = (C.c_float * 8)()
for lg in sizes:
    n = 1 << lg
    for w in windows:
        acc = [0.0] * 6
        reps = 3
        for r in range(reps + 1):
            out = tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32, window_bits=w, stream=st)
            torch.cuda.synchronize()
            k = lib.porla_stage_timing_read(buf)
            if r:
                acc = [a + buf[j] for j, a in enumerate(acc)]
        acc = [a / reps for a in acc]
        c = w or lib.porla_choose_window(0, n, 1)
        print("2^%d c=%2d total %8.3f ms | " % (lg, c, sum(acc)) + "  ".join("%s %.3f" % (nm, v) for nm, v in zip(names, acc)), flush=True)
