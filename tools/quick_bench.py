"""Ad-hoc timing of the device MSM at several sizes (development aid, not bench.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb

lib = pb.load()
lib.porla_device_init()
for v in ((0, 1, 2) if not os.environ.get("NOPINT") else ()):
    print("P_int variant", v, "%.4e MAC32/s" % lib.porla_measure_pint(v, 0.3), flush=True)

sizes = [int(x) for x in os.environ.get("SIZES", "16,20,22,24").split(",")]
nmax = 1 << max(sizes)
g = torch.Generator(device="cuda"); g.manual_seed(1)
ks = torch.randint(0, 2**31 - 1, (nmax, 8), dtype=torch.int32, device="cuda", generator=g)
ks[:, 7] &= 0x0FFFFFFF
t0 = time.time()
tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), nmax, pb.SCALAR_LE32, on_device=True)
torch.cuda.synchronize()
print("table of %d multiples: %.2fs" % (nmax, time.time() - t0), flush=True)
sc = torch.randint(0, 2**31 - 1, (nmax, 8), dtype=torch.int32, device="cuda", generator=g)
sc[:, 7] &= 0x0FFFFFFF
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
windows = [int(x) for x in os.environ.get("WINDOWS", "0").split(",")]
for lg in sizes:
    n = 1 << lg
    for w in windows:
        for _ in range(2):
            tab.msm_device(sc.data_ptr(), n, out.data_ptr(), scalar_fmt=pb.SCALAR_LE32, window_bits=w, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            tab.msm_device(sc.data_ptr(), n, out.data_ptr(), scalar_fmt=pb.SCALAR_LE32, window_bits=w, stream=st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        c = w or lib.porla_choose_window(0, n, 1)
        print("2^%d c=%d: %.3f ms  %.3e pts/s  (%.3e MAC32/s algorithmic)" % (lg, c, ms, n / ms * 1e3, n * 21760 / ms * 1e3), flush=True)
