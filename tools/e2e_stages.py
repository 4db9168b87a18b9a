"""Stage times of the LAST part of a streamed host-buffer MSM (and the whole-call time), for PORLA_STREAM_PARTS settings."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import porla_b200 as pb
lib = pb.load(); lib.porla_device_init()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
g = torch.Generator(device="cuda"); g.manual_seed(lg)
ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
pts = torch.from_numpy(np.frombuffer(tab.export(), dtype=np.uint8).copy()).pin_memory()
sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8).pin_memory()
d_sc = sc.cuda()
out = (C.c_ubyte * 64)()
names = ["count", "scan", "scatter", "accum", "reduce", "final"]
buf = (C.c_float * 8)()
lib.porla_stage_timing_enable(1)
ref = tab.msm_resident(d_sc.data_ptr(), n, scalar_fmt=pb.SCALAR_BE32)
ref = tab.msm_resident(d_sc.data_ptr(), n, scalar_fmt=pb.SCALAR_BE32)
torch.cuda.synchronize(); lib.porla_stage_timing_read(buf)
print("resident            : " + "  ".join("%s %.3f" % (nm, buf[j]) for j, nm in enumerate(names)), flush=True)
for parts in [x for x in os.environ.get("PARTS", "1,2,4,8").split(",")]:
    os.environ["PORLA_STREAM_PARTS"] = parts
    lib.porla_measure_pint(1, 0.1)
    ts = []
    for r in range(6):
        t0 = time.perf_counter()
        lib.porla_msm_host_devices(0, C.c_void_p(sc.data_ptr()), C.c_void_p(pts.data_ptr()), n, pb.SCALAR_BE32, pb.POINT_BE64, 1, C.cast(out, C.c_void_p))
        ts.append((time.perf_counter() - t0) * 1e3)
    assert bytes(out) == ref
    lib.porla_stage_timing_read(buf)
    print("parts %s: call %.3f ms (min of 6) | last part: " % (parts, min(ts)) + "  ".join("%s %.3f" % (nm, buf[j]) for j, nm in enumerate(names)), flush=True)
