"""Timing of BASELINE configs 3 (4096 MSMs x 2^12 over one shared SRS, one launch sequence) and
4 (secp256k1 2^18-point multi-exponentiation), device-resident inputs."""
import ctypes as C, os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb
lib = pb.load(); lib.porla_device_init()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda"); g.manual_seed(5)

def ev_time(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# config 3
n, nb = 1 << 12, int(os.environ.get("NB", "4096"))
ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
srs = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
sc = torch.randint(-2**31, 2**31 - 1, (nb * n, 8), dtype=torch.int32, device="cuda", generator=g)
out = torch.zeros(64 * nb, dtype=torch.uint8, device="cuda")
for w in [int(x) for x in os.environ.get("WINDOWS", "0").split(",")]:
    ms = ev_time(lambda: srs.msm_device(sc.data_ptr(), n, out.data_ptr(), nbatch=nb, scalar_fmt=pb.SCALAR_LE32, shared_points=True, window_bits=w, stream=st))
    c = w or lib.porla_choose_window(0, n, nb)
    lib.porla_stage_timing_enable(1)
    srs.msm_device(sc.data_ptr(), n, out.data_ptr(), nbatch=nb, scalar_fmt=pb.SCALAR_LE32, shared_points=True, window_bits=w, stream=st)
    torch.cuda.synchronize()
    buf = (C.c_float * 8)(); lib.porla_stage_timing_read(buf); lib.porla_stage_timing_enable(0)
    print("   stages ms: count %.2f scan %.2f scatter %.2f accum %.2f reduce %.2f final %.2f" % tuple(buf[i] for i in range(6)))
    print("config3: %d MSMs x 2^12, c=%d: %.2f ms  %.3e points/s  (%.3e MAC32/s algorithmic)" % (nb, c, ms, nb * n / ms * 1e3, nb * n * 21760 / ms * 1e3), flush=True)
# fixed-base expansion of the shared SRS
cfb = srs.precompute(0, n, nb)
out2 = torch.zeros(64 * nb, dtype=torch.uint8, device="cuda")
ms = ev_time(lambda: srs.msm_device(sc.data_ptr(), n, out2.data_ptr(), nbatch=nb, scalar_fmt=pb.SCALAR_LE32, shared_points=True, stream=st))
lib.porla_stage_timing_enable(1)
srs.msm_device(sc.data_ptr(), n, out2.data_ptr(), nbatch=nb, scalar_fmt=pb.SCALAR_LE32, shared_points=True, stream=st)
torch.cuda.synchronize()
buf = (C.c_float * 8)(); lib.porla_stage_timing_read(buf); lib.porla_stage_timing_enable(0)
print("   stages ms: count %.2f scan %.2f scatter %.2f accum %.2f reduce %.2f final %.2f" % tuple(buf[i] for i in range(6)))
print("config3 fixed-base tables c=%d: %.2f ms  %.3e points/s  (%.3e MAC32/s algorithmic)" % (cfb, ms, nb * n / ms * 1e3, nb * n * 21760 / ms * 1e3), flush=True)
assert torch.equal(out, out2), "fixed-base batch differs from the general path"
# spot-check one MSM of the batch against the single-MSM path
m = nb // 3
single = srs.msm_resident(sc[m * n:(m + 1) * n].data_ptr(), n, scalar_fmt=pb.SCALAR_LE32)
assert bytes(out[64 * m:64 * m + 64].cpu().numpy().tobytes()) == single
del sc

# config 4
n = 1 << 18
ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
tab = pb.Table.multiples_of_generator(pb.CURVE_SECP256K1, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
sc = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
t0 = time.perf_counter(); reps = 5
for _ in range(2): tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32, stream=st)
t0 = time.perf_counter()
for _ in range(reps): tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32, stream=st)
ms = (time.perf_counter() - t0) / reps * 1e3
print("config4: secp256k1 2^18: %.3f ms  %.3e points/s (%.3e MAC32/s algorithmic @12800/pt)" % (ms, n / ms * 1e3, n * 12800 / ms * 1e3), flush=True)
