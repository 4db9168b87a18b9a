"""End-to-end time of compute_multi_exp with pinned host buffers at 2^LG terms (development aid; PORLA_SPLIT_PERCENT)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import porla_b200 as pb

lib = pb.load(); lib.porla_device_init()
lg = int(os.environ.get("LG", "20")); n = 1 << lg
g = torch.Generator(device="cuda"); g.manual_seed(7)
ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
table = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
pts_host = torch.empty(n * 64, dtype=torch.uint8).pin_memory()
lib.porla_table_export(C.c_void_p(table.handle), pb.POINT_BE64, C.c_void_p(pts_host.data_ptr()), 0, None)
sc = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, generator=torch.Generator().manual_seed(3))
sc_host = torch.from_numpy(np.ascontiguousarray(sc.numpy().view(np.uint8).reshape(n, 32)[:, ::-1])).pin_memory()
res = (C.c_ubyte * 64)()
a, b, c = pb.GoSlice(sc_host.data_ptr(), n * 32, n * 32), pb.GoSlice(pts_host.data_ptr(), n * 64, n * 64), pb.GoSlice(C.cast(res, C.c_void_p), 64, 64)
lib.porla_measure_pint(1, 0.3)
for _ in range(3):
    lib.compute_multi_exp(C.byref(a), C.byref(b), n, C.byref(c))
ts = []
for _ in range(15):
    t0 = time.perf_counter()
    lib.compute_multi_exp(C.byref(a), C.byref(b), n, C.byref(c))
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print("split %s%%  2^%d: min %.3f median %.3f ms  result %s" % (os.environ.get("PORLA_SPLIT_PERCENT", "default"), lg, ts[0], ts[len(ts) // 2], bytes(res).hex()[:16]), flush=True)
