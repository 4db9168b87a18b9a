"""secp256k1 MSM timing over window sizes (dev aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb
lib = pb.load(); lib.porla_device_init()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda"); g.manual_seed(5)
for lg in [int(x) for x in os.environ.get("SIZES", "18,20").split(",")]:
    n = 1 << lg
    ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    tab = pb.Table.multiples_of_generator(pb.CURVE_SECP256K1, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
    sc = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    for w in [int(x) for x in os.environ.get("WINDOWS", "0,13,15,16,17").split(",")]:
        for _ in range(2): tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32, window_bits=w, stream=st)
        t0 = time.perf_counter(); reps = 5
        for _ in range(reps): tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32, window_bits=w, stream=st)
        ms = (time.perf_counter() - t0) / reps * 1e3
        print("secp 2^%d c=%d: %.3f ms %.3e pts/s" % (lg, w or lib.porla_choose_window(1, n, 1), ms, n / ms * 1e3), flush=True)
    tab.destroy()
