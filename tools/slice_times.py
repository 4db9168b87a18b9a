"""Development aid: on ONE GPU, the time of bucket slice r of S of a 2^lg-term MSM (what each of S devices would run, without
the gather), whole and in two parts, against the whole MSM and the range shard of the same size."""
import ctypes as C, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb
import bench
lib = pb.load(); lib.porla_device_init()
dev = torch.device("cuda", 0)
names = ["count", "scan", "scatter", "accum", "reduce", "final"]
buf = (C.c_float * 8)()


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for lg in [int(x) for x in sys.argv[1:]] or [24]:
    n = 1 << lg
    S = int(os.environ.get("SLICES", "8"))
    rnd = random.Random(1000 + lg)
    a = rnd.getrandbits(228) | (1 << 227) | 1
    b = rnd.getrandbits(255) | (1 << 254)
    ks, ss = bench.closed_form_inputs(torch, 0, n, a, b, dev)
    g = torch.Generator(device=dev); g.manual_seed(lg)
    ss_r = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device=dev, generator=g)
    tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
    del ks
    c_, w_ = C.c_int(0), C.c_int(0)
    lib.porla_msm_plan(pb.CURVE_BN254, n, 1, 0, C.byref(c_), C.byref(w_))
    ws = torch.zeros(w_.value * 128, dtype=torch.uint8, device=dev)
    bk = torch.empty(int(lib.porla_msm_slice_bucket_bytes(pb.CURVE_BN254, c_.value, S)), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    print("2^%d: plan c=%d nwin=%d, max slices %d" % (lg, c_.value & 0xff, w_.value, lib.porla_msm_max_slices(pb.CURVE_BN254, c_.value, S)), flush=True)
    for sname, sc in (("a*i+b", ss), ("random", ss_r)):
        t_whole = timed(lambda: tab.msm_resident(sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32), 3)
        lib.porla_stage_timing_enable(1)
        for r in (0, S - 1):
            def one():
                lib.porla_msm_slice_window_sums_device(C.c_void_p(tab.handle), 0, C.c_void_p(sc.data_ptr()), n, pb.SCALAR_LE32, c_.value,
                                                       r, S, 0, None, C.c_void_p(ws.data_ptr()), C.c_void_p(stream))
            t1 = timed(one)
            lib.porla_stage_timing_read(buf)
            st = "  ".join("%s %.3f" % (nm, buf[j]) for j, nm in enumerate(names))
            own = n // S
            def two():
                lib.porla_msm_slice_window_sums_device(C.c_void_p(tab.handle), r * own, C.c_void_p(sc.data_ptr() + r * own * 32), own,
                                                       pb.SCALAR_LE32, c_.value, r, S, 1, C.c_void_p(bk.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(stream))
                lib.porla_msm_slice_window_sums_device(C.c_void_p(tab.handle), 0, C.c_void_p(sc.data_ptr()), n, pb.SCALAR_LE32, c_.value,
                                                       r, S, 3, C.c_void_p(bk.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(stream))
            t2 = timed(two)     # (the own range is counted twice here: an upper bound for the two-part form)
            print("2^%d %-7s slice %d/%d: one part %7.3f ms (%s) | two parts <= %7.3f ms | whole MSM %7.3f ms -> %.2fx / %.2fx" %
                  (lg, sname, r, S, t1, st, t2, t_whole, t_whole / t1, t_whole / t2), flush=True)
        lib.porla_stage_timing_enable(0)
    tab.destroy()
