"""Data-side FFT stage (porla_data_butterfly_stage_device) on resident blocks: time per stage and effective bandwidth."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb

lib = pb.load(); lib.porla_device_init()
LCM = int("2049369031155707573937272810025244064710333118140408897690954651424664974620215782673575413484558574566298823256897068805013612518402283464943595715297281")
st = torch.cuda.current_stream().cuda_stream
for lg in (10, 14, 16):
    n, chunks, m = 1 << lg, 128, 1 << lg
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    blocks = torch.randint(-2**31, 2**31 - 1, (n * chunks, 16), dtype=torch.int32, device="cuda", generator=g)
    blocks[:, 15] &= 0x0FFFFFFF                       # below LCM (~2^510.3)
    tw = torch.randint(-2**31, 2**31 - 1, (m // 2, 8), dtype=torch.int32, device="cuda", generator=g)
    run = lambda: lib.porla_data_butterfly_stage_device(C.c_void_p(blocks.data_ptr()), n, chunks, m, C.c_void_p(tw.data_ptr()), LCM.to_bytes(64, "little"), C.c_void_p(st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = n * chunks * 64 * 2
    print("2^%d blocks x 128 chunks: %.3f ms per stage, %.0f GB/s (read + write), %.2e chunk butterflies/s" % (lg, ms, byts / ms / 1e6, n * chunks / 2 / ms * 1e3), flush=True)
