"""Workload for compute-sanitizer (memcheck / racecheck): touches every kernel of the pipeline once,
including the shared-memory radix partition (n >= 2^19) and the batched / fixed-base / butterfly paths."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import porla_b200 as pb
from oracle import curves_py as O
be = lambda v: v.to_bytes(32, "big")
lib = pb.load(); lib.porla_device_init()
g = torch.Generator(device="cuda"); g.manual_seed(3)
big = os.environ.get("SAN_BIG", "1") == "1"
for curve in (pb.CURVE_BN254, pb.CURVE_SECP256K1):
    n = (1 << 19) + 37 if big else 5000
    ks = torch.randint(0, 2**15, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    ks[:, 1:] = 0
    ks[5::7] = 0
    tab = pb.Table.multiples_of_generator(curve, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
    ss = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    a = tab.msm_resident(ss.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32)
    ss[:] = ss[0]
    b = tab.msm_resident(ss.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32)      # constant scalar: long stitch
    out = torch.zeros(64 * 8, dtype=torch.uint8, device="cuda")
    tab.msm_device(ss.data_ptr(), 600, out.data_ptr(), nbatch=8, scalar_fmt=pb.SCALAR_LE32, shared_points=True)
    tab2 = pb.Table.multiples_of_generator(curve, ks.data_ptr(), 256, pb.SCALAR_LE32, on_device=True)
    tab2.precompute(0, 256, 8)
    tab2.msm_device(ss.data_ptr(), 256, out.data_ptr(), nbatch=8, scalar_fmt=pb.SCALAR_LE32, shared_points=True)
    tab2.butterfly_stage(4, be(3) + be(5))
    torch.cuda.synchronize()
    print("curve", curve, a.hex()[:16], b.hex()[:16], flush=True)
    tab.destroy(); tab2.destroy()
k = pb.Kzg(bytes.fromhex("ffeeddccbbaa99887766554433221100"), bytes.fromhex("00112233445566778899aabbccddeeff"))
k.init_srs(32)
rnd = random.Random(1)
data = bytearray(b"".join(rnd.randrange(1 << 500).to_bytes(64, "little") for _ in range(32 * 3)))
print("align", k.align_mac_batch(data, 3).hex()[:16])
print("proof", k.create_proof(77, b"".join(be(rnd.randrange(1 << 256)) for _ in range(32)))[0].hex()[:16])
print("done")
