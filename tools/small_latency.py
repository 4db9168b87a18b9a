"""Where the time of a Porla-shaped call goes (development aid): per call wall clock through the C-ABI, the
kernel's share (CUDA events inside the library), for the audit MSMs (128 / 766 terms, 31-bit coefficients),
the 128-term commitment over the resident SRS and batches of commitments."""
import ctypes as C, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import porla_b200 as pb
from oracle import curves_py as O, loader
from porla_b200.lib import _slice

lib = pb.load(); lib.porla_device_init()
rnd = random.Random(1)
be = lambda v: v.to_bytes(32, "big")
k = pb.Kzg(bytes.fromhex("ffeeddccbbaa99887766554433221100"), bytes.fromhex("00112233445566778899aabbccddeeff"))
k.init_srs(128)
G = O.bn254_marshal((1, 2))
step = O.bn254_marshal(O.mul(O.BN254, 0xABCDEF12345, (1, 2)))
buf = (C.c_float * 8)()


def timeit(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    wall = (time.perf_counter() - t) / reps * 1e3
    lib.porla_stage_timing_enable(1)
    dev = 0.0
    for _ in range(20):
        fn()
        n = lib.porla_stage_timing_read(buf)
        dev += sum(buf[j] for j in range(n))
    lib.porla_stage_timing_enable(0)
    return wall, dev / 20


for npts in (128, 766):
    macs = bytearray(loader.bn254_point_chain(G, step, npts))
    coeff = bytearray(b"".join(pb.bn254_scalar_set_int(rnd.randrange(1 << 31)) for _ in range(npts)))
    out = bytearray(64)
    gs = [_slice(coeff), _slice(macs), _slice(out)]
    fn = lambda: lib.compute_multi_exp(C.byref(gs[0]), C.byref(gs[1]), npts, C.byref(gs[2]))
    w, d = timeit(fn)
    print("compute_multi_exp n=%d 31-bit: %.1f us per call, %.1f us on the device (launch sequence)" % (npts, w * 1e3, d * 1e3), flush=True)
    wide = bytearray(b"".join(be(rnd.randrange(O.BN254.n)) for _ in range(npts)))
    gs[0] = _slice(wide)
    w, d = timeit(fn)
    print("compute_multi_exp n=%d 254-bit: %.1f us per call, %.1f us on the device" % (npts, w * 1e3, d * 1e3), flush=True)
block = bytearray(b"".join(be(rnd.randrange(1 << 256)) for _ in range(128)))
out = bytearray(64)
g_in, g_out = _slice(block), _slice(out)
w, d = timeit(lambda: lib.compute_digest_from_srs(C.byref(g_in), C.byref(g_out)))
print("compute_digest_from_srs: %.1f us per call, %.1f us on the device" % (w * 1e3, d * 1e3), flush=True)
for batch in (8, 64, 1024):
    blocks = b"".join(be(rnd.randrange(1 << 256)) for _ in range(128 * batch))
    w, d = timeit(lambda: k.compute_digest_from_srs_batch(blocks, batch), reps=20, warm=3)
    print("compute_digest_from_srs_batch %d: %.3f ms per call, %.3f ms on the device" % (batch, w, d), flush=True)
# IPA mode: one inner-product proof over 128 resident generators + u (Server::inner_product_prove, Server.hpp:2279-2443)
c = O.SECP256K1
gens = [O.mul(c, rnd.randrange(1, c.n), (c.gx, c.gy)) for _ in range(129)]
tab = pb.SecpGenerators(gens)
a = [rnd.randrange(1 << 256) for _ in range(128)]
b = [rnd.randrange(c.n) for _ in range(128)]
lib.porla_measure_pint(1, 0.2)
for _ in range(3):
    tab.inner_product_prove(a, b)
t0 = time.perf_counter()
for _ in range(20):
    tab.inner_product_prove(a, b)
print("inner_product_prove (128 generators, 6 rounds, 12 multi-exponentiations): %.3f ms per proof" % ((time.perf_counter() - t0) / 20 * 1e3), flush=True)
tab.destroy()
