"""Porla's thread-pool call pattern (Server.hpp:1077-1078, Client.hpp:377-406): T host threads issue small calls at the
same time.  Wall clock per round of T calls, with the staging slots (default) and serialised (PORLA_SERIAL_CALLS=1)."""
import ctypes as C, os, random, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import porla_b200 as pb
from oracle import curves_py as O, loader
from porla_b200.lib import _slice

lib = pb.load(); lib.porla_device_init()
rnd = random.Random(1)
be = lambda v: v.to_bytes(32, "big")
k = pb.Kzg(bytes.fromhex("ffeeddccbbaa99887766554433221100"), bytes.fromhex("00112233445566778899aabbccddeeff"))
blob = k.init_srs(128)
srs = b"".join(O.bn254_marshal(O.bn254_unmarshal(blob[132 + 32 * i:164 + 32 * i])) for i in range(128))
T, ROUNDS = int(os.environ.get("T", "8")), 40
blocks = [bytearray(b"".join(be(rnd.randrange(1 << 256)) for _ in range(128))) for _ in range(T)]
outs = [bytearray(64) for _ in range(T)]
want = [loader.bn254_msm(bytes(b), srs, 128, 1) for b in blocks]
lib.porla_measure_pint(1, 0.3)


def worker(t, barrier):
    gi, go = _slice(blocks[t]), _slice(outs[t])
    for _ in range(ROUNDS):
        barrier.wait()
        lib.compute_digest_from_srs(C.byref(gi), C.byref(go))
        barrier.wait()


for label in ("warm-up", "timed"):
    barrier = threading.Barrier(T + 1)
    th = [threading.Thread(target=worker, args=(t, barrier)) for t in range(T)]
    for x in th: x.start()
    ts = []
    for _ in range(ROUNDS):
        barrier.wait(); t0 = time.perf_counter()
        barrier.wait(); ts.append((time.perf_counter() - t0) * 1e3)
    for x in th: x.join()
    ts.sort()
    if label == "timed":
        print("%d threads x compute_digest_from_srs: round of %d calls min %.3f median %.3f ms (serial=%s)" % (T, T, ts[0], ts[len(ts) // 2], os.environ.get("PORLA_SERIAL_CALLS", "0")), flush=True)
assert all(bytes(o) == w for o, w in zip(outs, want)), "concurrent results differ from the oracle"
