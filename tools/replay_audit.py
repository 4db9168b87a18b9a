"""Replay of the C-ABI call sequence of one Porla KZG-mode audit and one update (SURVEY.md
Appendix C; Server.hpp:564-931, Client.hpp:633-892, Server.hpp:401-476) against libmultiexp.so, with
the MSM calls also timed on the CPU port.  NTL / ZeroMQ / file work of the real Client/Server is
excluded (they cannot be built here); data are synthetic with the reference's shapes:
NUM_CHUNKS = 128 coefficients per block, 31-bit audit coefficients, n_points in {128, 766}."""
import json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import porla_b200 as pb
from oracle import curves_py as O, loader

rnd = random.Random(1)
be = lambda v: v.to_bytes(32, "big")
lib = pb.load(); lib.porla_device_init()
k = pb.Kzg(bytes.fromhex("ffeeddccbbaa99887766554433221100"), bytes.fromhex("00112233445566778899aabbccddeeff"))
t0 = time.perf_counter(); blob = k.init_srs(128); t_init = time.perf_counter() - t0
srs = b"".join(O.bn254_marshal(O.bn254_unmarshal(blob[132 + 32 * i:164 + 32 * i])) for i in range(128))
G = O.bn254_marshal((1, 2))
step = O.bn254_marshal(O.mul(O.BN254, 0xABCDEF12345, (1, 2)))


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e3


out = {"init_SRS_ms": t_init * 1e3, "cpu_threads": os.cpu_count()}
block = b"".join(be(rnd.randrange(1 << 256)) for _ in range(128))
for npts in (128, 766):
    macs = loader.bn254_point_chain(G, step, npts)
    macs = bytearray(macs)
    for i in range(0, npts, 7):
        macs[64 * i:64 * i + 64] = bytes(64)          # alignment MACs that are still infinity
    coeff = b"".join(pb.bn254_scalar_set_int(rnd.randrange(1 << 31)) for _ in range(npts))
    r = {}
    r["compute_multi_exp_ms"] = timeit(lambda: pb.bn254_multi_exp(bytes(macs), coeff, npts))
    r["compute_multi_exp_cpu_port_1thread_ms"] = timeit(lambda: loader.bn254_msm(coeff, bytes(macs), npts, 1), reps=5, warm=1)
    r["compute_multi_exp_cpu_port_allthreads_ms"] = timeit(lambda: loader.bn254_msm(coeff, bytes(macs), npts, os.cpu_count()), reps=5, warm=1)
    assert pb.bn254_multi_exp(bytes(macs), coeff, npts) == loader.bn254_msm(coeff, bytes(macs), npts, 1)
    out["n_points_%d" % npts] = r
out["compute_digest_from_srs_ms"] = timeit(lambda: k.compute_digest_from_srs(block))
out["compute_digest_from_srs_cpu_port_ms"] = timeit(lambda: loader.bn254_msm(block, srs, 128, 1), reps=5, warm=1)
out["create_proof_ms"] = timeit(lambda: k.create_proof(123456789, block))
c_, h_, z_, y_ = k.create_proof(123456789, block)
out["verify_proof_ms"] = timeit(lambda: k.verify_proof(c_, h_, z_, y_), reps=5, warm=1)
assert k.verify_proof(c_, h_, z_, y_)
buf = bytearray(c_)
out["mult_point_ms"] = timeit(lambda: pb.bn254_mult(buf, be(rnd.randrange(1 << 254))), reps=50)
out["add_point_ms"] = timeit(lambda: pb.bn254_add(buf, h_), reps=50)
out["compute_digest_ms"] = timeit(lambda: k.compute_digest(block))
# batched align_MAC: 1024 block commitments in one launch sequence
blocks = b"".join(be(rnd.randrange(1 << 256)) for _ in range(128 * 1024))
out["compute_digest_from_srs_batch_1024_ms"] = timeit(lambda: k.compute_digest_from_srs_batch(blocks, 1024), reps=3, warm=1)
# CRebuild_Cached on the MACs (Server.hpp:1548-1687): log2(n) butterfly stages over n = 1024 points.
# Reference call pattern: per butterfly mult_point + add_point + neg_point + add_point through the C-ABI
# (timed here on 64 butterflies and scaled); batched: one bn254 butterfly launch per stage on a resident table.
nblk = 1024
macs = loader.bn254_point_chain(G, step, nblk)
stages = []
m = 2
while m <= nblk:
    stages.append((m, b"".join(be(rnd.randrange(O.BN254.n)) for _ in range(m // 2))))
    m *= 2


def rebuild_batched():
    t = pb.Table.from_host(pb.CURVE_BN254, macs)
    for m_, tw in stages:
        t.butterfly_stage(m_, tw)
    res = t.export()
    t.destroy()
    return res


def butterflies_per_call(count):
    a0, a1 = bytearray(macs[:64]), bytearray(macs[64:128])
    w = stages[3][1][:32]
    for _ in range(count):
        tm = bytearray(a1); pb.bn254_mult(tm, w)
        um = bytearray(a0); pb.bn254_add(a0, bytes(tm))
        pb.bn254_neg(tm); a1[:] = um; pb.bn254_add(a1, bytes(tm))


out["crebuild_1024_batched_gpu_ms"] = timeit(rebuild_batched, reps=5, warm=2)
per = timeit(lambda: butterflies_per_call(64), reps=3, warm=1) / 64
out["butterfly_per_call_host_ms"] = per
out["crebuild_1024_per_call_host_1thread_ms"] = per * (nblk // 2) * len(stages)
# spot check: first stage against the per-call path
chk = bytearray(macs)
pb.bn254_butterfly_stage(chk, nblk, 2, stages[0][1])
a0, a1 = bytearray(macs[:64]), bytearray(macs[64:128])
tm = bytearray(a1); pb.bn254_mult(tm, stages[0][1][:32]); um = bytearray(a0); pb.bn254_add(a0, bytes(tm)); pb.bn254_neg(tm); a1[:] = um; pb.bn254_add(a1, bytes(tm))
assert bytes(chk[:128]) == bytes(a0) + bytes(a1)
for npts in (128, 766):
    r = out["n_points_%d" % npts]
    out["server_audit_msm_total_ms_n%d" % npts] = 2 * r["compute_multi_exp_ms"] + out["compute_digest_from_srs_ms"] + out["create_proof_ms"]
    out["client_audit_total_ms_n%d" % npts] = r["compute_multi_exp_ms"] + 2 * out["mult_point_ms"] + 2 * out["add_point_ms"] + out["verify_proof_ms"]
print(json.dumps(out, indent=1))
