"""Public EIP-196 (alt_bn128 ecAdd / ecMul) known-answer vectors: an independent pin for the BN254 group law and the
64-byte X||Y big-endian layout that crosses Porla's C-ABI (gnark-crypto's Marshal).  CPU tests run the
oracles and the library's host entry points (add_point / mult_point, main.go:196-212); the GPU tests run the same vectors
through compute_multi_exp (main.go:119-138) and the batched device kernels."""
import pytest

import porla_b200 as pb
from oracle import curves_py as O
from oracle import loader
from tests.common import golden

BN = O.BN254
VEC = golden("eip196_bn254.json")


def _pt(h):
    x, y = int(h[:64], 16), int(h[64:128], 16)
    return None if x == 0 and y == 0 else (x, y)


def test_vectors_are_on_the_curve():
    for v in VEC["add"] + VEC["mul"]:
        for off in (0, 128) if v in VEC["add"] else (0,):
            P = _pt(v["input"][off:off + 128])
            assert P is None or (P[1] * P[1] - P[0] ** 3 - 3) % BN.p == 0, v["name"]
        E = _pt(v["expected"])
        assert E is None or (E[1] * E[1] - E[0] ** 3 - 3) % BN.p == 0, v["name"]


def test_python_oracle_reproduces_eip196():
    for v in VEC["add"]:
        assert O.bn254_marshal(O.add(BN, _pt(v["input"][:128]), _pt(v["input"][128:]))).hex() == v["expected"], v["name"]
    for v in VEC["mul"]:
        s = int(v["input"][128:], 16)
        assert O.bn254_marshal(O.mul(BN, s % BN.n, _pt(v["input"][:128]))).hex() == v["expected"], v["name"]


def test_c_oracle_reproduces_eip196():
    import ctypes as C
    lib = loader.bn254()
    for v in VEC["add"]:
        out = C.create_string_buffer(64)
        lib.oracle_bn254_add(bytes.fromhex(v["input"][:128]), bytes.fromhex(v["input"][128:]), out)
        assert out.raw.hex() == v["expected"], v["name"]
    for v in VEC["mul"]:
        out = C.create_string_buffer(64)
        lib.oracle_bn254_mul(bytes.fromhex(v["input"][:128]), bytes.fromhex(v["input"][128:]), out)
        assert out.raw.hex() == v["expected"], v["name"]
        one = loader.bn254_msm(bytes.fromhex(v["input"][128:]), bytes.fromhex(v["input"][:128]), 1, 1)
        assert one.hex() == v["expected"], v["name"]


def test_library_host_point_ops_reproduce_eip196():
    """add_point / mult_point of the legacy ABI (host code inside libmultiexp.so; no GPU needed)."""
    for v in VEC["add"]:
        a = bytearray.fromhex(v["input"][:128])
        pb.bn254_add(a, bytes.fromhex(v["input"][128:]))
        assert a.hex() == v["expected"], v["name"]
    for v in VEC["mul"]:
        a = bytearray.fromhex(v["input"][:128])
        pb.bn254_mult(a, bytes.fromhex(v["input"][128:]))
        assert a.hex() == v["expected"], v["name"]


@pytest.mark.gpu
def test_compute_multi_exp_reproduces_eip196():
    """ecMul as a 1-term MSM, ecAdd as a 2-term MSM with scalars (1, 1), and all ecMul vectors as ONE MSM whose expected
    value is the sum of the published results; through the legacy symbol and through the forced bucket pipeline."""
    one = (1).to_bytes(32, "big")
    for v in VEC["mul"]:
        got = pb.bn254_multi_exp(bytes.fromhex(v["input"][:128]), bytes.fromhex(v["input"][128:]), 1)
        assert got.hex() == v["expected"], v["name"]
    for v in VEC["add"]:
        got = pb.bn254_multi_exp(bytes.fromhex(v["input"]), one + one, 2)
        assert got.hex() == v["expected"], v["name"]
    pts = b"".join(bytes.fromhex(v["input"][:128]) for v in VEC["mul"])
    scs = b"".join(bytes.fromhex(v["input"][128:]) for v in VEC["mul"])
    want = None
    for v in VEC["mul"]:
        want = O.add(BN, want, _pt(v["expected"]))
    assert pb.bn254_multi_exp(pts, scs, len(VEC["mul"])) == O.bn254_marshal(want)
    import os
    os.environ["PORLA_NO_SMALL"] = "1"
    try:
        assert pb.bn254_multi_exp(pts, scs, len(VEC["mul"])) == O.bn254_marshal(want)
    finally:
        del os.environ["PORLA_NO_SMALL"]


@pytest.mark.gpu
def test_scalar_mul_batch_device_reproduces_eip196_and_oracle():
    """porla_scalar_mul_batch_device (the batched form of mult_point / secp256k1_ecmult_const): out[i] =
    k_i * table[i], and the fixed-base form over a 1-entry table; against the published ecMul answers and the oracle."""
    import ctypes as C
    import random
    import torch
    lib = pb.load()
    pts = b"".join(bytes.fromhex(v["input"][:128]) for v in VEC["mul"])
    scs = b"".join(bytes.fromhex(v["input"][128:]) for v in VEC["mul"])
    n = len(VEC["mul"])
    tab = pb.Table.from_host(pb.CURVE_BN254, pts)
    d_sc = torch.frombuffer(bytearray(scs), dtype=torch.uint8).cuda()
    d_out = torch.zeros(64 * n, dtype=torch.uint8, device="cuda")
    lib.porla_scalar_mul_batch_device(C.c_void_p(tab.handle), C.c_void_p(d_sc.data_ptr()), n, pb.SCALAR_BE32, pb.POINT_BE64,
                                      C.c_void_p(d_out.data_ptr()), None)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().tobytes()
    for i, v in enumerate(VEC["mul"]):
        assert got[64 * i:64 * i + 64].hex() == v["expected"], v["name"]
    tab.destroy()
    # random scalars (zero, one, r - 1, above r) over random points with an infinity, both curves
    rnd = random.Random(11)
    for curve, c, fmt in ((pb.CURVE_BN254, BN, pb.SCALAR_BE32), (pb.CURVE_SECP256K1, O.SECP256K1, pb.SCALAR_LE32)):
        base = (c.gx, c.gy)
        P = [O.mul(c, rnd.randrange(1, c.n), base) for _ in range(20)] + [None]
        ks = [0, 1, c.n - 1, c.n + 5, (1 << 256) - 1] + [rnd.randrange(1 << 256) for _ in range(len(P) - 5)]
        enc = b"".join(bytes(64) if Q is None else Q[0].to_bytes(32, "big") + Q[1].to_bytes(32, "big") for Q in P)
        sc = b"".join(k.to_bytes(32, "big" if fmt == pb.SCALAR_BE32 else "little") for k in ks)
        tab = pb.Table.from_host(curve, enc)
        d_sc = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
        d_out = torch.zeros(64 * len(P), dtype=torch.uint8, device="cuda")
        lib.porla_scalar_mul_batch_device(C.c_void_p(tab.handle), C.c_void_p(d_sc.data_ptr()), len(P), fmt, pb.POINT_BE64,
                                          C.c_void_p(d_out.data_ptr()), None)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().tobytes()
        for i, (Q, k) in enumerate(zip(P, ks)):
            W = None if Q is None else O.mul(c, k % c.n, Q)
            want = bytes(64) if W is None else W[0].to_bytes(32, "big") + W[1].to_bytes(32, "big")
            assert got[64 * i:64 * i + 64] == want, (curve, i)
        tab.destroy()
        # fixed base: a 1-entry table, every scalar multiplies the same point
        tab1 = pb.Table.from_host(curve, enc[:64])
        lib.porla_scalar_mul_batch_device(C.c_void_p(tab1.handle), C.c_void_p(d_sc.data_ptr()), len(P), fmt, pb.POINT_BE64,
                                          C.c_void_p(d_out.data_ptr()), None)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().tobytes()
        for i, k in enumerate(ks):
            W = O.mul(c, k % c.n, P[0])
            want = bytes(64) if W is None else W[0].to_bytes(32, "big") + W[1].to_bytes(32, "big")
            assert got[64 * i:64 * i + 64] == want, (curve, "fixed", i)
        tab1.destroy()
