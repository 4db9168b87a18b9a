"""Host-only entry points of the C-ABI (no GPU needed): single-point operations, codecs, the
trapdoor digest and the KZG verifier, against the oracle -- mirroring how Client.hpp/Server.hpp
call them (utils.h:235-305)."""
import random

import porla_b200 as pb
from oracle import curves_py as O
from tests.common import be, bn254_points, golden

BN = O.BN254
G = (1, 2)
TAU = bytes.fromhex("ffeeddccbbaa99887766554433221100")     # config.hpp:39
ALPHA = bytes.fromhex("00112233445566778899aabbccddeeff")   # config.hpp:38


def test_point_ops_match_oracle():
    rnd = random.Random(5)
    pts = bn254_points(6)
    for t in range(12):
        a, b = rnd.choice(pts), rnd.choice(pts)
        if t == 0:
            b = a
        if t == 1:
            b = O.neg(BN, a)
        if t == 2:
            b = None
        if t == 3:
            a = None
        buf = bytearray(O.bn254_marshal(a))
        pb.bn254_add(buf, O.bn254_marshal(b))
        assert bytes(buf) == O.bn254_marshal(O.add(BN, a, b)), t
        k = rnd.randrange(1 << 256)                # >= r: reduced by fr.SetBytes (main.go:208)
        buf = bytearray(O.bn254_marshal(a))
        pb.bn254_mult(buf, be(k))
        assert bytes(buf) == O.bn254_marshal(O.mul(BN, k, a)), t
        buf = bytearray(O.bn254_marshal(a))
        pb.bn254_neg(buf)
        assert bytes(buf) == O.bn254_marshal(O.neg(BN, a))
    buf = bytearray(b"\xff" * 64)
    pb.bn254_set_infinity(buf)
    assert bytes(buf) == bytes(64)
    assert pb.bn254_compare(O.bn254_marshal(pts[0]), O.bn254_marshal(pts[0]))
    assert pb.bn254_scalar_set_int(0x01020304) == bytes(28) + b"\x01\x02\x03\x04"


def test_small_int_scalar_is_big_endian_word7():
    # utils.h:271-275 + main.go:208: bn254_mult(P, set_int(v)) == v*P
    P = bn254_points(1)[0]
    buf = bytearray(O.bn254_marshal(P))
    pb.bn254_mult(buf, pb.bn254_scalar_set_int(77777))
    assert bytes(buf) == O.bn254_marshal(O.mul(BN, 77777, P))


def test_srs_blob_digest_and_verify_on_cpu():
    g = golden("bn254.json")["kzg"]
    n = g["n"]
    k = pb.Kzg(TAU, ALPHA)
    blob = k.init_srs(n)
    assert len(blob) == n * 32 + 132                       # Client.hpp:350: NUM_CHUNKS*32+132
    assert blob[128:132] == n.to_bytes(4, "big")
    assert blob[132:].hex() == g["srs_compressed"]
    # compute_digest = [alpha * f(tau)] G  (main.go:71-89)
    rnd = random.Random(1)
    coef = [rnd.randrange(1 << 256) for _ in range(n)]
    tau, alpha = int.from_bytes(TAU, "big"), int.from_bytes(ALPHA, "big")
    fx = sum((cf % BN.n) * pow(tau, i, BN.n) for i, cf in enumerate(coef)) % BN.n
    assert k.compute_digest(b"".join(map(be, coef))) == O.bn254_marshal(O.mul(BN, fx * alpha, G))
    # complement = [s] h_MAC is linear in s even though h_MAC is random (main.go:92-101)
    c1 = k.compute_digest_complement(be(5))
    c2 = k.compute_digest_complement(be(7))
    c3 = bytearray(c1)
    pb.bn254_add(c3, c2)
    assert bytes(c3) == k.compute_digest_complement(be(12))
    # verifier accepts the oracle's opening and rejects a wrong claim (main.go:178-193)
    args = (bytes.fromhex(g["commit"]), bytes.fromhex(g["H"]), be(g["z"]), bytes.fromhex(g["claim"]))
    assert k.verify_proof(*args)
    bad = (int.from_bytes(args[3], "big") + 1) % BN.n
    assert not k.verify_proof(args[0], args[1], args[2], be(bad))
    # the server side parses the same blob (Server.hpp:179-188)
    k2 = pb.Kzg(TAU, ALPHA)
    k2.init_srs_from_data(n, blob)
    assert k2.verify_proof(*args)


def test_pairing_selfcheck_and_tampered_proofs():
    """The verifier's pairing (multi-pairing Miller loop, u-addition-chain hard part) against its plain
    counterparts on bilinearity identities, and kzg.Verify rejecting every tampered component."""
    assert pb.load().porla_debug_pairing_selfcheck(3) == 0
    g = golden("bn254.json")["kzg"]
    k = pb.Kzg(TAU, ALPHA)
    k.init_srs(g["n"])
    c, h, z, y = bytes.fromhex(g["commit"]), bytes.fromhex(g["H"]), be(g["z"]), bytes.fromhex(g["claim"])
    assert k.verify_proof(c, h, z, y)
    other = O.bn254_marshal(O.mul(BN, 31337, G))
    assert not k.verify_proof(other, h, z, y)
    assert not k.verify_proof(c, other, z, y)
    assert not k.verify_proof(c, h, be(g["z"] + 1), y)
    assert not k.verify_proof(c, bytes(64), z, y)               # H = infinity


def test_msm_plan_codes_carry_the_window_layout():
    """porla_msm_plan needs no GPU: the plan code (window size | GLV on/off) that all ranks of a sharded MSM pass back,
    and the number of window sums it implies."""
    import ctypes as C
    import porla_b200 as pb
    lib = pb.load()
    c, nwin = C.c_int(0), C.c_int(0)
    lib.porla_msm_plan(pb.CURVE_BN254, 1 << 20, 1, 0, C.byref(c), C.byref(nwin))
    assert c.value & 0xFF == 16 and c.value & 0x100 and nwin.value == 8          # GLV: 8 windows per 127-bit half
    code = c.value
    lib.porla_msm_plan(pb.CURVE_BN254, (1 << 20) - 12345, 1, code, C.byref(c), C.byref(nwin))
    assert c.value == code and nwin.value == 8                                   # a rank with another n keeps the layout
    lib.porla_msm_plan(pb.CURVE_BN254, 1 << 20, 1, 16 | 0x200, C.byref(c), C.byref(nwin))
    assert c.value == 16 | 0x200 and nwin.value == 16
    lib.porla_msm_plan(pb.CURVE_BN254, 1 << 24, 1, 0, C.byref(c), C.byref(nwin))
    assert c.value == 20 | 0x200 and nwin.value == 13
    lib.porla_msm_plan(pb.CURVE_SECP256K1, 1 << 18, 1, 0, C.byref(c), C.byref(nwin))
    assert c.value == 16 | 0x200 and nwin.value == 16


def test_mult_point_glv_corners():
    """mult_point (main.go:205-214) runs a GLV joint double-and-add on the host: corners of the scalar split, short
    audit coefficients, unreduced scalars, negated halves, the point at infinity."""
    import random
    import porla_b200 as pb
    from oracle import curves_py as O
    BN = O.BN254
    g = O.glv_constants(BN)
    lam = g["lambda"]
    rnd = random.Random(205)
    P = O.hash_point(BN, 11)
    ks = [0, 1, 2, 3, BN.n - 1, BN.n, BN.n + 7, lam, lam - 1, lam + 1, BN.n - lam, g["a2"], -g["b1"], (1 << 127) - 1, 1 << 127,
          (1 << 31) - 1, (1 << 256) - 1, BN.n // 2, BN.n // 2 + 1] + [rnd.randrange(1 << 256) for _ in range(40)]
    for k in ks:
        buf = bytearray(O.bn254_marshal(P))
        pb.bn254_mult(buf, k.to_bytes(32, "big"))
        assert bytes(buf) == O.bn254_marshal(O.mul(BN, k % BN.n, P)), hex(k)
    buf = bytearray(64)
    pb.bn254_mult(buf, (12345).to_bytes(32, "big"))
    assert bytes(buf) == bytes(64)
