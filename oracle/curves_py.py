"""Independent big-integer oracle for the two curves on Porla's MSM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``porla_b200/`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the fixture generators.

It restates, with Python integers and the textbook affine/Jacobian formulas, what the
reference computes at the boundary:

* BN254 G1 (KZG mode): ``/root/reference/porla/main.go:119-138`` (``compute_multi_exp``),
  ``:104-116`` (``compute_digest_from_srs`` = ``kzg.Commit``), ``:154-175`` (``create_proof``),
  ``:196-230`` (single-point ops), ``:43-68`` (SRS blob).  gnark-crypto v0.6.0 itself is NOT in
  ``/root/reference`` (Go module, fetched by ``auto_setup.sh:44-51``) -> **parity unpinned**:
  the byte formats follow SURVEY.md Appendix B; the group law is the public curve
  y^2 = x^3 + 3 over p (``PARITY: vs restatement, never vs gnark``).
* secp256k1 (IPA mode): ``porla/Utils/secp256k1_lib/ecmult_impl.h:814-860``
  (``secp256k1_ecmult_multi_var``) and the SEC1 codec ``eckey_impl.h:36-52``.  This one IS
  pinned: ``oracle/_ref`` compiles the vendored C and the tests compare all three.

Because every result is a canonical affine point, any correct MSM algorithm is bit-exact
after serialisation; this file therefore uses the simplest possible algorithms.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass


@dataclass(frozen=True)
class Curve:
    name: str
    p: int          # base field
    n: int          # group order (scalar field)
    b: int          # y^2 = x^3 + b
    gx: int
    gy: int


BN254 = Curve(
    "bn254",
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    3, 1, 2,
)
SECP256K1 = Curve(
    "secp256k1",
    2**256 - 2**32 - 977,
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    7,
    0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
)

INF = None  # affine point at infinity


# --------------------------------------------------------------------------- group law
def on_curve(c: Curve, P) -> bool:
    if P is INF:
        return True
    x, y = P
    return (y * y - x * x * x - c.b) % c.p == 0


def neg(c: Curve, P):
    if P is INF:
        return INF
    return (P[0], (-P[1]) % c.p)


def add(c: Curve, P, Q):
    """Textbook affine chord-and-tangent (a = 0 curves)."""
    if P is INF:
        return Q
    if Q is INF:
        return P
    x1, y1 = P
    x2, y2 = Q
    p = c.p
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return (x3, (lam * (x1 - x3) - y1) % p)


def _jdbl(c, X, Y, Z):
    p = c.p
    if Y == 0 or Z == 0:
        return (1, 1, 0)
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    X3 = (E * E - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def _jadd_affine(c, X1, Y1, Z1, x2, y2):
    p = c.p
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % p
    U2 = x2 * Z1Z1 % p
    S2 = y2 * Z1 * Z1Z1 % p
    H = (U2 - X1) % p
    R = (S2 - Y1) % p
    if H == 0:
        if R == 0:
            return _jdbl(c, X1, Y1, Z1)
        return (1, 1, 0)
    HH = H * H % p
    HHH = H * HH % p
    V = X1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - Y1 * HHH) % p
    Z3 = Z1 * H % p
    return (X3, Y3, Z3)


def _jaffine(c, X, Y, Z):
    if Z == 0:
        return INF
    zi = pow(Z, -1, c.p)
    zi2 = zi * zi % c.p
    return (X * zi2 % c.p, Y * zi2 * zi % c.p)


def mul(c: Curve, k: int, P):
    """k*P, left-to-right double-and-add in Jacobian coordinates; k is reduced mod n."""
    k %= c.n
    if P is INF or k == 0:
        return INF
    X, Y, Z = 1, 1, 0
    for bit in bin(k)[2:]:
        X, Y, Z = _jdbl(c, X, Y, Z)
        if bit == "1":
            X, Y, Z = _jadd_affine(c, X, Y, Z, P[0], P[1])
    return _jaffine(c, X, Y, Z)


def msm_naive(c: Curve, scalars, points):
    """sum_i s_i * P_i by independent scalar multiplications (ground truth, slow)."""
    acc = INF
    for s, P in zip(scalars, points):
        acc = add(c, acc, mul(c, s, P))
    return acc


def msm(c: Curve, scalars, points, window: int = 8):
    """Bucket method with unsigned digits; cross-checked against msm_naive in the tests."""
    scalars = [s % c.n for s in scalars]
    nbits = c.n.bit_length()
    nwin = (nbits + window - 1) // window
    total = (1, 1, 0)
    for w in reversed(range(nwin)):
        for _ in range(window):
            total = _jdbl(c, *total)
        buckets = [(1, 1, 0)] * (1 << window)
        for s, P in zip(scalars, points):
            if P is INF:
                continue
            d = (s >> (w * window)) & ((1 << window) - 1)
            if d:
                buckets[d] = _jadd_affine(c, *buckets[d], P[0], P[1])
        run = INF
        wsum = INF
        for d in range((1 << window) - 1, 0, -1):
            run = add(c, run, _jaffine(c, *buckets[d]))
            wsum = add(c, wsum, run)
        if wsum is not INF:
            total = _jadd_affine(c, *total, wsum[0], wsum[1])
    return _jaffine(c, *total)


def sqrt_mod(c: Curve, a: int):
    """Square root for p = 3 mod 4 (both curves); None if a is a non-residue."""
    a %= c.p
    y = pow(a, (c.p + 1) // 4, c.p)
    return y if y * y % c.p == a else None


# --------------------------------------------------------------------------- BN254 codecs
# gnark-crypto v0.6.0 layout ([memory]): 2 flag bits in byte 0.
M_UNCOMPRESSED = 0x00
M_COMPRESSED_INF = 0x40
M_COMPRESSED_SMALLEST = 0x80
M_COMPRESSED_LARGEST = 0xC0


def fr_set_bytes(b: bytes) -> int:
    """fr.Element.SetBytes: big-endian integer of any length reduced mod r (main.go:127)."""
    return int.from_bytes(b, "big") % BN254.n


def fr_marshal(x: int) -> bytes:
    return (x % BN254.n).to_bytes(32, "big")


def bn254_marshal(P) -> bytes:
    """G1Affine.Marshal() = RawBytes(): X||Y big-endian, infinity = 64 zero bytes (main.go:137)."""
    if P is INF:
        return bytes(64)
    return P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big")


def bn254_compress(P) -> bytes:
    """G1Affine.Bytes(): 32-byte compressed form used only inside the SRS blob (main.go:48)."""
    if P is INF:
        return bytes([M_COMPRESSED_INF]) + bytes(31)
    x, y = P
    flag = M_COMPRESSED_LARGEST if y > (BN254.p - 1) // 2 else M_COMPRESSED_SMALLEST
    out = bytearray(x.to_bytes(32, "big"))
    out[0] |= flag
    return bytes(out)


def bn254_unmarshal(b: bytes):
    """G1Affine.Unmarshal / SetBytes.  Accepts 64-byte uncompressed and 32-byte compressed.

    Coordinates are reduced mod p as fp.Element.SetBytes does.  64 zero bytes -> infinity.
    No curve check is enforced (main.go ignores the error, main.go:130)."""
    flag = b[0] & 0xC0
    if flag == M_UNCOMPRESSED:
        x = int.from_bytes(b[:32], "big") % BN254.p
        y = int.from_bytes(b[32:64], "big") % BN254.p
        if x == 0 and y == 0:
            return INF
        return (x, y)
    if flag == M_COMPRESSED_INF:
        return INF
    xb = bytearray(b[:32])
    xb[0] &= 0x3F
    x = int.from_bytes(xb, "big") % BN254.p
    y = sqrt_mod(BN254, x * x * x + 3)
    if y is None:
        raise ValueError("not on curve")
    if (y > (BN254.p - 1) // 2) != (flag == M_COMPRESSED_LARGEST):
        y = BN254.p - y
    return (x, y)


# --------------------------------------------------------------------------- secp256k1 codecs
def secp_sec1_compressed(P) -> bytes:
    """secp256k1_eckey_pubkey_serialize(compressed=1), eckey_impl.h:36-52.  Infinity has no
    SEC1 encoding (the reference returns 0); we use 33 zero bytes as the test sentinel."""
    if P is INF:
        return bytes(33)
    return bytes([0x03 if P[1] & 1 else 0x02]) + P[0].to_bytes(32, "big")


def secp_sec1_uncompressed(P) -> bytes:
    if P is INF:
        return bytes(65)
    return b"\x04" + P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big")


# --------------------------------------------------------------------------- synthetic inputs
def hash_point(c: Curve, i: int, tag: bytes = b"porla-pt"):
    """x <- SHA-256(tag || LE64(i) || LE32(ctr)) mod p until x^3+b is a QR;
    BN254 takes the root y <= (p-1)/2, secp256k1 the even root."""
    ctr = 0
    while True:
        h = hashlib.sha256(tag + i.to_bytes(8, "little") + ctr.to_bytes(4, "little")).digest()
        x = int.from_bytes(h, "big") % c.p
        y = sqrt_mod(c, x * x * x + c.b)
        if y is not None:
            if c is BN254:
                if y > (c.p - 1) // 2:
                    y = c.p - y
            elif y & 1:
                y = c.p - y
            return (x, y)
        ctr += 1


def hash_scalar(c: Curve, i: int, tag: bytes = b"porla-sc") -> int:
    return int.from_bytes(hashlib.sha256(tag + i.to_bytes(8, "little")).digest(), "big") % c.n


# --------------------------------------------------------------------------- KZG (BN254)
def kzg_srs_g1(tau: int, n: int):
    """kzg.NewSRS: G1[i] = [tau^i] G (main.go:46)."""
    G = (BN254.gx, BN254.gy)
    out, t = [], 1
    for _ in range(n):
        out.append(mul(BN254, t, G))
        t = t * tau % BN254.n
    return out


def kzg_commit(coeffs, srs):
    """kzg.Commit = MultiExp(srs.G1[:len(p)], p) (main.go:114)."""
    return msm(BN254, coeffs, srs[: len(coeffs)])


def kzg_open(coeffs, z: int):
    """kzg.Open: y = f(z) by Horner; h = (f - y)/(X - z) by synthetic division (main.go:170).
    Returns (y, h) with len(h) == len(coeffs) - 1."""
    r = BN254.n
    n = len(coeffs)
    y = 0
    for cf in reversed(coeffs):
        y = (y * z + cf) % r
    h = [0] * (n - 1)
    carry = 0
    for i in range(n - 1, 0, -1):
        carry = (coeffs[i] + carry * z) % r
        h[i - 1] = carry
    return y, h


# ---------------------------------------------------------------------------------------------
# GLV endomorphism constants (test infrastructure for porla_b200/csrc/ec.cuh: Bn254::glv_*).
# phi(x, y) = (beta x, y) = lambda (x, y); (a1, b1), (a2, b2) is a reduced basis of the lattice
# {(x, y): x + y lambda = 0 mod n}; k = k1 + k2 lambda with k1 = k - c1 a1 - c2 a2, k2 = -c1 b1 - c2 b2.
def glv_constants(c: Curve):
    import math

    def cube_root_of_unity(m):
        for g in range(2, 100):
            w = pow(g, (m - 1) // 3, m)
            if w != 1:
                return w
        raise ValueError

    beta, lam = cube_root_of_unity(c.p), cube_root_of_unity(c.n)
    G = (c.gx, c.gy)
    found = None
    for lm in (lam, lam * lam % c.n):
        lg = mul(c, lm, G)
        for bt in (beta, beta * beta % c.p):
            if lg == (bt * G[0] % c.p, G[1]):
                found = (bt, lm)
    # the smaller lambda / its matching beta (the pair the kernels were generated with)
    beta, lam = found
    alt = (beta * beta % c.p, lam * lam % c.n)
    if alt[1] < lam:
        beta, lam = alt
    rows, sq = [], math.isqrt(c.n)
    r0, r1, t0, t1 = c.n, lam, 0, 1
    while r1:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        t0, t1 = t1, t0 - q * t1
        rows.append((r0, t0))
    idx = max(i for i, (r, _) in enumerate(rows) if r >= sq)
    v1 = (rows[idx + 1][0], -rows[idx + 1][1])
    ca, cb = (rows[idx][0], -rows[idx][1]), (rows[idx + 2][0], -rows[idx + 2][1])
    v2 = ca if ca[0] ** 2 + ca[1] ** 2 <= cb[0] ** 2 + cb[1] ** 2 else cb
    return {"beta": beta, "lambda": lam, "a1": v1[0], "b1": v1[1], "a2": v2[0], "b2": v2[1]}


def glv_split(c: Curve, k: int, consts=None):
    """The device routine glv_split (msm_kernels.cuh) restated: floors instead of roundings."""
    g = consts or glv_constants(c)
    g1 = (g["b2"] << 256) // c.n
    g2 = ((-g["b1"]) << 256) // c.n
    c1, c2 = (k * g1) >> 256, (k * g2) >> 256
    k1 = k - c1 * g["a1"] - c2 * g["a2"]
    k2 = -c1 * g["b1"] - c2 * g["b2"]
    return k1, k2
