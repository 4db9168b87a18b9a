"""TEST INFRASTRUCTURE (oracle): restatement of Porla's IPA-mode prover,
/root/reference/porla/Server/Server.hpp:2279-2443 (Server::inner_product_prove), on Python integers.

Only tests/ may import this file.  The elliptic-curve arithmetic is oracle/curves_py.py (pinned to the reference's
own secp256k1, see tests/test_oracle.py); the Fiat-Shamir transcript below reproduces how the reference drives ONE
secp256k1_sha256 object -- written to and finalized repeatedly without re-initialisation
(/root/reference/porla/Utils/secp256k1_lib/hash_impl.h:151-165: finalize zeroes the eight state words but keeps the byte
counter) -- and is pinned against that very code compiled into oracle/_ref (tests/test_oracle.py).
"""
import struct

from . import curves_py as O

_K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2,
]
_IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
_M = 0xFFFFFFFF


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & _M


def _compress(s, block):
    w = list(struct.unpack(">16I", block))
    for i in range(16, 64):
        s0 = _rotr(w[i - 15], 7) ^ _rotr(w[i - 15], 18) ^ (w[i - 15] >> 3)
        s1 = _rotr(w[i - 2], 17) ^ _rotr(w[i - 2], 19) ^ (w[i - 2] >> 10)
        w.append((w[i - 16] + s0 + w[i - 7] + s1) & _M)
    a, b, c, d, e, f, g, h = s
    for i in range(64):
        t1 = (h + (_rotr(e, 6) ^ _rotr(e, 11) ^ _rotr(e, 25)) + ((e & f) ^ (~e & _M & g)) + _K[i] + w[i]) & _M
        t2 = ((_rotr(a, 2) ^ _rotr(a, 13) ^ _rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & _M
        h, g, f, e, d, c, b, a = g, f, e, (d + t1) & _M, c, b, a, (t1 + t2) & _M
    return [(x + y) & _M for x, y in zip(s, (a, b, c, d, e, f, g, h))]


class TranscriptSha256:
    """secp256k1_sha256 as the reference uses it: initialize once, then write / finalize any number of times
    (hash_impl.h:37-48, 132-149, 151-165)."""

    def __init__(self):
        self.s = list(_IV)
        self.buf = b""
        self.bytes = 0

    def write(self, data: bytes):
        self.bytes += len(data)
        self.buf += data
        while len(self.buf) >= 64:
            self.s = _compress(self.s, self.buf[:64])
            self.buf = self.buf[64:]

    def finalize(self) -> bytes:
        size = struct.pack(">II", (self.bytes >> 29) & _M, (self.bytes << 3) & _M)
        self.write(b"\x80" + bytes((119 - (self.bytes % 64)) % 64))
        self.write(size)
        out = struct.pack(">8I", *self.s)
        self.s = [0] * 8                       # hash_impl.h:162: the state is wiped, the byte counter is not
        return out


def _le_words(v: int) -> bytes:
    """convert_ZZ_to_arr (utils.h:353-364): eight 32-bit words, least significant first, native (little-endian) bytes."""
    return (v % (1 << 256)).to_bytes(32, "little")


def _from_le_words(b: bytes) -> int:
    """convert_arr_to_ZZ_p (utils.h:384-393)."""
    return int.from_bytes(b, "little")


SEED = b"hash of P, c, etc. all that jazz"     # Server.hpp:2284 (32 bytes are hashed)


def inner_product_prove(generators, u, a, b):
    """Server::inner_product_prove (Server.hpp:2279-2443).  generators: NUM_CHUNKS affine points, u: affine point,
    a, b: NUM_CHUNKS integers.  Returns the proof bytes: <a, b> | (L, R) per round | a0 b0 a1 b1."""
    c = O.SECP256K1
    n = c.n
    N = len(generators)
    a, b = [x % n for x in a], [x % n for x in b]
    proof = bytearray(_le_words(sum(x * y for x, y in zip(a, b)) % n))        # :2286-2288
    x_values = [1] * N                                                          # :2299-2301
    sha = TranscriptSha256()                                                    # :2306-2310
    sha.write(SEED[:32])
    sha.write(bytes(proof[:32]))
    random_str = sha.finalize()
    half, k = N // 2, 1
    while half > 1:                                                             # :2318
        x = _from_le_words(random_str) % n                                      # :2320-2323
        inv_x = pow(x, -1, n)
        cL = sum(a[i] * b[half + i] for i in range(half)) % n                   # :2326-2329
        cR = sum(a[half + i] * b[i] for i in range(half)) % n                   # :2331-2334
        for odd, a_off, cc, factor in ((1, 0, cL, x), (0, half, cR, inv_x)):    # L :2337-2389, R :2392-2432
            sc, pts = [], []
            for i in range(k):
                pos = 2 * i + odd
                for q, j in enumerate(range(pos * half, (pos + 1) * half)):
                    sc.append(a[a_off + q] * x_values[j] % n)
                    pts.append(generators[j])
                    x_values[j] = x_values[j] * factor % n
            point = O.add(c, O.msm(c, sc, pts), O.mul(c, cc, u))                # :2372-2378
            ser = O.secp_sec1_compressed(point)                                 # :2380-2382
            proof += ser
            sha.write(ser)                                                      # :2385-2387
            random_str = sha.finalize()
        a = [(a[i] * x + a[i + half] * inv_x) % n for i in range(half)] + a[half:]      # :2435-2436
        b = [(b[i] * inv_x + b[i + half] * x) % n for i in range(half)] + b[half:]      # :2439-2440
        half >>= 1
        k <<= 1
    for i in range(2):                                                          # :2443-2449
        proof += _le_words(a[i]) + _le_words(b[i])
    return bytes(proof)


def inner_product_verify(generators, u, commitment, proof) -> bool:
    """Client::inner_product_verify (Client.hpp:1465-1630): True when the two sides the reference compares with
    ge_equals_ge are the same point.  commitment: the affine point sum a_i g_i (or None for infinity)."""
    c = O.SECP256K1
    n = c.n
    N = len(generators)
    pos_ = 32
    ip = _from_le_words(proof[:32])                                           # :1479 convert_arr_to_scalar
    lhs = O.add(c, commitment, O.mul(c, ip % n, u))                           # :1481-1482
    x_values = [1] * N
    sha = TranscriptSha256()                                                  # :1493-1497
    sha.write(SEED[:32])
    sha.write(bytes(proof[:32]))
    random_str = sha.finalize()
    half, k = N // 2, 1
    while half > 1:                                                           # :1504
        x = _from_le_words(random_str) % n
        inv_x = pow(x, -1, n)
        for i in range(k):                                                    # :1511-1523
            for j in range((2 * i + 1) * half, (2 * i + 2) * half):
                x_values[j] = x_values[j] * x % n
            for j in range(2 * i * half, (2 * i + 1) * half):
                x_values[j] = x_values[j] * inv_x % n
        x2 = x * x % n                                                        # :1525-1528
        inv_x2 = pow(x2, -1, n)
        for factor in (x2, inv_x2):                                           # L :1534-1545, R :1547-1558
            ser = bytes(proof[pos_:pos_ + 33])
            px = int.from_bytes(ser[1:], "big")
            py = O.sqrt_mod(c, (px * px * px + c.b) % c.p)
            if py is None or ser[0] not in (2, 3):
                return False
            if (py & 1) != (ser[0] & 1):
                py = c.p - py
            sha.write(ser)
            random_str = sha.finalize()
            pos_ += 33
            lhs = O.add(c, lhs, O.mul(c, factor, (px, py)))                    # :1560-1561
        half >>= 1
        k <<= 1
    a0, b0, a1, b1 = (_from_le_words(proof[pos_ + 32 * i:pos_ + 32 * i + 32]) for i in range(4))   # :1567-1574
    ab = (a0 * b0 + a1 * b1) % n
    sc = [a0 * x_values[j] % n for j in range(0, N, 2)] + [a1 * x_values[j] % n for j in range(1, N, 2)]   # :1586-1603
    pts = [generators[j] for j in range(0, N, 2)] + [generators[j] for j in range(1, N, 2)]
    rhs = O.add(c, O.mul(c, ab, u), O.msm(c, sc, pts))                        # :1583, :1609-1626
    return lhs == rhs                                                         # :1628-1630 ge_equals_ge
