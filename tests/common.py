"""Shared deterministic inputs for the tests (same recipes as tests/golden/make_golden.py)."""
import hashlib
import json
import os

from oracle import curves_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def det_scalar(tag: bytes, i: int) -> int:
    return int.from_bytes(hashlib.sha256(tag + i.to_bytes(8, "little")).digest(), "big")


def be(x: int) -> bytes:
    return x.to_bytes(32, "big")


def le(x: int) -> bytes:
    return x.to_bytes(32, "little")


def enc_points(pts) -> bytes:
    return b"".join(bytes(64) if P is None else be(P[0]) + be(P[1]) for P in pts)


_secp_chain_cache = {}


def secp_chain(n: int):
    """P0 = G, P(i+1) = P(i) + q*G with q = sha256('porla-seed')."""
    c = O.SECP256K1
    if n not in _secp_chain_cache:
        q = int.from_bytes(hashlib.sha256(b"porla-seed").digest(), "big")
        Q = O.mul(c, q, (c.gx, c.gy))
        pts, cur = [], (c.gx, c.gy)
        for _ in range(n):
            pts.append(cur)
            cur = O.add(c, cur, Q)
        _secp_chain_cache[n] = pts
    return _secp_chain_cache[n]


_bn_pts = []


def bn254_points(n: int):
    while len(_bn_pts) < n:
        _bn_pts.append(O.hash_point(O.BN254, len(_bn_pts)))
    return _bn_pts[:n]
