"""The binary-GCD modular inversion of the affine bucket accumulation (porla_b200/csrc/fp_inv.cuh) is plain C++: the same
source is compiled for the host here and compared with Python's pow(x, -1, p) on both base fields -- random values, values
near 0 and p, powers of two, and the Montgomery-form wrapper.  (The device build of the same code is covered by the GPU
parity tests of the affine accumulation.)"""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle import curves_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "libfp_inv_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "native", "fp_inv_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def _limbs(v):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def _run(fn, curve, values):
    n = len(values)
    arr = (C.c_uint32 * (8 * n))(*[w for v in values for w in _limbs(v)])
    out = (C.c_uint32 * (8 * n))()
    fn(curve, arr, out, n)
    return [sum(out[8 * i + k] << (32 * k) for k in range(8)) for i in range(n)]


@pytest.mark.parametrize("curve,p", [(0, O.BN254.p), (1, O.SECP256K1.p)])
def test_plain_inverse_matches_python(lib, curve, p):
    rnd = random.Random(curve)
    vals = [1, 2, 3, p - 1, p - 2, (p + 1) // 2, (p - 1) // 2, 1 << 255 if (1 << 255) < p else 1 << 253, (1 << 200) + 1, 0xFFFFFFFF,
            1 << 32, (1 << 64) - 1, p >> 1, p >> 30, 0x40000000, 0x3FFFFFFF]
    vals += [rnd.randrange(1, p) for _ in range(3000)]
    vals += [rnd.randrange(1, 1 << rnd.randrange(1, 256)) % p or 1 for _ in range(1000)]     # short values
    vals += [p - (rnd.randrange(1, 1 << rnd.randrange(1, 200))) for _ in range(500)]         # just below p
    got = _run(lib.fp_inv_plain, curve, vals)
    for v, g in zip(vals, got):
        assert g == pow(v, -1, p), hex(v)
    assert _run(lib.fp_inv_plain, curve, [0]) == [0]


def test_internal_form_inverse(lib):
    rnd = random.Random(9)
    p = O.BN254.p
    R = 1 << 256
    plain = [rnd.randrange(1, p) for _ in range(500)]
    got = _run(lib.fp_inv_internal, 0, [v * R % p for v in plain])          # Montgomery in, Montgomery out
    for v, g in zip(plain, got):
        assert g == pow(v, -1, p) * R % p
    p = O.SECP256K1.p
    plain = [rnd.randrange(1, p) for _ in range(500)]
    got = _run(lib.fp_inv_internal, 1, plain)                               # plain representation
    for v, g in zip(plain, got):
        assert g == pow(v, -1, p)
