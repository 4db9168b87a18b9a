"""ctypes loaders for the compiled oracles.  TEST INFRASTRUCTURE ONLY.

* ``bn254()``    -> oracle/liboracle_bn254.so   (C restatement; rebuilt with gcc if missing)
* ``secp_ref()`` -> oracle/_ref/libsecp_ref.so  (the reference's vendored secp256k1, built in the
                    build container from /root/reference; on the GPU box only the prebuilt file
                    exists -- returns None if it is absent)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_bn = None
_secp = None


def build() -> None:
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def bn254() -> C.CDLL:
    global _bn
    if _bn is None:
        path = os.path.join(HERE, "liboracle_bn254.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", HERE, "liboracle_bn254.so"], check=True, stdout=subprocess.DEVNULL)
        lib = C.CDLL(path)
        P = C.c_char_p
        lib.oracle_bn254_msm.argtypes = [P, P, C.c_size_t, P, C.c_int]
        lib.oracle_bn254_msm.restype = None
        lib.oracle_bn254_add.argtypes = [P, P, P]
        lib.oracle_bn254_mul.argtypes = [P, P, P]
        lib.oracle_bn254_point_chain.argtypes = [P, P, C.c_size_t, P]
        _bn = lib
    return _bn


def bn254_msm(scalars: bytes, points: bytes, n: int, nthreads: int = 1) -> bytes:
    out = C.create_string_buffer(64)
    bn254().oracle_bn254_msm(scalars, points, n, out, nthreads)
    return out.raw


def bn254_point_chain(base: bytes, step: bytes, n: int) -> bytes:
    out = C.create_string_buffer(64 * n)
    bn254().oracle_bn254_point_chain(base, step, n, out)
    return out.raw


def secp_ref():
    global _secp
    if _secp is None:
        path = os.path.join(HERE, "_ref", "libsecp_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/porla/Utils/secp256k1_lib"):
                subprocess.run(["make", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
            if not os.path.exists(path):
                return None
        lib = C.CDLL(path)
        P = C.c_char_p
        lib.ref_secp_msm.argtypes = [P, P, C.c_size_t, P, P]
        lib.ref_secp_msm.restype = C.c_int
        lib.ref_secp_prepare.argtypes = [P, P, C.c_size_t]
        lib.ref_secp_prepare.restype = C.c_void_p
        lib.ref_secp_release.argtypes = [C.c_void_p]
        lib.ref_secp_msm_prepared.argtypes = [C.c_void_p, C.c_size_t, C.c_int, P, P]
        lib.ref_secp_msm_prepared.restype = C.c_int
        lib.ref_secp_kat.argtypes = [P, P]
        lib.ref_secp_kat.restype = C.c_size_t
        lib.ref_secp_point_chain.argtypes = [P, C.c_size_t, P]
        if hasattr(lib, "ref_sha256_sequence"):
            lib.ref_sha256_sequence.argtypes = [P, C.POINTER(C.c_uint32), C.c_int, P]
        for f in ("ref_secp_sizeof_ge", "ref_secp_sizeof_gej", "ref_secp_sizeof_scalar"):
            getattr(lib, f).restype = C.c_size_t
        _secp = lib
    return _secp


def secp_ref_msm(scalars_le: bytes, points_be: bytes, n: int):
    """(ok, X||Y 64 B, SEC1 33 B) through secp256k1_ecmult_multi_var of the reference."""
    lib = secp_ref()
    o64, o33 = C.create_string_buffer(64), C.create_string_buffer(33)
    ok = lib.ref_secp_msm(scalars_le, points_be, n, o64, o33)
    return ok, o64.raw, o33.raw


def secp_ref_sha256_sequence(segments):
    """Digests of the reference's re-finalized SHA-256 object after each segment (see secp_ref.c); None if the
    reference build is not available."""
    lib = secp_ref()
    if lib is None or not hasattr(lib, "ref_sha256_sequence"):
        return None
    lens = (C.c_uint32 * len(segments))(*[len(s) for s in segments])
    outs = C.create_string_buffer(32 * len(segments))
    lib.ref_sha256_sequence(b"".join(segments), lens, len(segments), outs)
    return [outs.raw[32 * i:32 * i + 32] for i in range(len(segments))]
