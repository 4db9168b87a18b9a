"""Parity of the one-launch small-MSM kernels (porla_b200/csrc/small_kernels.cuh) against the oracle and
against the sort / accumulate / reduce pipeline, through the C-ABI.

  k_small_bits : Porla's audit aggregation, compute_multi_exp with 128..766 terms and 31-bit coefficients
                 (/root/reference/porla/Server/Server.hpp:900-901, Client.hpp:795, main.go:119-138)
  k_lut_sum    : commitments over the resident SRS, compute_digest_from_srs / create_proof
                 (Server.hpp:558, main.go:104-116, 154-175) and the resident secp256k1 generator table

PORLA_NO_SMALL=1 forces the pipeline, so every case is computed twice and both must equal the oracle.
"""
import ctypes as C
import random

import pytest

import porla_b200 as pb
from oracle import curves_py as O
from oracle import loader

pytestmark = pytest.mark.gpu

BN, SE = O.BN254, O.SECP256K1
TAU = bytes.fromhex("ffeeddccbbaa99887766554433221100")
ALPHA = bytes.fromhex("00112233445566778899aabbccddeeff")


def be(x):
    return x.to_bytes(32, "big")


def le(x):
    return x.to_bytes(32, "little")


def enc(P):
    return bytes(64) if P is None else be(P[0]) + be(P[1])


def chain(c, n, seed):
    """n distinct points: P_0 = hash point, P_{i+1} = P_i + Q."""
    cur, Q = O.hash_point(c, seed), O.hash_point(c, seed + 1)
    out = []
    for _ in range(n):
        out.append(cur)
        cur = O.add(c, cur, Q)
    return out


@pytest.fixture(params=["small", "pipeline"])
def route(request, monkeypatch):
    if request.param == "pipeline":
        monkeypatch.setenv("PORLA_NO_SMALL", "1")
    return request.param


@pytest.mark.parametrize("n,bits", [(1, 31), (2, 31), (127, 31), (128, 31), (129, 31), (255, 31), (256, 31), (257, 31),
                                    (513, 31), (766, 31), (766, 256), (100, 256), (300, 64), (1000, 254)])
def test_bitwise_tree_sum_matches_oracle(n, bits, route):
    rnd = random.Random(1000 * n + bits)
    pts = chain(BN, n, n)
    for i in range(0, n, 7):
        pts[i] = None                                       # alignment MACs that are still infinity
    sc = [rnd.randrange(1 << bits) for _ in range(n)]
    if n > 4:
        sc[1] = 0
        sc[2] = (1 << bits) - 1
    got = pb.bn254_multi_exp(b"".join(map(enc, pts)), b"".join(map(be, sc)), n)
    want = loader.bn254_msm(b"".join(map(be, sc)), b"".join(map(enc, pts)), n, 1)
    assert got == want
    if n <= 130:
        assert got == O.bn254_marshal(O.msm(BN, sc, pts))


def test_bitwise_tree_sum_exceptional_additions(route):
    """Equal points meet in the tree (doubling branch), opposite points cancel, whole windows are empty."""
    P, Q = O.hash_point(BN, 5), O.hash_point(BN, 6)
    cases = [
        ([1, 1], [P, P]),
        ([1, 1], [P, O.neg(BN, P)]),
        ([3, 3, 3, 3], [P, P, P, P]),
        ([5, 5, 9, 9], [P, O.neg(BN, P), Q, O.neg(BN, Q)]),
        ([1 << 30, 1], [P, Q]),
        ([0, 0], [P, Q]),
        ([BN.n, BN.n + 1, BN.n - 1], [P, Q, P]),
        ([(1 << 256) - 1], [P]),
        ([7], [None]),
    ]
    for sc, pts in cases:
        got = pb.bn254_multi_exp(b"".join(map(enc, pts)), b"".join(map(be, sc)), len(sc))
        assert got == O.bn254_marshal(O.msm_naive(BN, sc, pts)), sc
    n = 300                                                  # every term the same point and scalar
    got = pb.bn254_multi_exp(enc(P) * n, be(12345) * n, n)
    assert got == O.bn254_marshal(O.mul(BN, 12345 * n, P))


def test_bitwise_tree_sum_batched(route):
    batch, n = 6, 90
    rnd = random.Random(8)
    pts = chain(BN, batch * n, 40)
    sc = [rnd.randrange(1 << 256) for _ in range(batch * n)]
    got = pb.bn254_multi_exp_batch(b"".join(map(enc, pts)), b"".join(map(be, sc)), n, batch)
    for m in range(batch):
        want = loader.bn254_msm(b"".join(map(be, sc[m * n:(m + 1) * n])), b"".join(map(enc, pts[m * n:(m + 1) * n])), n, 1)
        assert got[64 * m:64 * m + 64] == want, m


@pytest.mark.parametrize("n", [1, 5, 200, 766])
def test_secp256k1_small_matches_oracle(n, route):
    rnd = random.Random(n)
    pts = chain(SE, n, 3 * n)
    sc = [rnd.randrange(1 << 256) for _ in range(n)]
    if n > 4:
        sc[0], sc[1], sc[2], sc[3] = SE.n - 1, SE.n, (SE.n + 1) // 2, SE.n // 2      # the n - s recoding's corners
    got = pb.msm_host(pb.CURVE_SECP256K1, b"".join(map(le, sc)), b"".join(map(enc, pts)), n, scalar_fmt=pb.SCALAR_LE32)
    assert got == enc(O.msm(SE, sc, pts))


def _srs_points(blob, n):
    return [O.bn254_unmarshal(blob[132 + 32 * i:164 + 32 * i]) for i in range(n)]


def test_lookup_table_commitments_match_oracle(route):
    """compute_digest_from_srs over the resident 128-point SRS: digits on the signed-window rounding boundary
    (bytes 0x80 / 0x7f / 0x81 / 0xff), unreduced 256-bit data chunks, zeros, r - 1."""
    k = pb.Kzg(TAU, ALPHA)
    n = 128
    srs = _srs_points(k.init_srs(n), n)
    srs_bytes = b"".join(O.bn254_marshal(P) for P in srs)
    rnd = random.Random(99)
    rows = [
        [rnd.randrange(1 << 256) for _ in range(n)],
        [int.from_bytes(bytes([0x80]) * 32, "big") % BN.n] * n,
        [int.from_bytes(bytes([0x80] * 31 + [0x81]), "big") % BN.n, int.from_bytes(bytes([0x7f] * 32), "big") % BN.n] * (n // 2),
        [(1 << 256) - 1, BN.n - 1, BN.n, BN.n + 1, 0, 1, 0x80, 0x8000, 0x7fff, 0x80000000] + [0] * (n - 10),
        [int.from_bytes(bytes(rnd.choice([0x7f, 0x80, 0x81, 0x00, 0xff]) for _ in range(32)), "big") for _ in range(n)],
        [0] * n,
    ]
    for row in rows:
        data = b"".join(map(be, row))
        got = k.compute_digest_from_srs(data)
        assert got == loader.bn254_msm(data, srs_bytes, n, 1)
    assert k.compute_digest_from_srs(b"".join(map(be, rows[0]))) == O.bn254_marshal(O.msm(BN, rows[0], srs))


def test_lookup_table_batch_and_proof(route):
    k = pb.Kzg(TAU, ALPHA)
    n, batch = 128, 37
    srs = _srs_points(k.init_srs(n), n)
    srs_bytes = b"".join(O.bn254_marshal(P) for P in srs)
    rnd = random.Random(5)
    data = b"".join(be(rnd.randrange(1 << 256)) for _ in range(n * batch))
    got = k.compute_digest_from_srs_batch(data, batch)
    for m in range(batch):
        assert got[64 * m:64 * m + 64] == loader.bn254_msm(data[32 * n * m:32 * n * (m + 1)], srs_bytes, n, 1), m
    # create_proof: commitment of f and of the 127-term quotient (a prefix of the table), then verify
    block = data[:32 * n]
    c_, h_, z_, y_ = k.create_proof(987654321, block)
    assert c_ == got[:64]
    f = [int.from_bytes(block[32 * i:32 * i + 32], "big") % BN.n for i in range(n)]
    z = int.from_bytes(z_, "big")
    y, q = O.kzg_open(f, z)
    assert int.from_bytes(y_, "big") == y
    assert h_ == O.bn254_marshal(O.msm(BN, q, srs[:len(q)]))
    assert k.verify_proof(c_, h_, z_, y_)


def test_secp256k1_generator_table_lookup(route):
    """Resident secp256k1 generators (IPA mode, Client.hpp:377-406): sub-range MSMs with host scalars."""
    import ctypes as C
    n = 128
    rnd = random.Random(17)
    gens = chain(SE, n, 500)
    tab = pb.Table.from_host(pb.CURVE_SECP256K1, b"".join(map(enc, gens)))
    tab.precompute(0, n, 1)
    lib = pb.load()
    for first, cnt in ((0, 128), (0, 16), (16, 16), (0, 1)):
        sc = [rnd.randrange(1 << 256) for _ in range(cnt)]
        sc[0] = SE.n - 1
        out = (C.c_ubyte * 64)()
        buf = b"".join(map(le, sc))
        lib.porla_msm_table_host_scalars(C.c_void_p(tab.handle), first, buf, cnt, pb.SCALAR_LE32, pb.POINT_BE64, C.cast(out, C.c_void_p))
        assert bytes(out) == enc(O.msm(SE, sc, gens[first:first + cnt])), (first, cnt)
    tab.destroy()


# ------------------------------------------------------------------ Server::audit's block aggregation
@pytest.mark.parametrize("n", [1, 3, 128, 766])
def test_audit_aggregate_matches_reference_arithmetic(n):
    """B = sum coef_i * block_i over plain integers (Server.hpp:790-828), then align_MAC on B (Server.hpp:531-560):
    B % PRIME_MODULUS comes back as bn254_scalars, with the commitment of c = (B % PRIME - B) % r over the SRS."""
    PRIME = 207 * 2**248 + 1                                   # utils.h:40
    LCM = PRIME * BN.n                                         # utils.h:42-43
    chunks = 128
    k = pb.Kzg(TAU, ALPHA)
    srs = _srs_points(k.init_srs(chunks), chunks)
    srs_bytes = b"".join(O.bn254_marshal(P) for P in srs)
    rnd = random.Random(31 * n)
    coefs = [rnd.randrange(1 << 31) for _ in range(n)]
    blocks = [[rnd.randrange(LCM) for _ in range(chunks)] for _ in range(n)]
    coefs[0] = (1 << 31) - 1
    blocks[0][:5] = [LCM - 1, 0, PRIME, PRIME - 1, 1]
    if n > 2:
        blocks[1] = [rnd.randrange(1 << 256) for _ in range(chunks)]      # a raw U block: 256-bit chunks (Client.hpp:371)
        coefs[2] = 0
    for i in range(n):
        blocks[i][7] = LCM - 1                                  # the widest possible sum in chunk 7
    got_b, got_align = k.audit_aggregate(b"".join(c.to_bytes(4, "little") for c in coefs),
                                         b"".join(v.to_bytes(64, "little") for row in blocks for v in row), n)
    B = [sum(c * row[j] for c, row in zip(coefs, blocks)) for j in range(chunks)]
    for j in range(chunks):
        assert int.from_bytes(got_b[32 * j:32 * j + 32], "big") == B[j] % PRIME, j
    cs = b"".join(be(((b % PRIME) - b) % BN.n) for b in B)
    assert got_align == loader.bn254_msm(cs, srs_bytes, chunks, 1)


# ------------------------------------------------------------------ the IPA prover's round MSMs
def test_ipa_round_msms_over_resident_generators(route):
    """Server::inner_product_prove (Server.hpp:2318-2443): every round's L and R are multi-exponentiations of
    NUM_CHUNKS/2 generators picked in alternating blocks of half_width (L: odd blocks, R: even blocks) with scalars
    a[q] * x_values[j] mod n.  With the generators resident (porla_secp256k1_table_create expands them into the
    look-up table) each L / R is ONE call over the whole table whose unused generators carry a zero scalar."""
    N = 128
    rnd = random.Random(2318)
    G = (SE.gx, SE.gy)
    gens = [O.mul(SE, rnd.randrange(1, SE.n), G) for _ in range(N)]
    tab = pb.SecpGenerators(gens)
    a = [rnd.randrange(SE.n) for _ in range(N)]
    x_values = [1] * N
    half, k = N // 2, 1
    while half > 1:
        x = rnd.randrange(1, SE.n)
        inv_x = pow(x, -1, SE.n)
        for odd, a_off, factor in ((1, 0, x), (0, half, inv_x)):            # L then R, as the reference orders them
            sc = [0] * N
            for i in range(k):
                pos = 2 * i + odd
                for q, j in enumerate(range(pos * half, (pos + 1) * half)):
                    sc[j] = a[a_off + q] * x_values[j] % SE.n
                    x_values[j] = x_values[j] * factor % SE.n
            ok, got = tab.multi(0, sc)
            active = [(s, gens[j]) for j, s in enumerate(sc) if s]
            assert len(active) <= N // 2
            assert ok == 1 and got == O.msm(SE, [s for s, _ in active], [P for _, P in active]), (half, odd)
        a = [(a[i] * x + a[i + half] * inv_x) % SE.n for i in range(half)] + [0] * (N - half)
        half >>= 1
        k <<= 1
    # the per-thread sub-ranges of compute_commitment (16 generators per pool thread, Server.hpp:331-360)
    for t in range(8):
        sc = [rnd.randrange(1 << 256) for _ in range(16)]
        ok, got = tab.multi(16 * t, sc)
        assert ok == 1 and got == O.msm(SE, [s % SE.n for s in sc], gens[16 * t:16 * t + 16]), t
    tab.destroy()


# ------------------------------------------------------------------ GLV split in the pipeline (BN254)
@pytest.mark.parametrize("glv", ["on", "off"])
@pytest.mark.parametrize("n,window", [(3, 4), (600, 9), (600, 13), (600, 16), (5000, 0), (5000, 11), (40000, 0)])
def test_glv_pipeline_matches_oracle(n, window, glv, monkeypatch):
    """k = k1 + k2 lambda, 2n terms over (P_i, phi(P_i)) with half the windows: same bytes as the plain recoding and
    as the oracle, for window sizes that do and do not divide 128, scalars at the corners of the split, infinities."""
    monkeypatch.setenv("PORLA_NO_SMALL", "1")
    monkeypatch.setenv("PORLA_GLV" if glv == "on" else "PORLA_NO_GLV", "1")
    if window:
        monkeypatch.setenv("PORLA_WINDOW_BITS", str(window))
    rnd = random.Random(n * 31 + window)
    g = O.glv_constants(BN)
    lam = g["lambda"]
    pts = chain(BN, min(n, 700), n)
    pts = (pts * (n // len(pts) + 1))[:n]                    # repeated points: equal operands meet in the buckets
    for i in range(0, n, 11):
        pts[i] = None
    sc = [rnd.randrange(1 << 256) for _ in range(n)]
    corners = [0, 1, BN.n - 1, BN.n, lam, BN.n - lam, lam + 1, g["a2"], -g["b1"], BN.n // 2, (1 << 253) - 1, (1 << 127) - 1, 1 << 127]
    for i, v in enumerate(corners[:n]):
        sc[i] = v
    pbytes, sbytes = b"".join(map(enc, pts)), b"".join(map(be, sc))
    got = pb.bn254_multi_exp(pbytes, sbytes, n)
    assert got == loader.bn254_msm(sbytes, pbytes, n, 4)


def test_wide_window_lookup_table_for_large_batches(monkeypatch):
    """A table shared by a large batch gets the widest windows whose look-up table fits the budget (BASELINE config 3:
    4096 commitments over 4096 bases, c = 15, 73 GB); here 300 bases under a 1 GB budget (c = 12), batches and
    single MSMs against the general path and the closed form."""
    import torch
    monkeypatch.setenv("PORLA_LUT_BUDGET_GB", "1")
    n, nb = 300, 9
    rnd = random.Random(73)
    ks = [rnd.randrange(BN.n) for _ in range(n)]
    ks[5] = 0                                               # an infinity base
    tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, b"".join(map(le, ks)), n, pb.SCALAR_LE32)
    ss = [rnd.randrange(1 << 256) for _ in range(nb * n)]
    ss[0], ss[1], ss[2] = BN.n - 1, int.from_bytes(bytes([0x08, 0x00] * 16), "big"), (1 << 256) - 1
    d_sc = torch.frombuffer(bytearray(b"".join(map(le, ss))), dtype=torch.uint8).cuda()
    general = torch.zeros(64 * nb, dtype=torch.uint8, device="cuda")
    tab.msm_device(d_sc.data_ptr(), n, general.data_ptr(), nbatch=nb, scalar_fmt=pb.SCALAR_LE32, shared_points=True)
    torch.cuda.synchronize()
    c = tab.precompute(0, n, 1 << 15)                       # 300 x 32768 scalars per launch: worth a wide table
    assert 10 <= c <= 13
    fixed = torch.zeros(64 * nb, dtype=torch.uint8, device="cuda")
    tab.msm_device(d_sc.data_ptr(), n, fixed.data_ptr(), nbatch=nb, scalar_fmt=pb.SCALAR_LE32, shared_points=True)
    torch.cuda.synchronize()
    assert torch.equal(fixed, general)
    for m in range(nb):
        want = O.mul(BN, sum(s * k for s, k in zip(ss[m * n:(m + 1) * n], ks)) % BN.n, (1, 2))
        assert bytes(fixed[64 * m:64 * m + 64].cpu().numpy().tobytes()) == enc(want), m
    assert tab.msm_resident(d_sc.data_ptr(), n, scalar_fmt=pb.SCALAR_LE32) == bytes(fixed[:64].cpu().numpy().tobytes())
    tab.destroy()


def test_more_concurrent_callers_than_staging_slots():
    """24 host threads of short-lived pools (Porla builds a ThreadPool inside align_MAC, Server.hpp:487) issue small
    commitments, audit-sized MSMs and secp256k1 generator sub-range MSMs at once: 16 lease a staging slot, the rest
    fall back to the shared staging area; every result equals the sequential one and the oracle."""
    import threading
    n = 128
    k = pb.Kzg(TAU, ALPHA)
    srs = _srs_points(k.init_srs(n), n)
    srs_bytes = b"".join(O.bn254_marshal(P) for P in srs)
    rnd = random.Random(2424)
    T = 24
    blocks = [b"".join(be(rnd.randrange(1 << 256)) for _ in range(n)) for _ in range(T)]
    pts = b"".join(map(enc, chain(BN, 200, 9)))
    scs = [b"".join(pb.bn254_scalar_set_int(rnd.randrange(1 << 31)) for _ in range(200)) for _ in range(T)]
    want_c = [loader.bn254_msm(b, srs_bytes, n, 1) for b in blocks]
    want_m = [loader.bn254_msm(s, pts, 200, 1) for s in scs]
    gens_pts = chain(SE, 64, 77)
    gens = pb.SecpGenerators(gens_pts)
    sec_sc = [[rnd.randrange(SE.n) for _ in range(16)] for _ in range(T)]
    want_s = [O.msm(SE, sc, gens_pts[16 * (t % 4):16 * (t % 4) + 16]) for t, sc in enumerate(sec_sc)]
    errors = []

    def worker(t):
        for _ in range(3):                                   # three generations of fresh threads
            def body():
                if k.compute_digest_from_srs(blocks[t]) != want_c[t]:
                    errors.append(("commit", t))
                if pb.bn254_multi_exp(pts, scs[t], 200) != want_m[t]:
                    errors.append(("msm", t))
                if gens.multi(16 * (t % 4), sec_sc[t]) != (1, want_s[t]):
                    errors.append(("secp", t))
            inner = threading.Thread(target=body)
            inner.start()
            inner.join()
    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    gens.destroy()
    assert not errors, errors[:6]


def test_device_entry_on_alternating_streams():
    """porla_msm_device is asynchronous and all calls share one scratch arena: launches issued back to back on two
    different streams must not overlap on it (the engine orders them with an event)."""
    import torch
    n = 1 << 15
    g = torch.Generator(device="cuda")
    g.manual_seed(515)
    ks = torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g)
    tab = pb.Table.multiples_of_generator(pb.CURVE_BN254, ks.data_ptr(), n, pb.SCALAR_LE32, on_device=True)
    scs = [torch.randint(-2**31, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda", generator=g) for _ in range(6)]
    ref = []
    for s in scs:
        o = torch.zeros(64, dtype=torch.uint8, device="cuda")
        tab.msm_device(s.data_ptr(), n, o.data_ptr(), scalar_fmt=pb.SCALAR_LE32)
        torch.cuda.synchronize()
        ref.append(o.clone())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    outs = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in scs]
    for i, s in enumerate(scs):                                  # no synchronisation between the calls
        tab.msm_device(s.data_ptr(), n, outs[i].data_ptr(), scalar_fmt=pb.SCALAR_LE32, stream=streams[i % 2].cuda_stream)
    torch.cuda.synchronize()
    for i in range(len(scs)):
        assert torch.equal(outs[i], ref[i]), i
    tab.destroy()


@pytest.mark.parametrize("n", [128, 16, 4])
def test_inner_product_prove_matches_oracle(n, route):
    """porla_secp256k1_inner_product_prove (Server::inner_product_prove, Server.hpp:2279-2443) byte for byte against the
    oracle's restatement: inner product, the L / R points of every round (33-byte SEC1), the re-finalized SHA-256
    transcript, the folded a and b."""
    from oracle import ipa_py
    rnd = random.Random(2279 + n)
    G = (SE.gx, SE.gy)
    gens = [O.mul(SE, rnd.randrange(1, SE.n), G) for _ in range(n)]
    u = O.mul(SE, rnd.randrange(1, SE.n), G)
    tab = pb.SecpGenerators(gens + [u])
    for trial in range(2):
        a = [rnd.randrange(1 << 256) for _ in range(n)]          # data chunks: any 256-bit value (Client.hpp:371)
        b = [rnd.randrange(SE.n) for _ in range(n)]
        if trial == 1:
            a[0], a[1], b[0] = SE.n - 1, 0, SE.n + 5
        got = tab.inner_product_prove(a, b)
        want = ipa_py.inner_product_prove(gens, u, a, b)
        assert len(got) == 32 + 66 * (n.bit_length() - 2) + 128
        assert got == want, trial
    # ... and the proof satisfies the verifier's point equation (Client.hpp:1465-1630, restated in the oracle)
    if n <= 16:
        assert ipa_py.inner_product_verify(gens, u, O.msm(SE, [x % SE.n for x in a], gens), got)
    tab.destroy()


@pytest.mark.parametrize("n", [128, 8])
def test_inner_product_verify_matches_oracle(n):
    """porla_secp256k1_inner_product_verify (Client::inner_product_verify, Client.hpp:1465-1630): accepts the prover's
    proofs, and agrees with the oracle's restatement on tampered ones (every region of the proof, malformed points)."""
    from oracle import ipa_py
    rnd = random.Random(1465 + n)
    G = (SE.gx, SE.gy)
    gens = [O.mul(SE, rnd.randrange(1, SE.n), G) for _ in range(n)]
    u = O.mul(SE, rnd.randrange(1, SE.n), G)
    tab = pb.SecpGenerators(gens + [u])
    a = [rnd.randrange(1 << 256) for _ in range(n)]
    b = [rnd.randrange(SE.n) for _ in range(n)]
    proof = tab.inner_product_prove(a, b)
    commitment = O.msm(SE, [x % SE.n for x in a], gens)
    assert tab.inner_product_verify(commitment, proof)
    assert not tab.inner_product_verify(O.add(SE, commitment, G), proof)          # another commitment
    assert not tab.inner_product_verify(None, proof)
    cases = [0, 31, 32, 33, 40, 65, 66, len(proof) - 128, len(proof) - 1]
    for pos in cases:
        bad = bytearray(proof)
        bad[pos] ^= 1 if pos != 32 else 0x06                                      # position 32: the SEC1 tag of the first L
        want = ipa_py.inner_product_verify(gens, u, commitment, bytes(bad)) if n <= 8 else False
        assert tab.inner_product_verify(commitment, bytes(bad)) == want, pos
    tab.destroy()


# ------------------------------------------------------------------ SURVEY 8(f)4, second half: the FFT on the data blocks
LCM_KZG = int("2049369031155707573937272810025244064710333118140408897690954651424664974620215782673575413484558574566298823256897068805013612518402283464943595715297281")   # utils.h:42 (ENABLE_KZG)
LCM_IPA = int("10841469693352021873483684275893008392101031472050201500515861578010683886271238884283113399568804205471204971859923723932950084770981108620251449466962241")  # utils.h:32


@pytest.mark.parametrize("lcm", [LCM_KZG, LCM_IPA])
@pytest.mark.parametrize("n_blocks,chunks,m", [(2, 3, 2), (16, 128, 4), (64, 128, 64), (8, 5, 8)])
def test_data_butterfly_stage_matches_reference_arithmetic(lcm, n_blocks, chunks, m):
    """(u + v x) % LCM and (u - v x) % LCM on 512-bit chunks (Server.hpp:1582-1588) against Python integers, both moduli of
    utils.h, with the corners of the Barrett reduction (0, LCM - 1, twiddles 0 / 1 / 2^256 - 1)."""
    import ctypes as C
    assert lcm == (207 * 2**248 + 1) * (BN.n if lcm == LCM_KZG else SE.n)          # LCM = PRIME_MODULUS * group order
    rnd = random.Random(n_blocks * 1000 + m)
    X = [[rnd.randrange(lcm) for _ in range(chunks)] for _ in range(n_blocks)]
    tw = [rnd.randrange(1 << 256) for _ in range(m // 2)]
    X[0][0], X[m // 2][0] = lcm - 1, lcm - 1
    X[1 % n_blocks][1 % chunks] = 0
    tw[0] = (1 << 256) - 1
    if m >= 8:
        tw[1], tw[2] = 0, 1
    buf = bytearray(b"".join(v.to_bytes(64, "little") for row in X for v in row))
    pb.load().porla_data_butterfly_stage(C.cast((C.c_ubyte * len(buf)).from_buffer(buf), C.c_void_p), n_blocks, chunks, m,
                                         b"".join(v.to_bytes(32, "little") for v in tw), lcm.to_bytes(64, "little"))
    m2 = m // 2
    want = [row[:] for row in X]
    for j in range(m2):
        for k in range(j, n_blocks, m):
            for p in range(chunks):
                t = tw[j] * X[k + m2][p]
                want[k][p] = (X[k][p] + t) % lcm
                want[k + m2][p] = (X[k][p] - t) % lcm
    for i in range(n_blocks):
        for p in range(chunks):
            off = 64 * (i * chunks + p)
            assert int.from_bytes(buf[off:off + 64], "little") == want[i][p], (i, p)


# ---------------------------------------------------------------------------- four lanes per point operation (quad.cuh)
@pytest.mark.parametrize("curve", [pb.CURVE_BN254, pb.CURVE_SECP256K1])
def test_quad_point_operations_match_oracle(curve):
    """quad_add / quad_dbl (one coordinate per lane, one field product per lane and level) against the oracle's affine
    group law on random points, points at infinity on either side, equal points (P + P) and opposite points (P - P)."""
    c = O.BN254 if curve == pb.CURVE_BN254 else O.SECP256K1
    G = (c.gx, c.gy)
    rnd = random.Random(77 + curve)
    n = 61
    A = [O.mul(c, rnd.randrange(1, c.n), G) for _ in range(n)]
    B = [O.mul(c, rnd.randrange(1, c.n), G) for _ in range(n)]
    for i in range(0, n, 7):
        B[i] = A[i]                                    # P + P
    for i in range(1, n, 7):
        B[i] = (A[i][0], (c.p - A[i][1]) % c.p)        # P - P
    for i in range(2, n, 7):
        A[i] = None                                    # infinity + Q
    for i in range(3, n, 7):
        B[i] = None                                    # P + infinity
    A[4] = B[4] = None
    enc = lambda P: bytes(64) if P is None else P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big")
    ta = pb.Table.from_host(curve, b"".join(map(enc, A)))
    tb = pb.Table.from_host(curve, b"".join(map(enc, B)))
    add = lambda P, Q: O.add(c, P, Q)
    neg = lambda P: None if P is None else (P[0], (c.p - P[1]) % c.p)
    want = {
        0: [add(a, b) for a, b in zip(A, B)],
        1: [add(a, a) for a in A],
        2: [add(add(a, a), add(a, b)) for a, b in zip(A, B)],
        3: [O.mul(c, 32, a) if a is not None else None for a in A],
        4: [add(add(a, b), neg(b)) for a, b in zip(A, B)],
        5: [add(add(a, b), add(a, b)) for a, b in zip(A, B)],
    }
    out = (C.c_ubyte * (64 * n))()
    bad = []
    for op, exp in want.items():
        pb.load().porla_debug_quad_op(curve, op, C.c_void_p(ta.handle), C.c_void_p(tb.handle), n, pb.POINT_BE64, out)
        got = bytes(out)
        bad += [(op, i) for i in range(n) if got[64 * i:64 * i + 64] != enc(exp[i])]
    assert not bad, bad
    ta.destroy()
    tb.destroy()
