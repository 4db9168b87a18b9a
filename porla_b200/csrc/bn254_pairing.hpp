// Host-only BN254 G2 arithmetic and optimal-ate pairing check, needed by two O(1) entry points
// of the legacy C-ABI: init_SRS (the SRS blob carries [1]G2 and [tau]G2, main.go:46-49) and
// verify_proof (kzg.Verify, main.go:187).  Out of the MSM hot path by construction: affine G2, lockstep
// multi-pairing Miller loop with one shared inversion per step, Karatsuba tower products, BN addition-chain
// hard part (the plain 761-bit exponentiation is kept as its test reference).
//
// Tower: Fp2 = Fp[i]/(i^2+1), Fp12 = Fp2[w]/(w^6 - xi), xi = 9 + i.  Twist E': y^2 = x^3 + 3/xi,
// untwist (x', y') -> (x' w^2, y' w^3).
#pragma once
#include <vector>
#include "host_bn254.hpp"

namespace porla {
namespace host {

struct Fq2 {
    Fq a0, a1;  // a0 + a1*i
    static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
    bool operator==(const Fq2& o) const { return a0 == o.a0 && a1 == o.a1; }
    Fq2 operator+(const Fq2& o) const { return Fq2{a0 + o.a0, a1 + o.a1}; }
    Fq2 operator-(const Fq2& o) const { return Fq2{a0 - o.a0, a1 - o.a1}; }
    Fq2 neg() const { return Fq2{a0.neg(), a1.neg()}; }
    Fq2 conj() const { return Fq2{a0, a1.neg()}; }
    Fq2 operator*(const Fq2& o) const {
        Fq t0 = a0 * o.a0, t1 = a1 * o.a1;
        Fq t2 = (a0 + a1) * (o.a0 + o.a1);
        return Fq2{t0 - t1, t2 - t0 - t1};
    }
    Fq2 sqr() const {   // (a0 + a1 i)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 i
        Fq t = a0 * a1;
        return Fq2{(a0 + a1) * (a0 - a1), t + t};
    }
    Fq2 mul_xi() const {   // * (9 + i)
        Fq t0 = a0.dbl().dbl().dbl() + a0, t1 = a1.dbl().dbl().dbl() + a1;   // 9 a0, 9 a1
        return Fq2{t0 - a1, t1 + a0};
    }
    Fq2 scale(const Fq& k) const { return Fq2{a0 * k, a1 * k}; }
    Fq2 inverse() const {
        Fq n = (a0.sqr() + a1.sqr()).inverse();
        return Fq2{a0 * n, (a1 * n).neg()};
    }
    Fq2 dbl() const { return *this + *this; }
};

inline Fq2 fq2_pow(const Fq2& a, const uint32_t* e, int nlimbs) {
    Fq2 r = Fq2::one();
    for (int i = nlimbs * 32 - 1; i >= 0; i--) {
        r = r.sqr();
        if ((e[i >> 5] >> (i & 31)) & 1u) r = r * a;
    }
    return r;
}

inline Fq2 fq2_xi() { return Fq2{elem_from_u64<Fq>(9), elem_from_u64<Fq>(1)}; }

// sqrt in Fq2 via the norm (p = 3 mod 4); false if a is a non-residue
inline bool fq2_sqrt(const Fq2& a, Fq2* out) {
    if (a.is_zero()) {
        *out = a;
        return true;
    }
    Fq half = elem_from_u64<Fq>(2).inverse();
    if (a.a1.is_zero()) {
        Fq s;
        if (fq_sqrt(a.a0, &s)) {
            *out = Fq2{s, Fq::zero()};
            return true;
        }
        if (fq_sqrt(a.a0.neg(), &s)) {
            *out = Fq2{Fq::zero(), s};
            return true;
        }
        return false;
    }
    Fq n;
    if (!fq_sqrt(a.a0.sqr() + a.a1.sqr(), &n)) return false;
    Fq t = (a.a0 + n) * half, x0;
    if (!fq_sqrt(t, &x0)) {
        t = (a.a0 - n) * half;
        if (!fq_sqrt(t, &x0)) return false;
    }
    Fq x1 = a.a1 * (x0.dbl()).inverse();
    Fq2 r{x0, x1};
    if (!(r.sqr() == a)) return false;
    *out = r;
    return true;
}

// ---------------------------------------------------------------------------- G2 (affine)
struct G2A {
    Fq2 x, y;
    bool inf = true;
};

inline Fq2 g2_b() {
    static const Fq2 b = Fq2{elem_from_u64<Fq>(3), Fq::zero()} * fq2_xi().inverse();
    return b;
}

inline Fq fq_from_hex_be(const char* hex) {
    uint8_t b[32];
    for (int i = 0; i < 32; i++) {
        auto nib = [](char c) -> int { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
        b[i] = (uint8_t)((nib(hex[2 * i]) << 4) | nib(hex[2 * i + 1]));
    }
    return elem_from_be<Fq>(b, 32);
}

// the generator gnark-crypto / EIP-197 use (checked on-curve by tests/test_host_abi.py)
inline G2A g2_generator() {
    G2A g;
    g.x = Fq2{fq_from_hex_be("1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"),
              fq_from_hex_be("198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2")};
    g.y = Fq2{fq_from_hex_be("12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"),
              fq_from_hex_be("090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b")};
    g.inf = false;
    return g;
}

inline G2A g2_neg(const G2A& p) {
    G2A r = p;
    if (!r.inf) r.y = r.y.neg();
    return r;
}
inline bool g2_on_curve(const G2A& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + g2_b(); }

// chord/tangent; *lambda (optional) receives the slope used (for the Miller loop)
inline G2A g2_add(const G2A& p, const G2A& q, Fq2* lambda = nullptr) {
    if (p.inf) return q;
    if (q.inf) return p;
    Fq2 lam;
    if (p.x == q.x) {
        if (!(p.y == q.y) || p.y.is_zero()) return G2A{};
        Fq2 xx = p.x.sqr();
        lam = (xx.dbl() + xx) * p.y.dbl().inverse();
    } else {
        lam = (q.y - p.y) * (q.x - p.x).inverse();
    }
    if (lambda) *lambda = lam;
    G2A r;
    r.x = lam.sqr() - p.x - q.x;
    r.y = lam * (p.x - r.x) - p.y;
    r.inf = false;
    return r;
}
inline G2A g2_mul(const G2A& p, const Fr& k_internal) {
    Fr k = k_internal.from_internal();
    G2A r;
    int top = 255;  // skip leading zeros: kzg.Verify multiplies by the 64-bit opening point (main.go:157)
    while (top >= 0 && !((k.v[top >> 5] >> (top & 31)) & 1u)) top--;
    for (int i = top; i >= 0; i--) {
        r = g2_add(r, r);
        if ((k.v[i >> 5] >> (i & 31)) & 1u) r = g2_add(r, p);
    }
    return r;
}

// gnark E2 ordering: compare A1 first, then A0
inline bool fq2_lex_largest(const Fq2& y) { return y.a1.is_zero() ? fq_lex_largest(y.a0) : fq_lex_largest(y.a1); }

// G2Affine.Bytes(): X.A1 || X.A0 big-endian, flags in the top two bits of byte 0
inline void g2_compress(const G2A& p, uint8_t* out64) {
    if (p.inf) {
        memset(out64, 0, 64);
        out64[0] = kFlagCompressedInf;
        return;
    }
    elem_to_be(p.x.a1, out64);
    elem_to_be(p.x.a0, out64 + 32);
    out64[0] |= fq2_lex_largest(p.y) ? kFlagLargest : kFlagSmallest;
}
inline bool g2_decompress(const uint8_t* b, G2A* out) {
    uint8_t flag = b[0] & 0xC0;
    if (flag == kFlagCompressedInf) {
        *out = G2A{};
        return true;
    }
    if (flag == kFlagUncompressed) return false;  // 128-byte form never appears in the SRS blob
    uint8_t xb[32];
    memcpy(xb, b, 32);
    xb[0] &= 0x3F;
    G2A p;
    p.x = Fq2{elem_from_be<Fq>(b + 32, 32), elem_from_be<Fq>(xb, 32)};
    if (!fq2_sqrt(p.x.sqr() * p.x + g2_b(), &p.y)) return false;
    if (fq2_lex_largest(p.y) != (flag == kFlagLargest)) p.y = p.y.neg();
    p.inf = false;
    *out = p;
    return true;
}

// ---------------------------------------------------------------------------- Fp12
struct Fq12 {
    Fq2 c[6];  // sum c[k] w^k, w^6 = xi
    static Fq12 one() {
        Fq12 r;
        for (int k = 0; k < 6; k++) r.c[k] = Fq2::zero();
        r.c[0] = Fq2::one();
        return r;
    }
    bool is_one() const {
        if (!(c[0] == Fq2::one())) return false;
        for (int k = 1; k < 6; k++)
            if (!c[k].is_zero()) return false;
        return true;
    }
    Fq12 operator*(const Fq12& o) const {
        Fq2 t[11];
        for (int k = 0; k < 11; k++) t[k] = Fq2::zero();
        for (int i = 0; i < 6; i++) {
            if (c[i].is_zero()) continue;
            for (int j = 0; j < 6; j++) {
                if (o.c[j].is_zero()) continue;
                t[i + j] = t[i + j] + c[i] * o.c[j];
            }
        }
        Fq12 r;
        Fq2 xi = fq2_xi();
        for (int k = 0; k < 6; k++) r.c[k] = k < 5 ? t[k] + t[k + 6] * xi : t[k];
        return r;
    }
    // Dense product / square through the tower Fp12 = Fp6[w]/(w^2 - v), Fp6 = Fp2[v]/(v^3 - xi): with
    // a = (c0, c2, c4), b = (c1, c3, c5), f = a + b w.  Karatsuba at both levels: 18 Fp2 products for a
    // product, 12 for a square, against 36 for the schoolbook operator* (kept for the sparse line values).
    struct V6 { Fq2 c[3]; };
    static V6 v6_add(const V6& x, const V6& y) { return V6{{x.c[0] + y.c[0], x.c[1] + y.c[1], x.c[2] + y.c[2]}}; }
    static V6 v6_sub(const V6& x, const V6& y) { return V6{{x.c[0] - y.c[0], x.c[1] - y.c[1], x.c[2] - y.c[2]}}; }
    static V6 v6_mul_v(const V6& x) { return V6{{x.c[2].mul_xi(), x.c[0], x.c[1]}}; }
    static V6 v6_mul(const V6& x, const V6& y) {
        Fq2 v0 = x.c[0] * y.c[0], v1 = x.c[1] * y.c[1], v2 = x.c[2] * y.c[2];
        Fq2 t0 = (x.c[1] + x.c[2]) * (y.c[1] + y.c[2]) - v1 - v2;
        Fq2 t1 = (x.c[0] + x.c[1]) * (y.c[0] + y.c[1]) - v0 - v1;
        Fq2 t2 = (x.c[0] + x.c[2]) * (y.c[0] + y.c[2]) - v0 - v2;
        return V6{{v0 + t0.mul_xi(), t1 + v2.mul_xi(), t2 + v1}};
    }
    V6 lo6() const { return V6{{c[0], c[2], c[4]}}; }
    V6 hi6() const { return V6{{c[1], c[3], c[5]}}; }
    static Fq12 from6(const V6& a, const V6& b) {
        Fq12 r;
        r.c[0] = a.c[0]; r.c[2] = a.c[1]; r.c[4] = a.c[2];
        r.c[1] = b.c[0]; r.c[3] = b.c[1]; r.c[5] = b.c[2];
        return r;
    }
    Fq12 mul_dense(const Fq12& o) const {
        V6 a1 = lo6(), b1 = hi6(), a2 = o.lo6(), b2 = o.hi6();
        V6 aa = v6_mul(a1, a2), bb = v6_mul(b1, b2);
        V6 cross = v6_sub(v6_sub(v6_mul(v6_add(a1, b1), v6_add(a2, b2)), aa), bb);
        return from6(v6_add(aa, v6_mul_v(bb)), cross);
    }
    Fq12 sqr() const {   // (a + b w)^2 = (a + b)(a + v b) - ab - v ab + 2ab w
        V6 a = lo6(), b = hi6();
        V6 ab = v6_mul(a, b);
        V6 t = v6_mul(v6_add(a, b), v6_add(a, v6_mul_v(b)));
        return from6(v6_sub(v6_sub(t, ab), v6_mul_v(ab)), v6_add(ab, ab));
    }
    // f^(p^6): w -> -w
    Fq12 conj6() const {
        Fq12 r = *this;
        r.c[1] = r.c[1].neg();
        r.c[3] = r.c[3].neg();
        r.c[5] = r.c[5].neg();
        return r;
    }
};

struct PairingConsts {
    Fq2 g1, g2, g3, g4, g5;   // xi^(k (p-1)/6), k = 1..5 (Frobenius of the twist and of Fp12)
    Fq n[6];          // xi^(k (p^2-1)/6) in Fp, k = 0..5
    PairingConsts() {
        static const uint32_t e[8] = {0x2414d4e1u, 0x34b01759u, 0xe6bda1c2u, 0xee9591c2u,
                                      0xc0403964u, 0xf40d60f3u, 0xd032f006u, 0x0810b7bdu};  // (p-1)/6
        g1 = fq2_pow(fq2_xi(), e, 8);
        g2 = g1.sqr();
        g3 = g2 * g1;
        g4 = g3 * g1;
        g5 = g4 * g1;
        Fq nrm = (g1 * g1.conj()).a0;  // xi^((p^2-1)/6)
        n[0] = Fq::one();
        for (int k = 1; k < 6; k++) n[k] = n[k - 1] * nrm;
    }
};
inline const PairingConsts& pairing_consts() {
    static const PairingConsts c;
    return c;
}

// f^(p^2): coefficients lie in Fp2 (fixed by p^2), w^k picks up xi^(k (p^2-1)/6)
inline Fq12 frob2(const Fq12& f) {
    const PairingConsts& pc = pairing_consts();
    Fq12 r;
    for (int k = 0; k < 6; k++) r.c[k] = f.c[k].scale(pc.n[k]);
    return r;
}

// inverse through the norm to the quadratic subfield is overkill here; use f^(-1) = conj-chain:
// for the easy part only f^(p^6 - 1) is needed, so invert by solving with the degree-6 tower:
// f = a + b w with a,b in Fp6 = Fp2[v], v = w^2.  f^-1 = (a - b w) / (a^2 - b^2 v).
struct Fq6 {
    Fq2 c[3];  // sum c[k] v^k, v^3 = xi
};
inline Fq6 fq6_mul(const Fq6& x, const Fq6& y) {
    Fq2 xi = fq2_xi();
    Fq2 t[5];
    for (int k = 0; k < 5; k++) t[k] = Fq2::zero();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i + j] = t[i + j] + x.c[i] * y.c[j];
    return Fq6{{t[0] + t[3] * xi, t[1] + t[4] * xi, t[2]}};
}
inline Fq6 fq6_sub(const Fq6& x, const Fq6& y) { return Fq6{{x.c[0] - y.c[0], x.c[1] - y.c[1], x.c[2] - y.c[2]}}; }
inline Fq6 fq6_mul_v(const Fq6& x) { return Fq6{{x.c[2] * fq2_xi(), x.c[0], x.c[1]}}; }
inline Fq6 fq6_inv(const Fq6& x) {
    Fq2 xi = fq2_xi();
    Fq2 A = x.c[0].sqr() - x.c[1] * x.c[2] * xi;
    Fq2 B = x.c[2].sqr() * xi - x.c[0] * x.c[1];
    Fq2 C = x.c[1].sqr() - x.c[0] * x.c[2];
    Fq2 F = (x.c[0] * A + (x.c[2] * B + x.c[1] * C) * xi).inverse();
    return Fq6{{A * F, B * F, C * F}};
}
inline Fq12 fq12_inv(const Fq12& f) {
    Fq6 a{{f.c[0], f.c[2], f.c[4]}}, b{{f.c[1], f.c[3], f.c[5]}};
    Fq6 d = fq6_inv(fq6_sub(fq6_mul(a, a), fq6_mul_v(fq6_mul(b, b))));
    Fq6 ra = fq6_mul(a, d), rb = fq6_mul(b, d);
    Fq12 r;
    r.c[0] = ra.c[0]; r.c[2] = ra.c[1]; r.c[4] = ra.c[2];
    r.c[1] = rb.c[0].neg(); r.c[3] = rb.c[1].neg(); r.c[5] = rb.c[2].neg();
    return r;
}

// line through the untwisted T with slope lam*w, evaluated at P = (xp, yp) in E(Fp):
//   l = yp - lam*xp * w + (lam*xT - yT) * w^3
inline Fq12 line_eval(const Fq2& lam, const G2A& t, const G1A& p) {
    Fq12 l;
    for (int k = 0; k < 6; k++) l.c[k] = Fq2::zero();
    l.c[0] = Fq2{p.y, Fq::zero()};
    l.c[1] = lam.scale(p.x).neg();
    l.c[3] = lam * t.x - t.y;
    return l;
}

// f^p: conjugate every Fp2 coefficient, w^k picks up xi^(k (p-1)/6)
inline Fq12 frob1(const Fq12& f) {
    const PairingConsts& pc = pairing_consts();
    Fq12 r;
    r.c[0] = f.c[0].conj();
    r.c[1] = f.c[1].conj() * pc.g1;
    r.c[2] = f.c[2].conj() * pc.g2;
    r.c[3] = f.c[3].conj() * pc.g3;
    r.c[4] = f.c[4].conj() * pc.g4;
    r.c[5] = f.c[5].conj() * pc.g5;
    return r;
}

// Miller loops of several pairs in lockstep (optimal ate, 6u + 2): one Fp12 squaring per bit for the
// whole product and ONE Fp2 inversion per step for all pairs (Montgomery's trick on the slope denominators).
struct MillerPair {
    G1A p;
    G2A q, t;
    bool live;
};
inline void miller_step(MillerPair* pr, int n, bool doubling, const G2A* addend, Fq12& f) {
    // slope denominators: 2 y_T (tangent) or x_Q - x_T (chord); a zero denominator means T = +-Q or an
    // order-2 point, impossible for points of prime order r inside the loop -- treated as "pair dead"
    Fq2 den[4], pre[4];
    Fq2 acc = Fq2::one();
    for (int i = 0; i < n; i++) {
        if (!pr[i].live) continue;
        den[i] = doubling ? pr[i].t.y.dbl() : addend[i].x - pr[i].t.x;
        if (den[i].is_zero()) {
            pr[i].live = false;
            continue;
        }
        pre[i] = acc;
        acc = acc * den[i];
    }
    Fq2 inv = acc.inverse();
    for (int i = n - 1; i >= 0; i--) {
        if (!pr[i].live) continue;
        Fq2 dinv = inv * pre[i];
        inv = inv * den[i];
        G2A& t = pr[i].t;
        Fq2 lam;
        if (doubling) {
            Fq2 xx = t.x.sqr();
            lam = (xx.dbl() + xx) * dinv;
        } else {
            lam = (addend[i].y - t.y) * dinv;
        }
        f = f * line_eval(lam, t, pr[i].p);
        G2A r;
        const Fq2& ox = doubling ? t.x : addend[i].x;
        r.x = lam.sqr() - t.x - ox;
        r.y = lam * (t.x - r.x) - t.y;
        r.inf = false;
        t = r;
    }
}

inline Fq12 miller_loop_multi(const G1A* ps, const G2A* qs, int n) {
    const PairingConsts& pc = pairing_consts();
    MillerPair pr[4];
    G2A q1[4], q2[4], qq[4];
    if (n > 4) n = 4;
    for (int i = 0; i < n; i++) {
        pr[i].p = ps[i];
        pr[i].q = qs[i];
        pr[i].t = qs[i];
        pr[i].live = !(ps[i].is_inf() || qs[i].inf);
        qq[i] = qs[i];
        // Q1 = pi(Q), Q2 = -pi^2(Q)
        q1[i].x = qs[i].x.conj() * pc.g2;
        q1[i].y = qs[i].y.conj() * pc.g3;
        q1[i].inf = false;
        q2[i].x = qs[i].x.scale(pc.n[2]);
        q2[i].y = qs[i].y.scale(pc.n[3]).neg();  // xi^((p^2-1)/2) = n[3] (= -1)
        q2[i].inf = false;
    }
    // 6u + 2 = 0x19d797039be763ba8 (65 bits)
    const uint64_t lo = 0x9d797039be763ba8ull;
    Fq12 f = Fq12::one();
    for (int i = 63; i >= 0; i--) {  // bit 64 is the leading one
        f = f.sqr();
        miller_step(pr, n, true, nullptr, f);
        if ((lo >> i) & 1ull) miller_step(pr, n, false, qq, f);
    }
    miller_step(pr, n, false, q1, f);
    miller_step(pr, n, false, q2, f);   // the last point addition itself is not needed, only its line
    return f;
}

inline Fq12 miller_loop(const G1A& p, const G2A& q) { return miller_loop_multi(&p, &q, 1); }

// ---------------------------------------------------------------------------- fixed second arguments
// kzg.Verify pairs against the two G2 points of the SRS, [1]G2 and [tau]G2, in every call.  Everything the Miller
// loop does on the G2 side (tangent / chord slopes, the Fp2 inversions, the point updates) depends on Q alone, so
// it is done once per SRS: per step the slope lam and the constant lam * x_T - y_T.  A pairing evaluation then
// costs one line evaluation (two Fp products), one sparse Fp12 product per step and pair, and the squarings.
struct G2Lines {
    std::vector<Fq2> lam, c3;
    bool valid = false;
};
inline G2Lines g2_precompute_lines(const G2A& q) {
    G2Lines out;
    if (q.inf) return out;
    const PairingConsts& pc = pairing_consts();
    G2A q1, q2, t = q;
    q1.x = q.x.conj() * pc.g2;
    q1.y = q.y.conj() * pc.g3;
    q1.inf = false;
    q2.x = q.x.scale(pc.n[2]);
    q2.y = q.y.scale(pc.n[3]).neg();
    q2.inf = false;
    auto step = [&](bool doubling, const G2A& addend) {
        Fq2 den = doubling ? t.y.dbl() : addend.x - t.x;
        if (den.is_zero()) return false;                  // impossible for points of prime order r
        Fq2 lam;
        if (doubling) {
            Fq2 xx = t.x.sqr();
            lam = (xx.dbl() + xx) * den.inverse();
        } else {
            lam = (addend.y - t.y) * den.inverse();
        }
        out.lam.push_back(lam);
        out.c3.push_back(lam * t.x - t.y);
        const Fq2& ox = doubling ? t.x : addend.x;
        G2A r;
        r.x = lam.sqr() - t.x - ox;
        r.y = lam * (t.x - r.x) - t.y;
        r.inf = false;
        t = r;
        return true;
    };
    const uint64_t lo = 0x9d797039be763ba8ull;
    bool ok = true;
    for (int i = 63; i >= 0 && ok; i--) {
        ok = step(true, q);
        if (ok && ((lo >> i) & 1ull)) ok = step(false, q);
    }
    ok = ok && step(false, q1) && step(false, q2);
    out.valid = ok;
    return out;
}

// f *= yp + b w + c w^3  (w^6 = xi): twelve Fp2 products and six scalings instead of a dense product
inline void fq12_mul_line(Fq12& f, const Fq& yp, const Fq2& b, const Fq2& c) {
    const Fq2 f5x = f.c[5].mul_xi(), f3x = f.c[3].mul_xi(), f4x = f.c[4].mul_xi();
    Fq12 r;
    r.c[0] = f.c[0].scale(yp) + b * f5x + c * f3x;
    r.c[1] = f.c[1].scale(yp) + b * f.c[0] + c * f4x;
    r.c[2] = f.c[2].scale(yp) + b * f.c[1] + c * f5x;
    r.c[3] = f.c[3].scale(yp) + b * f.c[2] + c * f.c[0];
    r.c[4] = f.c[4].scale(yp) + b * f.c[3] + c * f.c[1];
    r.c[5] = f.c[5].scale(yp) + b * f.c[4] + c * f.c[2];
    f = r;
}

// Product of the Miller loops of (ps[i], Q_i) for precomputed Q_i; pairs with an infinite first argument contribute 1.
inline Fq12 miller_loop_fixed(const G1A* ps, const G2Lines* const* lines, int n) {
    Fq2 negx[4];
    bool live[4];
    if (n > 4) n = 4;
    for (int i = 0; i < n; i++) {
        live[i] = !ps[i].is_inf() && lines[i]->valid;
        negx[i] = Fq2{ps[i].x.neg(), Fq::zero()};
    }
    const uint64_t lo = 0x9d797039be763ba8ull;
    Fq12 f = Fq12::one();
    size_t k = 0;
    auto apply = [&]() {
        for (int i = 0; i < n; i++)
            if (live[i]) fq12_mul_line(f, ps[i].y, lines[i]->lam[k].scale(negx[i].a0), lines[i]->c3[k]);
        k++;
    };
    for (int i = 63; i >= 0; i--) {
        f = f.sqr();
        apply();
        if ((lo >> i) & 1ull) apply();
    }
    apply();
    apply();
    return f;
}

// Squaring in the cyclotomic subgroup (Granger-Scott): for f with f^(p^6 + 1) = 1 -- everything after the easy part of the
// final exponentiation -- nine Fp2 squarings replace the twelve Fp2 products of the generic square.  In the tower
// Fp12 = Fp6[w]/(w^2 - v), Fp6 = Fp2[v]/(v^3 - xi) with f = (x0 + x1 v + x2 v^2) + (x3 + x4 v + x5 v^2) w
// (x0, x1, x2 = c[0], c[2], c[4]; x3, x4, x5 = c[1], c[3], c[5]):
//   f^2 = (3 (x4^2 xi + x0^2) - 2 x0,  3 (x2^2 xi + x3^2) - 2 x1,  3 (x5^2 xi + x1^2) - 2 x2,
//          6 x1 x5 xi + 2 x3,          6 x0 x4 + 2 x4,             6 x2 x3 + 2 x5)
// porla_debug_pairing_selfcheck compares it with Fq12::sqr on cyclotomic elements.
inline Fq12 fq12_cyclotomic_sqr(const Fq12& f) {
    const Fq2 &x0 = f.c[0], &x1 = f.c[2], &x2 = f.c[4], &x3 = f.c[1], &x4 = f.c[3], &x5 = f.c[5];
    Fq2 t0 = x4.sqr(), t1 = x0.sqr(), t6 = (x4 + x0).sqr() - t0 - t1;          // 2 x0 x4
    Fq2 t2 = x2.sqr(), t3 = x3.sqr(), t7 = (x2 + x3).sqr() - t2 - t3;          // 2 x2 x3
    Fq2 t4 = x5.sqr(), t5 = x1.sqr(), t8 = ((x5 + x1).sqr() - t4 - t5).mul_xi();   // 2 x1 x5 xi
    t0 = t0.mul_xi() + t1;
    t2 = t2.mul_xi() + t3;
    t4 = t4.mul_xi() + t5;
    auto three_minus_two = [](const Fq2& t, const Fq2& x) { return (t - x).dbl() + t; };   // 3 t - 2 x
    auto three_plus_two = [](const Fq2& t, const Fq2& x) { return (t + x).dbl() + t; };    // 3 t + 2 x
    Fq12 r;
    r.c[0] = three_minus_two(t0, x0);
    r.c[2] = three_minus_two(t2, x1);
    r.c[4] = three_minus_two(t4, x2);
    r.c[1] = three_plus_two(t8, x3);
    r.c[3] = three_plus_two(t6, x4);
    r.c[5] = three_plus_two(t7, x5);
    return r;
}

inline Fq12 fq12_pow_u(const Fq12& a) {   // u = 4965661367192848881 (BN254 curve parameter, 63 bits); a cyclotomic
    const uint64_t u = 4965661367192848881ull;
    Fq12 r = a;
    for (int i = 61; i >= 0; i--) {
        r = fq12_cyclotomic_sqr(r);
        if ((u >> i) & 1ull) r = r.mul_dense(a);
    }
    return r;
}

// plain square-and-multiply over (p^4 - p^2 + 1)/r: the reference routine for the addition chain below
inline Fq12 hard_part_plain(const Fq12& b) {
    static const uint32_t h[24] = {
        0xccdf42b1u, 0xe81bb482u, 0xf49c36d4u, 0x5abf5cc4u, 0x1da014fdu, 0xf1154e7eu, 0x87cdbacfu, 0xdcc7b44cu,
        0x954bcf8au, 0xaaa441e3u, 0xd5095f23u, 0x6b887d56u, 0xf3fd90c6u, 0x79581e16u, 0xd189227du, 0x3b1b1355u,
        0x61876f6bu, 0x4e529a58u, 0xd5b12278u, 0x6c0eb522u, 0x83177fafu, 0x331ec151u, 0x0b0759adu, 0x01baaa71u};
    Fq12 r = Fq12::one();
    for (int i = 760; i >= 0; i--) {
        r = r.sqr();
        if ((h[i >> 5] >> (i & 31)) & 1u) r = r.mul_dense(b);
    }
    return r;
}

// Hard part through the BN addition chain in u (three 63-bit exponentiations and a dozen products instead
// of a 761-bit exponentiation).  The value is b^(k (p^4 - p^2 + 1)/r) for a small constant k prime to r,
// which is 1 exactly when the plain hard part is 1; tests/test_host_abi.py checks accept/reject behaviour
// and the C++ self-test below compares it with hard_part_plain on the verifier's own values.
inline Fq12 hard_part_chain(const Fq12& t1) {
    Fq12 fp = frob1(t1), fp2 = frob2(t1), fp3 = frob1(fp2);
    Fq12 fu = fq12_pow_u(t1), fu2 = fq12_pow_u(fu), fu3 = fq12_pow_u(fu2);
    Fq12 y3 = frob1(fu), fu2p = frob1(fu2), fu3p = frob1(fu3), y2 = frob2(fu2);
    Fq12 y0 = fp.mul_dense(fp2).mul_dense(fp3);
    Fq12 y1 = t1.conj6(), y5 = fu2.conj6();
    y3 = y3.conj6();
    Fq12 y4 = fu.mul_dense(fu2p).conj6();
    Fq12 y6 = fu3.mul_dense(fu3p).conj6();
    Fq12 t0 = y6.sqr().mul_dense(y4).mul_dense(y5);
    Fq12 t1b = y3.mul_dense(y5).mul_dense(t0);
    t0 = t0.mul_dense(y2);
    t1b = t1b.sqr().mul_dense(t0).sqr();
    t0 = t1b.mul_dense(y1);
    t1b = t1b.mul_dense(y0);
    t0 = t0.sqr();
    return t0.mul_dense(t1b);
}

inline Fq12 final_exponentiation(const Fq12& f) {
    // easy part: f^((p^6 - 1)(p^2 + 1))
    Fq12 a = f.conj6().mul_dense(fq12_inv(f));
    Fq12 b = frob2(a).mul_dense(a);
    return hard_part_chain(b);
}

inline bool pairing_product_is_one(const G1A* ps, const G2A* qs, int n) {
    return final_exponentiation(miller_loop_multi(ps, qs, n)).is_one();
}
inline bool pairing_product_is_one_fixed(const G1A* ps, const G2Lines* const* lines, int n) {
    return final_exponentiation(miller_loop_fixed(ps, lines, n)).is_one();
}

}  // namespace host
}  // namespace porla
