// Short-Weierstrass group law for a = 0 curves (BN254 G1: y^2 = x^3 + 3, secp256k1: y^2 = x^3 + 7)
// in extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2).
//
// This is the coordinate system gnark-crypto's MultiExp keeps its buckets in (SURVEY.md
// Appendix B) and replaces secp256k1's Jacobian gej_add_ge_var
// (/root/reference/porla/Utils/secp256k1_lib/group_impl.h:389) as the bucket update.
// All routines handle the exceptional cases (infinity, P + P, P - P) exactly, because Porla's
// inputs contain them (alignment MACs start as infinity, Server.hpp:1534-1535) and parity is
// judged on canonical affine bytes.
#pragma once
#include "fp.cuh"

namespace porla {

template <class F>
struct alignas(16) Affine {
    F x, y;  // infinity is encoded as x = y = 0 (not on either curve)
    PORLA_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    PORLA_HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
    PORLA_HD Affine neg() const { return Affine{x, y.neg()}; }
};

template <class F>
struct alignas(16) XYZZ {
    F x, y, zz, zzz;

    PORLA_HD bool is_inf() const { return zz.is_zero(); }
    PORLA_HD static XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    PORLA_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
    PORLA_HD XYZZ neg() const { return XYZZ{x, y.neg(), zz, zzz}; }

    // 2 * (affine p), p finite.  mdbl-2008-s-1 with a = 0.
    PORLA_HD static XYZZ dbl_affine(const Affine<F>& p) {
        if (p.y.is_zero()) return inf();  // order-2 point: cannot occur on prime-order curves
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * p.y;
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // 2 * this.  dbl-2008-s-1 with a = 0.
    PORLA_HD XYZZ dbl() const {
        if (is_inf()) return *this;
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += affine q.  madd-2008-s: 8M + 2S on the generic path.
    PORLA_HD void madd(const Affine<F>& q) {
        if (q.is_inf()) return;
        if (is_inf()) {
            *this = XYZZ{q.x, q.y, F::one(), F::one()};
            return;
        }
        madd_finite(q);
    }

    // this += affine q where neither operand is infinity (the bucket-accumulation hot op).
    PORLA_HD void madd_finite(const Affine<F>& q) {
        F p = q.x * zz - x;
        F r = q.y * zzz - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(q);
            else *this = inf();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F qq = x * pp;
        F x3 = r.sqr() - ppp - qq.dbl();
        y = F::mul2add(r, qq - x3, y.neg(), ppp);   // r (qq - x3) - y ppp with one reduction
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o.  add-2008-s: 12M + 2S on the generic path.
    PORLA_HD void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        F u1 = x * o.zz;
        F u2 = o.x * zz;
        F s1 = y * o.zzz;
        F s2 = o.y * zzz;
        F p = u2 - u1;
        F r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = inf();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F qq = u1 * pp;
        F x3 = r.sqr() - ppp - qq.dbl();
        y = r * (qq - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }

    // canonical affine form: one field inversion.  1/ZZ = (ZZ / ZZZ)^2 because ZZ^3 = ZZZ^2.
    PORLA_HD Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        F i3 = zzz.inverse();
        F t = zz * i3;
        Affine<F> a;
        a.x = x * t.sqr();
        a.y = y * i3;
        return a;
    }
};

// k * p for a small non-negative k (used by the bucket reduction to weight a chunk sum)
template <class F>
PORLA_HD XYZZ<F> mul_small(const XYZZ<F>& p, uint32_t k) {
    XYZZ<F> r = XYZZ<F>::inf();
    if (k == 0 || p.is_inf()) return r;
    int top = 31;
    while (!((k >> top) & 1u)) top--;
    for (int i = top; i >= 0; i--) {
        r = r.dbl();
        if ((k >> i) & 1u) r.add(p);
    }
    return r;
}

// Curve tags ------------------------------------------------------------------------------
using Bn254Fp = Fp<Bn254FpParams>;
using SecpFp = Fp<Secp256k1FpParams>;
using Bn254FpCompact = Fp<Bn254FpParams, true>;
using SecpFpCompact = Fp<Secp256k1FpParams, true>;

struct Bn254 {
    using F = Bn254Fp;
    using FC = Bn254FpCompact;  // same layout, non-inlined multiplier (cold kernels)
    static constexpr int kFieldBits = 254;      // bit length of the base-field modulus
    static constexpr int kScalarBits = 254;     // bits the window recoder covers (r < 2^254)
    static constexpr bool kHalveScalar = false;
    // GLV: phi(x, y) = (beta x, y) = lambda (x, y) with beta^3 = 1 in Fp, lambda^3 = 1 mod r.  A scalar k splits as
    // k = k1 + k2 lambda with |k1|, |k2| < 2^127 (lattice basis (a1, b1), (a2, b2), a1 b2 - a2 b1 = r; constants
    // re-derived independently and compared limb by limb in tests/test_oracle.py), so an n-term MSM becomes a
    // 2n-term MSM over (P_i, phi(P_i)) with half as many windows, i.e. half as many buckets to reduce.
    static constexpr bool kGlv = true;
    static constexpr int kGlvBits = 127;
    // bucket count up to which the bucket reduction runs on quads (k_reduce_scan): measured crossover against k_reduce
    // (0.93 against 0.92 ms at 983 k buckets, 0.66 against 0.71 ms for one set of 524 k)
    static constexpr uint32_t kReduceQuadMax = 600000;
    PORLA_HD static constexpr uint32_t glv_beta_mont(int i) {   // beta * 2^256 mod p
        constexpr uint32_t m[8] = {0xd782e155u, 0x71930c11u, 0xffbe3323u, 0xa6bb947cu,
                                   0xd4741444u, 0xaa303344u, 0x26594943u, 0x2c3b3f0du};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t glv_g1(int i) {   // floor(2^256 * b2 / r)
        constexpr uint32_t m[3] = {0xc7e0b3d7u, 0xd91d232eu, 0x2u};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t glv_g2(int i) {   // floor(2^256 * (-b1) / r)
        constexpr uint32_t m[5] = {0x391eb18du, 0x7a7bd9d4u, 0xa773d2cfu, 0x4ccef014u, 0x2u};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t glv_a1(int i) {   // a1 = b2
        constexpr uint32_t m[2] = {0x94d213e3u, 0x89d32568u};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t glv_a2(int i) {
        constexpr uint32_t m[4] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t glv_nb1(int i) {  // -b1
        constexpr uint32_t m[4] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u};
        return m[i];
    }
    // r (scalar field order), little-endian 32-bit limbs
    PORLA_HD static constexpr uint32_t order(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static constexpr uint32_t kB = 3;
};

struct Secp256k1 {
    using F = SecpFp;
    using FC = SecpFpCompact;
    // n is just below 2^256: scalars above n/2 are recoded as -(n - s), so 255 magnitude bits (plus the
    // carry of the signed digits) span exactly 16 windows of 16 bits instead of 17 with a 1-bit top window.
    static constexpr int kFieldBits = 256;
    static constexpr int kScalarBits = 255;
    static constexpr bool kHalveScalar = true;
    // no GLV here: the split halves of a secp256k1 scalar reach 2^128 (they need 129 bits with the signed digits'
    // carry), so 2n terms x 9 windows of 16 bits would cost more bucket updates than n terms x 16 windows
    static constexpr bool kGlv = false;
    static constexpr int kGlvBits = 0;
    // quads up to 300 k buckets (the cheaper special-form product moves the crossover: at 524 k buckets, 2^18 / 2^20 terms,
    // k_reduce 1.231 / 2.880 ms per MSM against 1.254 / 2.904 ms on quads; 2^16: quads 0.63 against 0.70 ms)
    static constexpr uint32_t kReduceQuadMax = 300000;
    PORLA_HD static constexpr uint32_t order(int i) {
        constexpr uint32_t m[8] = {0xd0364141u, 0xbfd25e8cu, 0xaf48a03bu, 0xbaaedce6u,
                                   0xfffffffeu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        return m[i];
    }
    static constexpr uint32_t kB = 7;
};

}  // namespace porla
