// Host-side BN254 helpers for the C-ABI: byte codecs (gnark-crypto v0.6.0 layouts, SURVEY.md
// Appendix B), single-point group operations and Fr polynomial arithmetic.  These back the O(1)
// entry points of /root/reference/porla/main.go (add_point :196, mult_point :205, neg_point :217,
// compute_digest :71, kzg.Open inside create_proof :170); the MSMs never come through here.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "ec.cuh"
#include "host_fp64.hpp"

namespace porla {
namespace host {

// Base field on 4 x 64-bit limbs (unsigned __int128 products): every host-side group operation
// (mult_point alone is called O(n log n) times per rebuild, Server.hpp:1548-1687) and the pairing run
// 3-4x faster than on the portable 8 x 32 code.  Same in-memory layout as the device type.
using Fq = Fp64<Bn254Fq64Params>;
using Fq32 = Fp<Bn254FpParams>;
using Fr = Fp<Bn254FrParams>;
// the 8 x 32-limb type with the same modulus and memory layout (used for byte <-> limb conversion)
template <class F> struct Limb32Of { using type = F; };
template <> struct Limb32Of<Fq> { using type = Fq32; };
using G1A = Affine<Fq>;
using G1X = XYZZ<Fq>;

// ---- raw 256-bit <-> big-endian bytes
inline void be32_to_limbs(const uint8_t* b, uint32_t* l) {
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = b + 4 * (7 - i);
        l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
}
inline void limbs_to_be32(const uint32_t* l, uint8_t* b) {
    for (int i = 0; i < 8; i++) {
        uint8_t* p = b + 4 * (7 - i);
        p[0] = (uint8_t)(l[i] >> 24);
        p[1] = (uint8_t)(l[i] >> 16);
        p[2] = (uint8_t)(l[i] >> 8);
        p[3] = (uint8_t)l[i];
    }
}

// canonical (non-Montgomery) limbs of a value < 2^256, reduced below the modulus
template <class F>
inline void reduce_canonical(uint32_t* l) {
    uint32_t m[8], t[8];
    for (int i = 0; i < 8; i++) m[i] = F::Params::mod(i);
    for (int k = 0; k < 8; k++) {
        if (sub256(t, l, m)) break;
        memcpy(l, t, 32);
    }
}

// fp/fr Element.SetBytes: big-endian integer of ANY length reduced mod the modulus; returns the
// element in internal (Montgomery) form.
template <class F>
inline F elem_from_be(const uint8_t* b, size_t len) {
    if (len <= 32) {
        uint8_t buf[32] = {0};
        memcpy(buf + (32 - len), b, len);
        using L = typename Limb32Of<F>::type;
        L x;
        be32_to_limbs(buf, x.v);
        reduce_canonical<L>(x.v);
        F r;
        memcpy(&r, &x, 32);
        return r.to_internal();
    }
    // long inputs: Horner in base 2^8
    F acc = F::zero();
    F c256 = F::zero();
    c256.v[0] = 256;
    c256 = c256.to_internal();
    for (size_t i = 0; i < len; i++) {
        F d = F::zero();
        d.v[0] = b[i];
        acc = acc * c256 + d.to_internal();
    }
    return acc;
}
template <class F>
inline void elem_to_be(const F& x, uint8_t* out32) {
    F c = x.from_internal();
    uint32_t l[8];
    memcpy(l, &c, 32);
    limbs_to_be32(l, out32);
}
template <class F>
inline F elem_from_u64(uint64_t v) {
    F x = F::zero();
    memcpy(&x, &v, 8);   // little-endian limbs, 32- or 64-bit
    return x.to_internal();
}

template <class F>
inline F pow_limbs(const F& a, const uint32_t* e) {
    F r = F::one();
    for (int i = 255; i >= 0; i--) {
        r = r.sqr();
        if ((e[i >> 5] >> (i & 31)) & 1u) r = r * a;
    }
    return r;
}

// sqrt in Fq (p = 3 mod 4): a^((p+1)/4); returns false for non-residues
inline bool fq_sqrt(const Fq& a, Fq* out) {
    uint32_t e[8], one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, m[8];
    for (int i = 0; i < 8; i++) m[i] = Bn254FpParams::mod(i);
    add256(e, m, one);
    for (int i = 0; i < 8; i++) e[i] = (e[i] >> 2) | (i < 7 ? (e[i + 1] << 30) : 0u);
    Fq y = pow_limbs(a, e);
    if (y.sqr() != a) return false;
    *out = y;
    return true;
}

// lexicographic "largest" test of gnark: y > (p-1)/2 on the canonical integer
inline bool fq_lex_largest(const Fq& y_internal) {
    Fq yc = y_internal.from_internal();
    Fq32 y;
    memcpy(&y, &yc, 32);
    uint32_t half[8], m[8], t[8];
    for (int i = 0; i < 8; i++) m[i] = Bn254FpParams::mod(i);
    for (int i = 0; i < 8; i++) half[i] = (m[i] >> 1) | (i < 7 ? (m[i + 1] << 31) : 0u);  // (p-1)/2
    return sub256(t, half, y.v) != 0;  // half < y
}

enum : uint8_t { kFlagUncompressed = 0x00, kFlagCompressedInf = 0x40, kFlagSmallest = 0x80, kFlagLargest = 0xC0 };

inline Fq curve_b() { return elem_from_u64<Fq>(3); }

// G1Affine.SetBytes on a 64-byte (uncompressed) or 32-byte (compressed) buffer.
inline bool g1_unmarshal(const uint8_t* b, size_t len, G1A* out) {
    if (len < 32) return false;
    uint8_t flag = b[0] & 0xC0;
    if (flag == kFlagUncompressed) {
        if (len < 64) return false;
        uint8_t xb[32];
        memcpy(xb, b, 32);
        out->x = elem_from_be<Fq>(xb, 32);
        out->y = elem_from_be<Fq>(b + 32, 32);
        return true;
    }
    if (flag == kFlagCompressedInf) {
        *out = G1A::inf();
        return true;
    }
    uint8_t xb[32];
    memcpy(xb, b, 32);
    xb[0] &= 0x3F;
    Fq x = elem_from_be<Fq>(xb, 32);
    Fq y;
    if (!fq_sqrt(x.sqr() * x + curve_b(), &y)) return false;
    if (fq_lex_largest(y) != (flag == kFlagLargest)) y = y.neg();
    out->x = x;
    out->y = y;
    return true;
}
inline void g1_marshal(const G1A& p, uint8_t* out64) {
    elem_to_be(p.x, out64);
    elem_to_be(p.y, out64 + 32);
}
inline void g1_compress(const G1A& p, uint8_t* out32) {
    if (p.is_inf()) {
        memset(out32, 0, 32);
        out32[0] = kFlagCompressedInf;
        return;
    }
    elem_to_be(p.x, out32);
    out32[0] |= fq_lex_largest(p.y) ? kFlagLargest : kFlagSmallest;
}

inline G1A g1_generator() { return G1A{elem_from_u64<Fq>(1), elem_from_u64<Fq>(2)}; }

inline G1A g1_add(const G1A& a, const G1A& b) {
    G1X r = G1X::from_affine(a);
    r.madd(b);
    return r.to_affine();
}
// out = a * b mod 2^(32*no) on 32-bit limbs (host twin of mul_limbs in msm_kernels.cuh)
inline void mul_limbs_host(const uint32_t* a, int na, const uint32_t* b, int nb, uint32_t* out, int no) {
    for (int i = 0; i < no; i++) out[i] = 0;
    for (int i = 0; i < na; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < nb && i + j < no; j++) {
            uint64_t v = (uint64_t)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (uint32_t)v;
            carry = v >> 32;
        }
        if (i + nb < no) out[i + nb] = (uint32_t)carry;
    }
}
// k (canonical, < r) -> |k1|, |k2| < 2^127 and their signs with k = k1 + k2 * lambda mod r: the device routine
// glv_split (msm_kernels.cuh) on the host, same constants (Bn254::glv_*, pinned in tests/test_oracle.py).
inline void glv_split_host(const uint32_t* k, uint32_t* k1, uint32_t* k2, bool* neg1, bool* neg2) {
    uint32_t g1[3], g2[5], a1[2], a2[4], nb1[4];
    for (int i = 0; i < 3; i++) g1[i] = Bn254::glv_g1(i);
    for (int i = 0; i < 5; i++) g2[i] = Bn254::glv_g2(i);
    for (int i = 0; i < 2; i++) a1[i] = Bn254::glv_a1(i);
    for (int i = 0; i < 4; i++) {
        a2[i] = Bn254::glv_a2(i);
        nb1[i] = Bn254::glv_nb1(i);
    }
    uint32_t t1[11], t2[13], p1[6], p2[6], q1[6], q2[6], v1[6], v2[6];
    mul_limbs_host(k, 8, g1, 3, t1, 11);
    mul_limbs_host(k, 8, g2, 5, t2, 13);
    const uint32_t *c1 = t1 + 8, *c2 = t2 + 8;
    mul_limbs_host(c1, 2, a1, 2, p1, 6);
    mul_limbs_host(c2, 4, a2, 4, p2, 6);
    mul_limbs_host(c1, 2, nb1, 4, q1, 6);
    mul_limbs_host(c2, 4, a1, 2, q2, 6);
    uint32_t b1 = 0, b2 = 0, b3 = 0;
    for (int i = 0; i < 6; i++) {
        uint64_t d = (uint64_t)k[i] - p1[i] - b1;
        b1 = (uint32_t)(d >> 63);
        uint64_t e = (uint64_t)(uint32_t)d - p2[i] - b2;
        b2 = (uint32_t)(e >> 63);
        v1[i] = (uint32_t)e;
        uint64_t f = (uint64_t)q1[i] - q2[i] - b3;
        b3 = (uint32_t)(f >> 63);
        v2[i] = (uint32_t)f;
    }
    *neg1 = (v1[5] >> 31) != 0;
    *neg2 = (v2[5] >> 31) != 0;
    uint32_t cy1 = *neg1 ? 1u : 0u, cy2 = *neg2 ? 1u : 0u;
    for (int i = 0; i < 6; i++) {
        uint64_t x = (uint64_t)(*neg1 ? ~v1[i] : v1[i]) + cy1;
        v1[i] = (uint32_t)x;
        cy1 = (uint32_t)(x >> 32);
        uint64_t y = (uint64_t)(*neg2 ? ~v2[i] : v2[i]) + cy2;
        v2[i] = (uint32_t)y;
        cy2 = (uint32_t)(y >> 32);
    }
    for (int i = 0; i < 4; i++) {
        k1[i] = v1[i];
        k2[i] = v2[i];
    }
}

// k * p, k an Fr element (internal form).  GLV: k = k1 + k2 lambda, phi(p) = (beta x, y) = lambda p, then joint
// 2 + 2-bit windows over (k1, k2): 15 precomputed combinations i p + j phi(p), 64 steps of two doublings and at
// most one addition -- 128 doublings instead of 252 (mult_point is called O(n log n) times per rebuild by an
// unchanged Porla, Server.hpp:1548-1687).  Short scalars (31-bit audit coefficients, utils.h:271-275) skip the
// leading zero windows.
inline G1A g1_mul(const G1A& p, const Fr& k_internal) {
    Fr k = k_internal.from_internal();
    if (p.is_inf() || k.is_zero()) return G1A::inf();
    uint32_t k1[4], k2[4];
    bool n1, n2;
    glv_split_host(k.v, k1, k2, &n1, &n2);
    G1A p1 = p, p2 = p;
    Fq beta;
    for (int i = 0; i < 4; i++) beta.v[i] = (uint64_t)Bn254::glv_beta_mont(2 * i) | ((uint64_t)Bn254::glv_beta_mont(2 * i + 1) << 32);
    p2.x = p.x * beta;
    if (n1) p1.y = p1.y.neg();
    if (n2) p2.y = p2.y.neg();
    // tab[4 j + i] = i p1 + j p2
    G1X tab[16];
    tab[0] = G1X::inf();
    for (int i = 1; i < 4; i++) {
        tab[i] = tab[i - 1];
        tab[i].madd(p1);
    }
    for (int j = 1; j < 4; j++)
        for (int i = 0; i < 4; i++) {
            tab[4 * j + i] = tab[4 * (j - 1) + i];
            tab[4 * j + i].madd(p2);
        }
    auto digit = [&](int w) {   // 2 bits of k1 and of k2 at bit 2w
        return ((k1[w >> 4] >> (2 * (w & 15))) & 3u) | (((k2[w >> 4] >> (2 * (w & 15))) & 3u) << 2);
    };
    int top = 63;
    while (top > 0 && digit(top) == 0) top--;
    G1X r = tab[digit(top)];
    for (int w = top - 1; w >= 0; w--) {
        r = r.dbl().dbl();
        uint32_t d = digit(w);
        if (d) r.add(tab[d]);
    }
    return r.to_affine();
}

// polynomial.Eval: Horner
inline Fr poly_eval(const std::vector<Fr>& f, const Fr& z) {
    Fr acc = Fr::zero();
    for (size_t i = f.size(); i-- > 0;) acc = acc * z + f[i];
    return acc;
}
// kzg.Open's quotient (f - f(z)) / (X - z): synthetic division, degree n-2
inline std::vector<Fr> poly_quotient(const std::vector<Fr>& f, const Fr& z) {
    std::vector<Fr> h(f.size() > 0 ? f.size() - 1 : 0);
    Fr carry = Fr::zero();
    for (size_t i = f.size(); i-- > 1;) {
        carry = f[i] + carry * z;
        h[i - 1] = carry;
    }
    return h;
}

}  // namespace host
}  // namespace porla
