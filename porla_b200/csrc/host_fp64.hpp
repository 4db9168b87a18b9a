// Host-only 4 x 64-bit field arithmetic with the same interface as Fp<> (fp.cuh), so XYZZ<> /
// Affine<> (ec.cuh) instantiate over it unchanged.  The in-memory representation is identical to
// the device's 8 x 32-bit little-endian limbs (same Montgomery radix 2^256 for BN254), so device
// buffers can be reinterpreted directly.  Used for the serial tail of an MSM -- the Horner
// combination of the per-window sums (c doublings per window, ~254 in total) and the final affine
// normalisation -- which is latency-bound O(1) work that a single GPU thread runs ~15x slower than
// a host core.  The data-parallel stages (recoding, sorting, bucket accumulation, bucket
// reduction) never come through here.
#pragma once
#include <stdint.h>
#include <string.h>

namespace porla {
namespace host {

typedef unsigned __int128 u128;

struct Bn254Fq64Params {
    static constexpr bool kMontgomery = true;
    static constexpr uint64_t mod(int i) {
        constexpr uint64_t m[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
        return m[i];
    }
    static constexpr uint64_t one(int i) {
        constexpr uint64_t m[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
        return m[i];
    }
    static constexpr uint64_t r2(int i) {
        constexpr uint64_t m[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
        return m[i];
    }
    static constexpr uint64_t kInv = 0x87d20782e4866389ull;
};

struct SecpFq64Params {
    static constexpr bool kMontgomery = false;
    static constexpr uint64_t mod(int i) {
        constexpr uint64_t m[4] = {0xfffffffefffffc2full, 0xffffffffffffffffull, 0xffffffffffffffffull, 0xffffffffffffffffull};
        return m[i];
    }
    static constexpr uint64_t one(int i) { return i == 0 ? 1 : 0; }
    static constexpr uint64_t r2(int i) { return i == 0 ? 1 : 0; }
    static constexpr uint64_t kInv = 0;
};

template <class P>
struct alignas(16) Fp64 {
    uint64_t v[4];
    using Params = P;

    static Fp64 zero() { return Fp64{{0, 0, 0, 0}}; }
    static Fp64 one() { return Fp64{{P::one(0), P::one(1), P::one(2), P::one(3)}}; }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const Fp64& b) const { return ((v[0] ^ b.v[0]) | (v[1] ^ b.v[1]) | (v[2] ^ b.v[2]) | (v[3] ^ b.v[3])) == 0; }
    bool operator!=(const Fp64& b) const { return !(*this == b); }

    static uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)a[i] + b[i];
            r[i] = (uint64_t)c;
            c >>= 64;
        }
        return (uint64_t)c;
    }
    static uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 t = (u128)a[i] - b[i] - borrow;
            r[i] = (uint64_t)t;
            borrow = (uint64_t)(t >> 64) & 1;
        }
        return borrow;
    }
    void final_sub(uint64_t carry) {
        uint64_t m[4] = {P::mod(0), P::mod(1), P::mod(2), P::mod(3)}, t[4];
        uint64_t borrow = sub4(t, v, m);
        if (carry || !borrow) memcpy(v, t, 32);
    }
    friend Fp64 operator+(const Fp64& a, const Fp64& b) {
        Fp64 r;
        uint64_t c = add4(r.v, a.v, b.v);
        r.final_sub(c);
        return r;
    }
    friend Fp64 operator-(const Fp64& a, const Fp64& b) {
        Fp64 r;
        if (sub4(r.v, a.v, b.v)) {
            uint64_t m[4] = {P::mod(0), P::mod(1), P::mod(2), P::mod(3)};
            add4(r.v, r.v, m);
        }
        return r;
    }
    Fp64 neg() const {
        if (is_zero()) return *this;
        Fp64 r;
        uint64_t m[4] = {P::mod(0), P::mod(1), P::mod(2), P::mod(3)};
        sub4(r.v, m, v);
        return r;
    }
    Fp64 dbl() const { return *this + *this; }

    static Fp64 mul_mont(const Fp64& a, const Fp64& b) {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (u128)a.v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P::kInv;
            c = (u128)m * P::mod(0) + t[0];
            c >>= 64;
            for (int j = 1; j < 4; j++) {
                c += (u128)m * P::mod(j) + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp64 r;
        memcpy(r.v, t, 32);
        r.final_sub(t[4]);
        return r;
    }
    // p = 2^256 - C, C = 2^32 + 977
    static Fp64 mul_special(const Fp64& a, const Fp64& b) {
        const uint64_t C = 0x1000003d1ull;
        uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (u128)a.v[j] * b.v[i] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
            t[i + 4] = (uint64_t)c;
        }
        uint64_t r[4];
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[4 + i] * C + t[i];
            r[i] = (uint64_t)c;
            c >>= 64;
        }
        // c < 2^34: fold again
        u128 d = (u128)(uint64_t)c * C + r[0];
        r[0] = (uint64_t)d;
        d >>= 64;
        for (int i = 1; i < 4; i++) {
            d += r[i];
            r[i] = (uint64_t)d;
            d >>= 64;
        }
        if ((uint64_t)d) {  // wrapped past 2^256: the value is tiny, add C once more
            u128 e = (u128)r[0] + C;
            r[0] = (uint64_t)e;
            e >>= 64;
            for (int i = 1; i < 4; i++) {
                e += r[i];
                r[i] = (uint64_t)e;
                e >>= 64;
            }
        }
        Fp64 out;
        memcpy(out.v, r, 32);
        out.final_sub(0);
        return out;
    }
    static Fp64 mul(const Fp64& a, const Fp64& b) { return P::kMontgomery ? mul_mont(a, b) : mul_special(a, b); }
    friend Fp64 operator*(const Fp64& a, const Fp64& b) { return mul(a, b); }
    Fp64 sqr() const { return mul(*this, *this); }
    static Fp64 mul2add(const Fp64& a, const Fp64& b, const Fp64& c, const Fp64& d) { return mul(a, b) + mul(c, d); }

    Fp64 to_internal() const {
        if (!P::kMontgomery) return *this;
        Fp64 r2{{P::r2(0), P::r2(1), P::r2(2), P::r2(3)}};
        return mul(*this, r2);
    }
    Fp64 from_internal() const {
        if (!P::kMontgomery) return *this;
        Fp64 o{{1, 0, 0, 0}};
        return mul(*this, o);
    }
    // a^(p-2) by square-and-multiply: the reference routine the fast inverse below is tested against
    Fp64 inverse_fermat() const {
        uint64_t e[4] = {P::mod(0) - 2, P::mod(1), P::mod(2), P::mod(3)};
        Fp64 r = one();
        for (int i = 255; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, *this);
        }
        return r;
    }

    // helpers of the binary extended Euclid below (plain 256-bit integers, bit 256 in `hi`)
    static bool geq4(const uint64_t* a, const uint64_t* b) {
        for (int i = 3; i >= 0; i--)
            if (a[i] != b[i]) return a[i] > b[i];
        return true;
    }
    static void shr1(uint64_t* a, uint64_t hi) {
        a[0] = (a[0] >> 1) | (a[1] << 63);
        a[1] = (a[1] >> 1) | (a[2] << 63);
        a[2] = (a[2] >> 1) | (a[3] << 63);
        a[3] = (a[3] >> 1) | (hi << 63);
    }
    // x <- x / 2 mod p  (p odd)
    static void half_mod(uint64_t* x, const uint64_t* m) {
        uint64_t hi = 0;
        if (x[0] & 1) hi = add4(x, x, m);
        shr1(x, hi);
    }
    // x <- x - y mod p, both below p
    static void sub_mod(uint64_t* x, const uint64_t* y, const uint64_t* m) {
        if (sub4(x, x, y)) add4(x, x, m);
    }

    // Modular inverse by the binary extended Euclidean algorithm (about 2 x 256 shift/subtract steps on
    // four limbs: ~10x faster than the Fermat exponentiation; variable time, which is fine here: nothing
    // secret flows through the MSM engine or the verifier).  Works on the stored representation: for a
    // Montgomery element aR it computes (aR)^-1 and multiplies by R^2 twice to return a^-1 R.
    // The inverse of zero is zero, like the Fermat routine.
    Fp64 inverse() const {
        if (is_zero()) return zero();
        const uint64_t m[4] = {P::mod(0), P::mod(1), P::mod(2), P::mod(3)};
        uint64_t u[4] = {v[0], v[1], v[2], v[3]}, w[4] = {m[0], m[1], m[2], m[3]};
        uint64_t x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
        auto is_one = [](const uint64_t* a) { return a[0] == 1 && (a[1] | a[2] | a[3]) == 0; };
        while (!is_one(u) && !is_one(w)) {
            while (!(u[0] & 1)) {
                shr1(u, 0);
                half_mod(x1, m);
            }
            while (!(w[0] & 1)) {
                shr1(w, 0);
                half_mod(x2, m);
            }
            if (geq4(u, w)) {
                sub4(u, u, w);
                sub_mod(x1, x2, m);
            } else {
                sub4(w, w, u);
                sub_mod(x2, x1, m);
            }
        }
        Fp64 r;
        memcpy(r.v, is_one(u) ? x1 : x2, 32);
        if (!P::kMontgomery) return r;
        Fp64 r2{{P::r2(0), P::r2(1), P::r2(2), P::r2(3)}};
        return mul(mul(r, r2), r2);
    }
};

}  // namespace host
}  // namespace porla
