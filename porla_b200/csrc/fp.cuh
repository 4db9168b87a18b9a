// 256-bit prime-field arithmetic on 8 x 32-bit limbs for sm_100a (and a portable host path).
//
// Two reduction families, selected by the parameter struct:
//   * Montgomery CIOS for BN254 Fp (p < 2^254, two spare bits => no extra carry limb), replacing
//     gnark-crypto's 4x64 Montgomery `fp.Element` that /root/reference/porla/main.go reaches
//     through bn254.G1Affine (main.go:130,136).
//   * Special-form fold for secp256k1 Fp (p = 2^256 - 2^32 - 977), replacing the 5x52 + __int128
//     code of /root/reference/porla/Utils/secp256k1_lib/field_5x52_int128_impl.h:18.
//
// Device path: inline-PTX carry chains.  Products are accumulated in two interleaved
// "even"/"odd" rows so that every 32x32->64 product lands on a 64-bit aligned slot; ptxas
// turns each mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32(.X) (checked with cuobjdump).
// Host path (no __CUDA_ARCH__): plain uint64_t arithmetic, used by the C-ABI's single-point
// operations and by the CPU-side unit tests of the formulas.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PORLA_HD __host__ __device__ __forceinline__
#define PORLA_D __device__ __forceinline__
#else
#define PORLA_HD inline
#define PORLA_D inline
#endif

namespace porla {

// ------------------------------------------------------------------------------------------
// Parameter packs.  Constants are returned through switch-free constexpr lookups so that fully
// unrolled code sees immediates.
// ------------------------------------------------------------------------------------------
struct Bn254FpParams {
    static constexpr bool kMontgomery = true;
    // p = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    PORLA_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    // R^2 mod p, R = 2^256
    PORLA_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    // R mod p (Montgomery one)
    PORLA_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static constexpr uint32_t kInv = 0xe4866389u;  // -p^{-1} mod 2^32
};

// BN254 scalar field Fr (order of G1); used on the host for KZG polynomial arithmetic
// (polynomial.Eval, kzg.Open in /root/reference/porla/main.go:81,170).
struct Bn254FrParams {
    static constexpr bool kMontgomery = true;
    PORLA_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static constexpr uint32_t kInv = 0xefffffffu;
};

struct Secp256k1FpParams {
    static constexpr bool kMontgomery = false;
    // p = 2^256 - 2^32 - 977
    PORLA_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xfffffc2fu, 0xfffffffeu, 0xffffffffu, 0xffffffffu,
                                   0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        return m[i];
    }
    PORLA_HD static constexpr uint32_t one(int i) { return i == 0 ? 1u : 0u; }
    PORLA_HD static constexpr uint32_t r2(int i) { return i == 0 ? 1u : 0u; }
    static constexpr uint32_t kInv = 0;
};

// ------------------------------------------------------------------------------------------
// Raw 256-bit helpers (shared by base fields and by scalar recoding).
// ------------------------------------------------------------------------------------------
// r = a + b, returns carry
PORLA_HD uint32_t add256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}

// r = a - b, returns borrow (1 if a < b)
PORLA_HD uint32_t sub256(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)(c & 1);
#endif
}

// ------------------------------------------------------------------------------------------
// Field element
// ------------------------------------------------------------------------------------------
// kCompact = true routes every device-side product through ONE non-inlined copy of the multiplier
// (8 + 8 registers in, 8 out).  The bucket-accumulation kernel wants everything inlined; the cold
// kernels (bucket reduction, stitching, window combine) inline a dozen XYZZ additions and become
// instruction-fetch bound (ncu: stall_no_instruction dominates) unless the code is kept small.
template <class P, bool kCompact>
struct Fp;
#ifdef __CUDACC__
template <class P>
__device__ __noinline__ Fp<P, true> fp_mul_outlined(Fp<P, true> a, Fp<P, true> b);
#endif

template <class P, bool kCompact = false>
struct alignas(16) Fp {
    uint32_t v[8];
    using Params = P;

    PORLA_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    PORLA_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r;
    }
    PORLA_HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
        return r;
    }
    PORLA_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i];
        return o == 0;
    }
    PORLA_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    PORLA_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // canonical reduce of a value known to be < 2p (carry = bit 256 of the value)
    PORLA_HD void final_sub(uint32_t carry) {
        uint32_t t[8], m[8];
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = P::mod(i);
        uint32_t borrow = sub256(t, v, m);
        // value >= p  <=>  carry || !borrow
        bool take = carry | (borrow ^ 1u);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = take ? t[i] : v[i];
    }

    PORLA_HD friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t c = add256(r.v, a.v, b.v);
        r.final_sub(c);
        return r;
    }
    PORLA_HD friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t borrow = sub256(r.v, a.v, b.v);
        uint32_t t[8], m[8];
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = P::mod(i);
        add256(t, r.v, m);
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = borrow ? t[i] : r.v[i];
        return r;
    }
    PORLA_HD Fp neg() const {
        Fp r;
        uint32_t m[8];
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = P::mod(i);
        sub256(r.v, m, v);
        bool z = is_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
        return r;
    }
    PORLA_HD Fp dbl() const { return *this + *this; }

    // -------------------------------------------------------------------------- multiplication
#ifdef __CUDA_ARCH__
    // One CIOS round of the even/odd-row Montgomery product (see mul_mont_device below).
    // Invariant on entry (round >= 1):  T = E + O[1]*2^0 + sum_{k<6} O[k+2]*2^(32(k+1)),
    // i.e. O is the row that was reduced in the previous round (O[0] == 0) viewed one limb lower.
    // Each carry chain lives inside ONE asm statement so the compiler cannot split it.
    template <bool FIRST>
    static PORLA_D void mont_round(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                uint64_t po = (uint64_t)a[j + 1] * bi, pe = (uint64_t)a[j] * bi;
                O[j] = (uint32_t)po;
                O[j + 1] = (uint32_t)(po >> 32);
                E[j] = (uint32_t)pe;
                E[j + 1] = (uint32_t)(pe >> 32);
            }
        } else {
            // E[0] += O[1]; O[j] = a_odd*bi + O[j+2] + carry   (row O moves up by 64 bits)
            asm("add.cc.u32 %0, %0, %2;\n\t"
                "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
                "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
                "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
                "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
                "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
                "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
                "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
                "madc.hi.u32 %8, %12, %13, 0;"
                : "+r"(E[0]), "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]),
                  "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
                : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(bi));
            // E += a_even*bi ; carry out lands on limb 8 = O[7]
            asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
                "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
                "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
                "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
                "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
                "addc.u32 %8, %8, 0;"
                : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]),
                  "+r"(E[6]), "+r"(E[7]), "+r"(O[7])
                : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(bi));
        }
        uint32_t mi = E[0] * P::kInv;
        // O += p_odd * mi  (cannot overflow: O <= T / 2^32 < 2^256)
        asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
            "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
            "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
            "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
            "madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]),
              "+r"(O[7])
            : "r"(P::mod(1)), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)), "r"(mi));
        // E += p_even * mi ; E[0] becomes 0 ; carry out lands on O[7]
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]),
              "+r"(E[7]), "+r"(O[7])
            : "r"(P::mod(0)), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)), "r"(mi));
    }

    // Montgomery product a*b/2^256 mod p, canonical output, for moduli below 2^254.
    static PORLA_D Fp mul_mont_device(const Fp& a, const Fp& b) {
        uint32_t even[8], odd[8];
        mont_round<true>(even, odd, a.v, b.v[0]);
        mont_round<false>(odd, even, a.v, b.v[1]);
        mont_round<false>(even, odd, a.v, b.v[2]);
        mont_round<false>(odd, even, a.v, b.v[3]);
        mont_round<false>(even, odd, a.v, b.v[4]);
        mont_round<false>(odd, even, a.v, b.v[5]);
        mont_round<false>(even, odd, a.v, b.v[6]);
        mont_round<false>(odd, even, a.v, b.v[7]);
        // T = even + (odd[1..7] one limb lower); T < 2p < 2^255
        Fp r;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, 0;"
            : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
              "=r"(r.v[6]), "=r"(r.v[7])
            : "r"(even[0]), "r"(even[1]), "r"(even[2]), "r"(even[3]), "r"(even[4]), "r"(even[5]),
              "r"(even[6]), "r"(even[7]), "r"(odd[1]), "r"(odd[2]), "r"(odd[3]), "r"(odd[4]),
              "r"(odd[5]), "r"(odd[6]), "r"(odd[7]));
        r.final_sub(0);
        return r;
    }
#endif

    // portable Montgomery CIOS (host, and reference for the device path)
    PORLA_HD static Fp mul_mont_portable(const Fp& a, const Fp& b) {
        uint32_t t[10];
#pragma unroll
        for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)a.v[j] * b.v[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[8] = (uint32_t)c;
            t[9] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * P::kInv;
            c = (uint64_t)m * P::mod(0) + t[0];
            c >>= 32;
#pragma unroll
            for (int j = 1; j < 8; j++) {
                c += (uint64_t)m * P::mod(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[7] = (uint32_t)c;
            t[8] = t[9] + (uint32_t)(c >> 32);
        }
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = t[i];
        r.final_sub(t[8]);
        return r;
    }

    // portable special-form product for p = 2^256 - C, C = 2^32 + 977 (secp256k1)
    PORLA_HD static Fp mul_special_portable(const Fp& a, const Fp& b) {
        uint32_t t[16];
#pragma unroll
        for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)a.v[j] * b.v[i] + t[i + j];
                t[i + j] = (uint32_t)c;
                c >>= 32;
            }
            t[i + 8] = (uint32_t)c;
        }
        // fold 1: r = lo + hi*977 + (hi << 32), 9 limbs + 1 spill bit
        uint32_t r[8];
        uint64_t c = 0;
        uint32_t prev = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)t[8 + i] * 977u + t[i] + prev;
            r[i] = (uint32_t)c;
            c >>= 32;
            prev = t[8 + i];
        }
        c += prev;  // value of limbs 8.. (< 2^34)
        // fold 2: add c*(2^32 + 977)
        uint64_t d = (c & 0xffffffffull) * 977u + r[0];
        r[0] = (uint32_t)d;
        d >>= 32;
        d += (c >> 32) * 977u + (c & 0xffffffffull) + r[1];
        r[1] = (uint32_t)d;
        d >>= 32;
        d += (c >> 32) + r[2];
        r[2] = (uint32_t)d;
        d >>= 32;
#pragma unroll
        for (int i = 3; i < 8; i++) {
            d += r[i];
            r[i] = (uint32_t)d;
            d >>= 32;
        }
        // fold 3: a wrap past 2^256 leaves a tiny value; add C once more (cannot wrap again)
        uint64_t e = (uint64_t)r[0] + (d ? 977u : 0u);
        r[0] = (uint32_t)e;
        e >>= 32;
        e += (uint64_t)r[1] + (d ? 1u : 0u);
        r[1] = (uint32_t)e;
        e >>= 32;
#pragma unroll
        for (int i = 2; i < 8; i++) {
            e += r[i];
            r[i] = (uint32_t)e;
            e >>= 32;
        }
        Fp out;
#pragma unroll
        for (int i = 0; i < 8; i++) out.v[i] = r[i];
        out.final_sub(0);
        return out;
    }

    PORLA_HD friend Fp operator*(const Fp& a, const Fp& b) { return mul(a, b); }
    PORLA_HD Fp sqr() const { return mul(*this, *this); }

    PORLA_HD static Fp mul_inlined(const Fp& a, const Fp& b) {
        if (P::kMontgomery) {
#ifdef __CUDA_ARCH__
            return mul_mont_device(a, b);
#else
            return mul_mont_portable(a, b);
#endif
        } else {
            return mul_special_portable(a, b);
        }
    }
    PORLA_HD static Fp mul(const Fp& a, const Fp& b) {
#ifdef __CUDA_ARCH__
        if constexpr (kCompact) return fp_mul_outlined<P>(a, b);
#endif
        return mul_inlined(a, b);
    }

    // to/from the internal representation (Montgomery for BN254, identity for secp256k1)
    PORLA_HD Fp to_internal() const {
        if (!P::kMontgomery) return *this;
        Fp r2;
#pragma unroll
        for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
        return mul(*this, r2);
    }
    PORLA_HD Fp from_internal() const {
        if (!P::kMontgomery) return *this;
        Fp o = zero();
        o.v[0] = 1;
        return mul(*this, o);
    }

    // a^(p-2); variable time, simple square-and-multiply (only on cold paths)
    PORLA_HD Fp inverse() const {
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = P::mod(i);
        e[0] -= 2;  // both moduli end in ...47 / ...2f: no borrow
        Fp r = one();
        for (int i = 255; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1u) r = mul(r, *this);
        }
        return r;
    }
};

#ifdef __CUDACC__
template <class P>
__device__ __noinline__ Fp<P, true> fp_mul_outlined(Fp<P, true> a, Fp<P, true> b) {
    return Fp<P, true>::mul_inlined(a, b);
}
#endif

}  // namespace porla
