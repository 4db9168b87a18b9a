// secp256k1 (Porla IPA mode) boundary: a drop-in for the static secp256k1_ecmult_multi_var of
// /root/reference/porla/Utils/secp256k1_lib/ecmult_impl.h:814-860 operating on the reference's
// own struct layouts (field_5x52.h:12-21, group.h:13-28, scalar_4x64.h:13-15), plus the SEC1
// serialiser (eckey_impl.h:36-52) used as the parity format.  Conversion rules: SURVEY.md App. E.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/porla_multiexp.h"
#include "host_bn254.hpp"
#include "msm.h"

using namespace porla;

namespace {

typedef unsigned __int128 u128;

// secp256k1_fe (5x52, any magnitude) -> canonical 8x32 LE limbs below p
// (the effect of secp256k1_fe_normalize_var, field_5x52_impl.h:128-170, then fe_get_b32 :344-381)
void fe_to_canonical(const porla_secp256k1_fe* a, uint32_t out[8]) {
    uint32_t w[11] = {0};  // 352 bits are plenty: sum of n[i] << 52i < 2^273
    for (int i = 0; i < 5; i++) {
        int bit = 52 * i;
        u128 v = (u128)a->n[i] << (bit & 31);
        int k = bit >> 5;
        uint64_t carry = 0;
        for (int j = 0; j < 4 && k + j < 11; j++) {
            uint64_t s = (uint64_t)w[k + j] + (uint32_t)(v >> (32 * j)) + carry;
            w[k + j] = (uint32_t)s;
            carry = s >> 32;
        }
        for (int j = k + 4; carry && j < 11; j++) {
            uint64_t s = (uint64_t)w[j] + carry;
            w[j] = (uint32_t)s;
            carry = s >> 32;
        }
    }
    // fold limbs 8.. : 2^256 = 2^32 + 977 (mod p), twice
    for (int round = 0; round < 3; round++) {
        uint64_t hi = (uint64_t)w[8] | ((uint64_t)w[9] << 32);
        if (w[10]) { fprintf(stderr, "[libmultiexp/porla_b200] FATAL: secp256k1_fe magnitude too large\n"); abort(); }
        if (!hi) break;
        w[8] = w[9] = 0;
        u128 add0 = (u128)hi * 977u;          // at limb 0
        u128 add1 = (u128)hi;                 // at limb 1
        uint64_t carry = 0;
        for (int j = 0; j < 10; j++) {
            u128 s = (u128)w[j] + carry;
            if (j < 4) s += (uint32_t)(add0 >> (32 * j));
            if (j >= 1 && j < 4) s += (uint32_t)(add1 >> (32 * (j - 1)));
            w[j] = (uint32_t)s;
            carry = (uint64_t)(s >> 32);
        }
    }
    memcpy(out, w, 32);
    host::reduce_canonical<SecpFp>(out);
}

// canonical 8x32 -> 5x52 (fe_set_b32 layout, field_5x52_impl.h:294-341)
void canonical_to_fe(const uint32_t in[8], porla_secp256k1_fe* r) {
    uint64_t q[4];
    for (int i = 0; i < 4; i++) q[i] = (uint64_t)in[2 * i] | ((uint64_t)in[2 * i + 1] << 32);
    const uint64_t M = 0xFFFFFFFFFFFFFull;
    r->n[0] = q[0] & M;
    r->n[1] = ((q[0] >> 52) | (q[1] << 12)) & M;
    r->n[2] = ((q[1] >> 40) | (q[2] << 24)) & M;
    r->n[3] = ((q[2] >> 28) | (q[3] << 36)) & M;
    r->n[4] = q[3] >> 16;
}

void gej_set_infinity(porla_secp256k1_gej* r) {
    memset(r, 0, sizeof(*r));
    r->infinity = 1;
}

const uint8_t kGenLE[64] = {
    // x, y of G as 8 LE 32-bit limbs each (little-endian bytes)
    0x98, 0x17, 0xF8, 0x16, 0x5B, 0x81, 0xF2, 0x59, 0xD9, 0x28, 0xCE, 0x2D, 0xDB, 0xFC, 0x9B, 0x02,
    0x07, 0x0B, 0x87, 0xCE, 0x95, 0x62, 0xA0, 0x55, 0xAC, 0xBB, 0xDC, 0xF9, 0x7E, 0x66, 0xBE, 0x79,
    0xB8, 0xD4, 0x10, 0xFB, 0x8F, 0xD0, 0x47, 0x9C, 0x19, 0x54, 0x85, 0xA6, 0x48, 0xB4, 0x17, 0xFD,
    0xA8, 0x08, 0x11, 0x0E, 0xFC, 0xFB, 0xA4, 0x5D, 0x65, 0xC4, 0xA3, 0x26, 0x77, 0xDA, 0x3A, 0x48};

template <class C>
void point_add_host(const uint8_t* a, const uint8_t* b, int64_t n, int fmt, uint8_t* out) {
    using F = typename C::F;
    for (int64_t i = 0; i < n; i++) {
        Affine<F> p, q;
        const uint8_t* pa = a + 64 * i;
        const uint8_t* pb = b + 64 * i;
        auto load = [&](const uint8_t* s, F* x) {
            if (fmt == PORLA_POINT_BE64) host::be32_to_limbs(s, x->v);
            else memcpy(x->v, s, 32);
            host::reduce_canonical<F>(x->v);
            *x = x->to_internal();
        };
        load(pa, &p.x); load(pa + 32, &p.y); load(pb, &q.x); load(pb + 32, &q.y);
        XYZZ<F> r = XYZZ<F>::from_affine(p);
        r.madd(q);
        Affine<F> s = r.to_affine();
        F x = s.x.from_internal(), y = s.y.from_internal();
        if (fmt == PORLA_POINT_BE64) {
            host::limbs_to_be32(x.v, out + 64 * i);
            host::limbs_to_be32(y.v, out + 64 * i + 32);
        } else {
            memcpy(out + 64 * i, x.v, 32);
            memcpy(out + 64 * i + 32, y.v, 32);
        }
    }
}

}  // namespace

extern "C" {

int porla_secp256k1_ecmult_multi_var(const porla_secp256k1_callback* error_callback, void* scratch,
                                     porla_secp256k1_gej* r, const porla_secp256k1_scalar* inp_g_sc,
                                     porla_secp256k1_ecmult_multi_callback cb, void* cbdata, size_t n) {
    (void)error_callback;
    (void)scratch;
    gej_set_infinity(r);  // ecmult_impl.h:821
    if (inp_g_sc == NULL && n == 0) return 1;
    std::vector<uint8_t> scalars, points;
    scalars.reserve((n + 1) * 32);
    points.reserve((n + 1) * 64);
    if (inp_g_sc) {
        const uint64_t* d = inp_g_sc->d;
        if (d[0] | d[1] | d[2] | d[3]) {
            scalars.insert(scalars.end(), (const uint8_t*)d, (const uint8_t*)d + 32);
            points.insert(points.end(), kGenLE, kGenLE + 64);
        }
    }
    for (size_t i = 0; i < n; i++) {
        porla_secp256k1_scalar sc;
        porla_secp256k1_ge pt;
        if (!cb(&sc, &pt, i, cbdata)) return 0;  // ecmult_impl.h:696
        if (pt.infinity) continue;               // ecmult_impl.h:500-502 skips them
        if (!(sc.d[0] | sc.d[1] | sc.d[2] | sc.d[3])) continue;
        uint32_t xy[16];
        fe_to_canonical(&pt.x, xy);
        fe_to_canonical(&pt.y, xy + 8);
        scalars.insert(scalars.end(), (const uint8_t*)sc.d, (const uint8_t*)sc.d + 32);
        points.insert(points.end(), (const uint8_t*)xy, (const uint8_t*)xy + 64);
    }
    size_t m = scalars.size() / 32;
    if (m == 0) return 1;
    uint8_t res[64];
    porla_msm_host(PORLA_CURVE_SECP256K1, scalars.data(), points.data(), (int64_t)m, 1, PORLA_SCALAR_LE32,
                   PORLA_POINT_LE64, res);
    uint32_t xy[16];
    memcpy(xy, res, 64);
    bool inf = true;
    for (int i = 0; i < 16; i++) inf = inf && xy[i] == 0;
    if (inf) return 1;
    canonical_to_fe(xy, &r->x);
    canonical_to_fe(xy + 8, &r->y);
    memset(&r->z, 0, sizeof(r->z));
    r->z.n[0] = 1;
    r->infinity = 0;
    return 1;
}

// ---- generators resident in HBM (BASELINE north star: "SRS and generator tables resident in HBM, uploaded
// once at init").  Porla's IPA mode multiplies the SAME generator array in every commitment, align_MAC and
// inner-product round (data.pt = &generators[start_chunk]: Server.hpp:347,507,2345,2395, Client.hpp:393,1588);
// the callback adapter above re-uploads them on every call.
porla_table* porla_secp256k1_table_create(const porla_secp256k1_ge* points, size_t n) {
    std::vector<uint8_t> le(n * 64 + 64, 0);
    for (size_t i = 0; i < n; i++) {
        if (points[i].infinity) continue;           // stays 64 zero bytes = infinity
        uint32_t xy[16];
        fe_to_canonical(&points[i].x, xy);
        fe_to_canonical(&points[i].y, xy + 8);
        memcpy(le.data() + 64 * i, xy, 64);
    }
    porla_table* t = porla_table_create(PORLA_CURVE_SECP256K1, le.data(), (int64_t)n, PORLA_POINT_LE64, 0, nullptr);
    // generators are fixed bases: expand them once (small sets such as Porla's 128 get the full look-up table of
    // window multiples, so every commitment / IPA round over them is a plain sum of table entries)
    if (n) porla_table_precompute(t, 0, (int64_t)n, 1, nullptr);
    return t;
}

int porla_secp256k1_ecmult_multi_table(const porla_table* t, size_t first, const porla_secp256k1_scalar* scalars, size_t n,
                                       porla_secp256k1_gej* r) {
    gej_set_infinity(r);
    if (n == 0) return 1;
    if ((int64_t)(first + n) > porla_table_len(t)) return 0;
    uint8_t res[64];
    porla_msm_table_host_scalars(t, (int64_t)first, scalars, (int64_t)n, PORLA_SCALAR_LE32, PORLA_POINT_LE64, res);
    uint32_t xy[16];
    memcpy(xy, res, 64);
    bool inf = true;
    for (int i = 0; i < 16; i++) inf = inf && xy[i] == 0;
    if (inf) return 1;
    canonical_to_fe(xy, &r->x);
    canonical_to_fe(xy + 8, &r->y);
    memset(&r->z, 0, sizeof(r->z));
    r->z.n[0] = 1;
    r->infinity = 0;
    return 1;
}

int porla_secp256k1_gej_serialize(const porla_secp256k1_gej* a, unsigned char out33[33]) {
    if (a->infinity) return 0;
    SecpFp x, y, z;
    fe_to_canonical(&a->x, x.v);
    fe_to_canonical(&a->y, y.v);
    fe_to_canonical(&a->z, z.v);
    SecpFp zi = z.inverse(), zi2 = zi.sqr();
    x = x * zi2;
    y = y * zi2 * zi;
    out33[0] = (y.v[0] & 1u) ? 0x03 : 0x02;  // SECP256K1_TAG_PUBKEY_ODD / EVEN
    host::limbs_to_be32(x.v, out33 + 1);
    return 1;
}

void porla_debug_point_add_host(int curve, const void* a, const void* b, int64_t n, int point_fmt, void* out) {
    if (curve == PORLA_CURVE_BN254) point_add_host<Bn254>((const uint8_t*)a, (const uint8_t*)b, n, point_fmt, (uint8_t*)out);
    else point_add_host<Secp256k1>((const uint8_t*)a, (const uint8_t*)b, n, point_fmt, (uint8_t*)out);
}

}  // extern "C"
