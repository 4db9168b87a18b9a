// Config-1 replay: the C-ABI call census of Porla's KZG mode for ./Client 1024 (NUM_CHUNKS = 128, TOP_CACHING_LEVEL = 10),
// SURVEY.md Appendix C, issued by a C++ caller exactly as the reference's glue issues it (GoSlices over caller-owned heap
// buffers, /root/reference/porla/Utils/utils.h:235-305), against libmultiexp.so.  Porla's own Client/Server need NTL and
// ZeroMQ and cannot be built here; this program restates their MAC-side control flow with synthetic data:
//
//   server update        Server::update -> HAdd -> HRebuildX / HRebuildY -> mix   (Server.hpp:401-476, 1388-1478, 1330-1386, 1209-1328)
//   server C rebuild     Server::CRebuild_Cached                                 (Server.hpp:1487-1830)
//   server audit         Server::audit, KZG branch                               (Server.hpp:564-931)
//   client audit check   Client::audit's verification calls                      (Client.hpp:633-892; census of SURVEY App. C)
//   one round            Client::self_test: 1024 updates, then 100 audits        (Client.hpp:894-919)
//
// Three passes over the same inputs:
//   legacy   every call through the 14 symbols of libmultiexp.h, butterflies on an 8-thread pool as the reference runs them
//   batched  the same work through the batched symbols (bn254_butterfly_stage, bn254_audit_aggregate)
//   cpu      the per-call pattern against the CPU restatement (oracle/liboracle_bn254.so, dlopen'ed; test infrastructure)
// The MAC arrays all passes end with must be identical, byte for byte; so must every audit reply.
//
// Built with -DPORLA_USE_REFERENCE_HEADER and -I/root/reference/porla/Utils when that tree is present: the legacy
// prototypes then come from the reference's own cgo header.
#ifdef PORLA_USE_REFERENCE_HEADER
#include "libmultiexp.h"
extern "C" {
void bn254_butterfly_stage(GoSlice* points, GoInt n, GoInt m, GoSlice* twiddles);
void bn254_audit_aggregate(GoSlice* coefs, GoSlice* blocks, GoInt n, GoSlice* b_out, GoSlice* align_out);
int porla_device_count(void);
}
#else
#include "porla_multiexp.h"
#endif

#include <dlfcn.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int kChunks = 128;          // NUM_CHUNKS = BLOCK_SIZE >> 5 (config.hpp:21-22)
constexpr int kThreads = 8;           // MAX_NUM_THREADS_SERVER (config.hpp:16)
constexpr int kAuditPoints = 128;     // NUM_CHECK_AUDIT (config.hpp:32)

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

struct Pt {
    uint8_t b[64];
    bool operator==(const Pt& o) const { return memcmp(b, o.b, 64) == 0; }
};
struct Sc {
    uint8_t b[32];   // bn254_scalar: 32-byte big-endian (utils.h:307-318)
};

GoSlice slice(void* p, long long n) {
    GoSlice s;
    s.data = p;
    s.len = s.cap = n;
    return s;
}

// ---------------------------------------------------------------------------- 256-bit helpers (twiddles mod PRIME_MODULUS)
struct U256 {
    uint32_t w[8];
};
const U256 kPrime = {{0x00000001u, 0, 0, 0, 0, 0, 0, 0xcf000000u}};   // PRIME_MODULUS = 207 * 2^248 + 1 (utils.h:40)
bool geq(const U256& a, const U256& b) {
    for (int i = 7; i >= 0; i--)
        if (a.w[i] != b.w[i]) return a.w[i] > b.w[i];
    return true;
}
uint32_t add_to(U256& a, const U256& b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.w[i] + b.w[i];
        a.w[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
void sub_from(U256& a, const U256& b) {
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.w[i] - b.w[i] - br;
        a.w[i] = (uint32_t)d;
        br = (d >> 63) & 1;
    }
}
// a * b mod PRIME_MODULUS = 207 * 2^248 + 1 (values below the modulus).  Schoolbook product on 64-bit limbs, then the
// special form: 207 * 2^248 = -1, so with the product t = hi * 2^248 + lo and hi = 207 q + r it is  r * 2^248 + lo - q.
// (The caller-side twiddle arithmetic is NTL's ZZ_p in the reference; a bit-serial product here cost 0.6 ms per power and
// hid the library's share of an update.)
U256 mulmod(const U256& a, const U256& b) {
    typedef unsigned __int128 u128;
    uint64_t x[4], y[4], t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        x[i] = (uint64_t)a.w[2 * i] | ((uint64_t)a.w[2 * i + 1] << 32);
        y[i] = (uint64_t)b.w[2 * i] | ((uint64_t)b.w[2 * i + 1] << 32);
    }
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)x[j] * y[i] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + 4] = (uint64_t)c;
    }
    // hi = t >> 248 (at most 264 bits: 5 limbs), lo = t mod 2^248
    uint64_t hi[5], lo[4] = {t[0], t[1], t[2], t[3] & 0x00ffffffffffffffull};
    for (int i = 0; i < 5; i++) hi[i] = (t[i + 3] >> 56) | (i + 4 < 9 ? t[i + 4] << 8 : 0);
    // q = hi / 207, r = hi % 207
    uint64_t q[5];
    u128 rem = 0;
    for (int i = 4; i >= 0; i--) {
        u128 cur = (rem << 64) | hi[i];
        q[i] = (uint64_t)(cur / 207);
        rem = cur % 207;
    }
    // v = r * 2^248 + lo - q  as a signed 5-limb value, then bring it into [0, p)
    uint64_t v[5] = {lo[0], lo[1], lo[2], lo[3] | ((uint64_t)rem << 56), 0};
    uint64_t br = 0;
    for (int i = 0; i < 5; i++) {
        u128 d = (u128)v[i] - q[i] - br;
        v[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
    const uint64_t pm[5] = {1, 0, 0, 0xcf00000000000000ull, 0};
    while ((int64_t)v[4] < 0) {           // negative: add p (q < 2^257 / 1, so a few rounds at most)
        u128 c = 0;
        for (int i = 0; i < 5; i++) {
            c += (u128)v[i] + pm[i];
            v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    for (;;) {                            // subtract p while v >= p
        bool ge = v[4] != 0;
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--)
                if (v[i] != pm[i]) { ge = v[i] > pm[i]; break; }
        }
        if (!ge) break;
        uint64_t b2 = 0;
        for (int i = 0; i < 5; i++) {
            u128 d = (u128)v[i] - pm[i] - b2;
            v[i] = (uint64_t)d;
            b2 = (uint64_t)(d >> 64) & 1;
        }
    }
    U256 r;
    for (int i = 0; i < 4; i++) {
        r.w[2 * i] = (uint32_t)v[i];
        r.w[2 * i + 1] = (uint32_t)(v[i] >> 32);
    }
    return r;
}
U256 powmod(U256 base, uint64_t e_lo, int extra_pow2 = 0) {   // base^(e_lo * 2^extra_pow2)
    U256 r = {{1, 0, 0, 0, 0, 0, 0, 0}};
    for (int bit = 63; bit >= 0; bit--) {
        r = mulmod(r, r);
        if ((e_lo >> bit) & 1u) r = mulmod(r, base);
    }
    for (int i = 0; i < extra_pow2; i++) r = mulmod(r, r);
    return r;
}
Sc to_scalar(const U256& v) {   // convert_ZZ_to_scalar (utils.h:307-318): word 0 most significant, htonl each
    Sc s;
    for (int i = 0; i < 8; i++) {
        uint32_t w = v.w[7 - i];
        s.b[4 * i] = (uint8_t)(w >> 24);
        s.b[4 * i + 1] = (uint8_t)(w >> 16);
        s.b[4 * i + 2] = (uint8_t)(w >> 8);
        s.b[4 * i + 3] = (uint8_t)w;
    }
    return s;
}
Sc scalar_from_int(uint32_t v) {   // bn254_scalar_set_int (utils.h:271-275)
    Sc s;
    memset(s.b, 0, 32);
    s.b[28] = (uint8_t)(v >> 24);
    s.b[29] = (uint8_t)(v >> 16);
    s.b[30] = (uint8_t)(v >> 8);
    s.b[31] = (uint8_t)v;
    return s;
}
int reverse_bits(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++)
        if ((v >> i) & 1) r |= 1 << (bits - 1 - i);
    return r;
}

// rem = a mod m for a little-endian limb array (align_MAC's "% PRIME_MODULUS" and "% GROUP_ORDER", Server.hpp:531-540)
const uint32_t kOrder[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
void mod_wide(const uint32_t* a, int limbs, const uint32_t* m, uint32_t* rem) {
    memset(rem, 0, 32);
    for (int bit = limbs * 32 - 1; bit >= 0; bit--) {
        uint32_t top = rem[7] >> 31;
        for (int k = 7; k > 0; k--) rem[k] = (rem[k] << 1) | (rem[k - 1] >> 31);
        rem[0] = (rem[0] << 1) | ((a[bit >> 5] >> (bit & 31)) & 1u);
        uint32_t t[8];
        uint64_t br = 0;
        for (int k = 0; k < 8; k++) {
            uint64_t d = (uint64_t)rem[k] - m[k] - br;
            t[k] = (uint32_t)d;
            br = (d >> 63) & 1;
        }
        if (top || !br) memcpy(rem, t, 32);
    }
}

// ---------------------------------------------------------------------------- the three back ends
struct Ops {
    const char* name;
    std::function<void(Pt&, const Sc&)> mult;                               // bn254_mult
    std::function<void(Pt&, const Pt&)> add;                                // bn254_add
    std::function<void(Pt&)> neg;                                           // bn254_neg
    std::function<void(Pt&)> set_inf;                                       // bn254_set_infinity
    std::function<void(Pt&, const Pt*, const Sc*, int)> multi_exp;          // bn254_multi_exp
    std::function<void(const Sc*, Pt&)> digest_from_srs;                    // compute_digest_from_srs over 128 scalars
    bool batched = false;
    bool device = true;
};

Ops library_ops(bool batched) {
    Ops o;
    o.name = batched ? "batched" : "legacy";
    o.batched = batched;
    o.mult = [](Pt& a, const Sc& s) {
        GoSlice pa = slice(a.b, 64), ps = slice((void*)s.b, 32);
        mult_point(&pa, &ps);
    };
    o.add = [](Pt& a, const Pt& b) {
        GoSlice pa = slice(a.b, 64), pb = slice((void*)b.b, 64);
        add_point(&pa, &pb);
    };
    o.neg = [](Pt& a) {
        GoSlice pa = slice(a.b, 64);
        neg_point(&pa);
    };
    o.set_inf = [](Pt& a) {
        GoSlice pa = slice(a.b, 64);
        set_inf_point(&pa);
    };
    o.multi_exp = [](Pt& r, const Pt* pts, const Sc* sc, int n) {
        GoSlice gs = slice((void*)sc, (long long)n << 5), gp = slice((void*)pts, (long long)n << 6), gr = slice(r.b, 64);
        compute_multi_exp(&gs, &gp, n, &gr);
    };
    o.digest_from_srs = [](const Sc* sc, Pt& out) {
        GoSlice gi = slice((void*)sc, kChunks * 32), go = slice(out.b, 64);
        compute_digest_from_srs(&gi, &go);
    };
    return o;
}

// CPU restatement behind the same per-call pattern (oracle: test infrastructure, loaded only on request)
typedef void (*oracle_msm_fn)(const uint8_t*, const uint8_t*, size_t, uint8_t*, int);
typedef void (*oracle_add_fn)(const uint8_t*, const uint8_t*, uint8_t*);
typedef void (*oracle_mul_fn)(const uint8_t*, const uint8_t*, uint8_t*);
bool cpu_ops(const char* path, const std::vector<Pt>& srs, Ops* o) {
    void* h = dlopen(path, RTLD_NOW);
    if (!h) return false;
    auto msm = (oracle_msm_fn)dlsym(h, "oracle_bn254_msm");
    auto add = (oracle_add_fn)dlsym(h, "oracle_bn254_add");
    auto mul = (oracle_mul_fn)dlsym(h, "oracle_bn254_mul");
    if (!msm || !add || !mul) return false;
    const int hw = (int)std::thread::hardware_concurrency() > 0 ? (int)std::thread::hardware_concurrency() : 1;
    o->name = "cpu";
    o->device = false;
    o->mult = [mul](Pt& a, const Sc& s) {
        Pt r;
        mul(a.b, s.b, r.b);
        a = r;
    };
    o->add = [add](Pt& a, const Pt& b) {
        Pt r;
        add(a.b, b.b, r.b);
        a = r;
    };
    // -P = (x, p - y); the restatement has no negation entry, the byte arithmetic is done here
    o->neg = [](Pt& a) {
        static const uint8_t p[32] = {0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45, 0xb6, 0x81, 0x81, 0x58, 0x5d,
                                      0x97, 0x81, 0x6a, 0x91, 0x68, 0x71, 0xca, 0x8d, 0x3c, 0x20, 0x8c, 0x16, 0xd8, 0x7c, 0xfd, 0x47};
        bool zero = true;
        for (int i = 0; i < 32; i++) zero = zero && a.b[32 + i] == 0;
        if (zero) return;
        int br = 0;
        for (int i = 31; i >= 0; i--) {
            int d = (int)p[i] - a.b[32 + i] - br;
            br = d < 0;
            a.b[32 + i] = (uint8_t)(d + (br ? 256 : 0));
        }
    };
    o->set_inf = [](Pt& a) { memset(a.b, 0, 64); };
    o->multi_exp = [msm, hw](Pt& r, const Pt* pts, const Sc* sc, int n) { msm(sc->b, pts->b, (size_t)n, r.b, n >= 4096 ? hw : 1); };
    o->digest_from_srs = [msm, &srs](const Sc* sc, Pt& out) { msm(sc->b, srs[0].b, kChunks, out.b, 1); };
    return true;
}

// 8-thread pool pattern of the reference (a ThreadPool built per loop, Server.hpp:1215, :1510): ranges of [0, n)
void parallel_ranges(int n, const std::function<void(int, int)>& body) {
    const int nt = n >= kThreads ? kThreads : n;
    if (nt <= 1) {
        body(0, n);
        return;
    }
    std::vector<std::thread> th;
    const int each = n / nt;
    for (int t = 0; t < nt; t++) th.emplace_back(body, t * each, t == nt - 1 ? n : (t + 1) * each);
    for (auto& x : th) x.join();
}

// ---------------------------------------------------------------------------- server state (MAC side)
struct Level {
    std::vector<Pt> mac_x, mac_y, al_x, al_y;   // 2 * 2^level entries each (Server.hpp: MAC_commitments_H / MAC_alignments_H)
    bool empty = true;
};

struct Inputs {
    int n_blocks, height;
    std::vector<Pt> mac_u;                 // client MACs of the blocks, as Server::update receives them
    std::vector<Pt> complements;           // MAC hiding parts (a pool; each update adds 2^(L+1) of them)
    std::vector<std::vector<Sc>> align_sc; // per update: the 128 scalars align_MAC commits to
    std::vector<uint8_t> blocks;           // data blocks for the audit aggregation: n_blocks x 128 chunks x 64 B (LE)
    U256 w;                                // 2 * n_blocks-th ... root of unity mod PRIME_MODULUS used for the twiddles
    std::vector<Pt> srs;                   // [tau^i] G, uncompressed (for the CPU pass)
};

struct Server {
    const Inputs& in;
    const Ops& op;
    std::vector<Level> H;
    int write_step = 0;
    double t_butterfly_ms = 0, t_align_ms = 0, t_crebuild_ms = 0;
    long n_butterflies = 0;

    Server(const Inputs& i, const Ops& o) : in(i), op(o), H(i.height) {
        for (int l = 0; l < in.height; l++) {
            const size_t n = (size_t)2 << l;
            // the second half of a level is the mixing target (2 * length offset, Server.hpp:1227-1229): 4 * 2^l entries
            H[l].mac_x.assign(2 * n, Pt{});
            H[l].mac_y.assign(2 * n, Pt{});
            H[l].al_x.assign(2 * n, Pt{});
            H[l].al_y.assign(2 * n, Pt{});
        }
    }

    std::vector<Sc> twiddles(int count, int stride_exp_blocks) {   // v^j, v = w^(num_blocks / stride)
        std::vector<Sc> tw((size_t)count);
        const U256 v = powmod(in.w, (uint64_t)stride_exp_blocks);
        U256 vi = {{1, 0, 0, 0, 0, 0, 0, 0}};
        for (int j = 0; j < count; j++) {
            tw[(size_t)j] = to_scalar(vi);
            vi = mulmod(vi, v);
        }
        return tw;
    }

    // One butterfly as the reference issues it (Server.hpp:1289-1310): t = v * A1; out0 = A0 + t; out1 = A0 - t
    void butterfly(const Pt& a0, const Pt& a1, const Sc& v, Pt& out0, Pt& out1) {
        Pt t = a1;
        op.mult(t, v);
        out0 = a0;
        op.add(out0, t);
        op.neg(t);
        out1 = a0;
        op.add(out1, t);
    }

    // Server::mix for one side: level `lvl` (first half A0, second half A1) -> second half of level lvl + 1
    void mix(bool is_x, int lvl) {
        const int length = 1 << lvl;
        std::vector<Pt>& src_m = is_x ? H[lvl].mac_x : H[lvl].mac_y;
        std::vector<Pt>& src_a = is_x ? H[lvl].al_x : H[lvl].al_y;
        std::vector<Pt>& dst_m = is_x ? H[lvl + 1].mac_x : H[lvl + 1].mac_y;
        std::vector<Pt>& dst_a = is_x ? H[lvl + 1].al_x : H[lvl + 1].al_y;
        const std::vector<Sc> tw = twiddles(length, in.n_blocks / length);
        auto t0 = Clock::now();
        if (op.batched) {
            // MAC and alignment arrays of the side in ONE stage call: [A0 | A1] per array, m = 2 * length
            std::vector<Pt> buf((size_t)4 * length);
            memcpy(&buf[0], &src_m[0], (size_t)2 * length * 64);
            memcpy(&buf[(size_t)2 * length], &src_a[0], (size_t)2 * length * 64);
            GoSlice gp = slice(buf.data(), (long long)buf.size() * 64), gt = slice((void*)tw.data(), (long long)length * 32);
            bn254_butterfly_stage(&gp, 4 * length, 2 * length, &gt);
            memcpy(&dst_m[(size_t)2 * length], &buf[0], (size_t)2 * length * 64);
            memcpy(&dst_a[(size_t)2 * length], &buf[(size_t)2 * length], (size_t)2 * length * 64);
        } else {
            parallel_ranges(length, [&](int lo, int hi) {
                for (int i = lo; i < hi; i++) {
                    butterfly(src_m[i], src_m[length + i], tw[i], dst_m[2 * length + i], dst_m[2 * length + i + length]);
                    butterfly(src_a[i], src_a[length + i], tw[i], dst_a[2 * length + i], dst_a[2 * length + i + length]);
                }
            });
        }
        t_butterfly_ms += ms_since(t0);
        n_butterflies += 2 * length;
    }

    void hrebuild(bool is_x, int level) {   // HRebuildX / HRebuildY (Server.hpp:1330-1386)
        for (int i = 0; i < level; i++) {
            mix(is_x, i);
            if (!is_x) H[i].empty = true;
        }
        const int n = 1 << level;
        std::vector<Pt>& m = is_x ? H[level].mac_x : H[level].mac_y;
        std::vector<Pt>& a = is_x ? H[level].al_x : H[level].al_y;
        for (int i = 0; i < n; i++) {
            m[i] = m[n + i];
            a[i] = a[n + i];
        }
        if (!is_x) H[level].empty = false;
    }

    // align_MAC (Server.hpp:478-562, KZG branch): one 128-term commitment over the SRS, added to B
    void align_mac(const std::vector<Sc>& sc, Pt& b) {
        auto t0 = Clock::now();
        Pt v;
        op.digest_from_srs(sc.data(), v);
        op.add(b, v);
        t_align_ms += ms_since(t0);
    }

    int hadd(int block) {   // Server::HAdd (Server.hpp:1388-1478)
        const Pt& mac = in.mac_u[(size_t)block];
        const Sc wt = to_scalar(powmod(in.w, (uint64_t)reverse_bits(write_step % in.n_blocks, in.height - 1)));
        Pt mac_b2 = mac;
        op.mult(mac_b2, wt);
        Pt al, al_b2;
        op.set_inf(al);
        op.set_inf(al_b2);
        align_mac(in.align_sc[(size_t)block], al_b2);
        int level = 0;
        if (H[0].empty) {
            H[0].mac_x[0] = mac;
            H[0].mac_y[0] = mac_b2;
            H[0].al_x[0] = al;
            H[0].al_y[0] = al_b2;
            H[0].empty = false;
        } else {
            level = 1;
            while (!H[level].empty) level++;
            H[0].mac_x[1] = mac;
            H[0].mac_y[1] = mac_b2;
            H[0].al_x[1] = al;
            H[0].al_y[1] = al_b2;
            hrebuild(true, level);
            hrebuild(false, level);
        }
        return level;
    }

    void crebuild() {   // Server::CRebuild_Cached (Server.hpp:1487-1830)
        auto t0 = Clock::now();
        const int top = in.height - 1, n = in.n_blocks;
        for (int l = 0; l < top; l++) H[l].empty = true;
        const Sc wt = to_scalar(powmod(in.w, (uint64_t)reverse_bits(write_step % in.n_blocks, in.height - 1)));
        Level& T = H[top];
        parallel_ranges(n, [&](int lo, int hi) {
            for (int i = lo; i < hi; i++) {
                T.mac_x[i] = in.mac_u[(size_t)i];
                T.mac_y[i] = in.mac_u[(size_t)i];
                op.mult(T.mac_y[i], wt);
                op.set_inf(T.al_x[i]);
                op.set_inf(T.al_y[i]);
            }
        });
        for (int side = 0; side < 2; side++) {
            std::vector<Pt>& a = side == 0 ? T.mac_x : T.mac_y;
            for (int s = 1; s < in.height; s++) {
                const int m = 1 << s, m2 = m >> 1;
                const std::vector<Sc> tw = twiddles(m2, in.n_blocks / m2);
                if (op.batched) {
                    GoSlice gp = slice(a.data(), (long long)n * 64), gt = slice((void*)tw.data(), (long long)m2 * 32);
                    bn254_butterfly_stage(&gp, n, m, &gt);
                } else {
                    // n/2 butterflies of the stage, spread over the pool like Server.hpp:1558-1680
                    parallel_ranges(n / 2, [&](int lo, int hi) {
                        for (int b = lo; b < hi; b++) {
                            const int j = b % m2, k = (b / m2) * m + j;
                            Pt o0, o1;
                            butterfly(a[k], a[k + m2], tw[j], o0, o1);
                            a[k] = o0;
                            a[k + m2] = o1;
                        }
                    });
                }
                n_butterflies += n / 2;
            }
        }
        T.empty = false;
        t_crebuild_ms += ms_since(t0);
    }

    void update(int block) {   // Server::update (Server.hpp:401-476), MAC side
        write_step++;
        int level = in.height - 1;
        if (write_step % in.n_blocks == 0) crebuild();
        else level = hadd(block);
        const int l = 1 << level;
        for (int i = 0; i < (l << 1); i++) {
            const Pt& c = in.complements[(size_t)((write_step * 131 + i) % (int)in.complements.size())];
            if (i >= l) op.add(H[level].mac_y[i - l], c);
            else op.add(H[level].mac_x[i], c);
        }
    }
};

struct AuditReply {
    Pt combined_mac, combined_align, commitment, proof_h;
    uint8_t point[32], claim[32];
    std::vector<uint8_t> b_mod;   // B % PRIME_MODULUS, 128 x 32 B big-endian
};

// Server::audit after a C rebuild: only the top level is live, n_points = NUM_CHECK_AUDIT
AuditReply server_audit(const Inputs& in, const Ops& op, const Server& sv, uint64_t seed, double* ms_msm, double* ms_agg,
                        double* ms_proof) {
    std::mt19937_64 rng(seed);
    const int top = in.height - 1, l = 1 << top;
    const Level& T = sv.H[top];
    std::vector<Sc> sc(kAuditPoints);
    std::vector<Pt> ptc(kAuditPoints), pta(kAuditPoints);
    std::vector<uint32_t> coefs(kAuditPoints);
    std::vector<int> rows(kAuditPoints);
    for (int j = 0; j < kAuditPoints; j++) {
        const int index = (int)(rng() % (uint64_t)(l << 1));
        const uint32_t coeff = (uint32_t)(rng() & 0x7fffffffu);   // abs(int) (Server.hpp:611)
        coefs[j] = coeff;
        sc[j] = scalar_from_int(coeff);
        if (index >= l) {
            ptc[j] = T.mac_y[index - l];
            pta[j] = T.al_y[index - l];
        } else {
            ptc[j] = T.mac_x[index];
            pta[j] = T.al_x[index];
        }
        rows[j] = index % in.n_blocks;
    }
    AuditReply r;
    auto t0 = Clock::now();
    op.multi_exp(r.combined_mac, ptc.data(), sc.data(), kAuditPoints);      // Server.hpp:900
    op.multi_exp(r.combined_align, pta.data(), sc.data(), kAuditPoints);    // Server.hpp:901
    *ms_msm += ms_since(t0);
    // B = sum coef_i * block_i (Server.hpp:790-828), then align_MAC(B, combined_align) (Server.hpp:903)
    t0 = Clock::now();
    r.b_mod.assign(kChunks * 32, 0);
    std::vector<uint8_t> gathered((size_t)kAuditPoints * kChunks * 64);
    for (int j = 0; j < kAuditPoints; j++)
        memcpy(&gathered[(size_t)j * kChunks * 64], &in.blocks[(size_t)rows[j] * kChunks * 64], (size_t)kChunks * 64);
    if (op.batched) {
        Pt al;
        GoSlice gc = slice(coefs.data(), kAuditPoints * 4), gb = slice(gathered.data(), (long long)gathered.size()),
                gout = slice(r.b_mod.data(), kChunks * 32), gal = slice(al.b, 64);
        bn254_audit_aggregate(&gc, &gb, kAuditPoints, &gout, &gal);
        op.add(r.combined_align, al);
    } else {
        std::vector<Sc> csc(kChunks);
        parallel_ranges(kChunks, [&](int lo, int hi) {
            for (int c = lo; c < hi; c++) {
                uint32_t acc[18] = {0};
                for (int j = 0; j < kAuditPoints; j++) {
                    const uint32_t* a = reinterpret_cast<const uint32_t*>(&gathered[((size_t)j * kChunks + c) * 64]);
                    uint64_t carry = 0;
                    for (int k = 0; k < 16; k++) {
                        uint64_t v = (uint64_t)a[k] * coefs[j] + acc[k] + carry;
                        acc[k] = (uint32_t)v;
                        carry = v >> 32;
                    }
                    for (int k = 16; k < 18; k++) {
                        uint64_t v = (uint64_t)acc[k] + carry;
                        acc[k] = (uint32_t)v;
                        carry = v >> 32;
                    }
                }
                uint32_t rem[8], d[18], t[8], cc[8];
                mod_wide(acc, 18, kPrime.w, rem);
                uint64_t br = 0;
                for (int k = 0; k < 18; k++) {
                    uint64_t v = (uint64_t)acc[k] - (k < 8 ? rem[k] : 0u) - br;
                    d[k] = (uint32_t)v;
                    br = (v >> 63) & 1;
                }
                mod_wide(d, 18, kOrder, t);
                bool zero = true;
                for (int k = 0; k < 8; k++) zero = zero && t[k] == 0;
                br = 0;
                for (int k = 0; k < 8; k++) {
                    uint64_t v = (uint64_t)kOrder[k] - t[k] - br;
                    cc[k] = zero ? 0u : (uint32_t)v;
                    br = (v >> 63) & 1;
                }
                U256 ru, cu;
                memcpy(ru.w, rem, 32);
                memcpy(cu.w, cc, 32);
                csc[c] = to_scalar(cu);
                const Sc bs = to_scalar(ru);
                memcpy(&r.b_mod[(size_t)c * 32], bs.b, 32);
            }
        });
        Pt al;
        op.digest_from_srs(csc.data(), al);
        op.add(r.combined_align, al);
    }
    *ms_agg += ms_since(t0);
    // create_kzg_proof (Server.hpp:363-398): library host + device code in every pass (the proof needs the SRS state)
    t0 = Clock::now();
    GoSlice gi = slice(r.b_mod.data(), kChunks * 32), gc2 = slice(r.commitment.b, 64), gh = slice(r.proof_h.b, 64),
            gpnt = slice(r.point, 32), gcl = slice(r.claim, 32);
    create_proof((GoUint64)(rng() & 0x7fffffffu), &gi, &gc2, &gh, &gpnt, &gcl);
    *ms_proof += ms_since(t0);
    return r;
}

// Client side of one audit: one MSM over the client's own n_points MACs, 2 mult_point, 2 add_point,
// verify_proof, compare_commitment.
bool client_audit(const Inputs& in, const Ops& op, const AuditReply& r, uint64_t seed, double* ms) {
    std::mt19937_64 rng(seed ^ 0x9e3779b97f4a7c15ull);
    std::vector<Sc> sc(kAuditPoints);
    std::vector<Pt> pts(kAuditPoints);
    for (int j = 0; j < kAuditPoints; j++) {
        sc[j] = scalar_from_int((uint32_t)(rng() & 0x7fffffffu));
        pts[j] = in.complements[(size_t)(rng() % in.complements.size())];
    }
    auto t0 = Clock::now();
    Pt acc;
    op.multi_exp(acc, pts.data(), sc.data(), kAuditPoints);
    Pt a = r.combined_mac, b = r.combined_align;
    op.mult(a, sc[0]);
    op.mult(b, sc[1]);
    op.add(a, acc);
    op.add(b, acc);
    GoSlice gc = slice((void*)r.commitment.b, 64), gh = slice((void*)r.proof_h.b, 64), gp = slice((void*)r.point, 32),
            gl = slice((void*)r.claim, 32);
    const bool ok = verify_proof(&gc, &gh, &gp, &gl) != 0;
    GoSlice ga = slice(a.b, 64), ga2 = slice(a.b, 64);
    const bool same = compare_commitment(&ga, &ga2) != 0;
    *ms += ms_since(t0);
    return ok && same;
}

struct PassResult {
    std::string name;
    double update_total_ms = 0, crebuild_ms = 0, butterfly_ms = 0, align_ms = 0;
    double audit_msm_ms = 0, audit_agg_ms = 0, audit_proof_ms = 0, client_ms = 0;
    long butterflies = 0;
    int audits = 0, updates = 0;
    bool verified = true;
    std::vector<Pt> final_state;
    std::vector<AuditReply> replies;
};

PassResult run_pass(const Inputs& in, const Ops& op, int n_updates, int n_audits) {
    PassResult pr;
    pr.name = op.name;
    Server sv(in, op);
    auto t0 = Clock::now();
    for (int i = 0; i < n_updates; i++) sv.update(i % in.n_blocks);
    pr.update_total_ms = ms_since(t0);
    pr.updates = n_updates;
    pr.crebuild_ms = sv.t_crebuild_ms;
    pr.butterfly_ms = sv.t_butterfly_ms;
    pr.align_ms = sv.t_align_ms;
    pr.butterflies = sv.n_butterflies;
    for (int j = 0; j < n_audits; j++) {
        AuditReply r = server_audit(in, op, sv, 1000 + (uint64_t)j, &pr.audit_msm_ms, &pr.audit_agg_ms, &pr.audit_proof_ms);
        pr.verified = client_audit(in, op, r, 1000 + (uint64_t)j, &pr.client_ms) && pr.verified;
        pr.replies.push_back(r);
    }
    pr.audits = n_audits;
    const Level& T = sv.H[in.height - 1];
    pr.final_state = T.mac_x;
    pr.final_state.insert(pr.final_state.end(), T.mac_y.begin(), T.mac_y.end());
    pr.final_state.insert(pr.final_state.end(), T.al_x.begin(), T.al_x.end());
    pr.final_state.insert(pr.final_state.end(), T.al_y.begin(), T.al_y.end());
    return pr;
}

void print_pass(const PassResult& p, bool last) {
    const double upd = p.updates ? p.update_total_ms / p.updates : 0;
    const double upd_ex = p.updates > 1 ? (p.update_total_ms - p.crebuild_ms) / (p.updates - 1) : 0;
    printf("  \"%s\": {\"updates\": %d, \"update_ms_amortised\": %.4f, \"update_ms_amortised_excl_crebuild\": %.4f, "
           "\"crebuild_ms\": %.3f, \"hierarchy_butterfly_ms_total\": %.3f, \"align_mac_ms_total\": %.3f, \"butterflies\": %ld, "
           "\"audits\": %d, \"audit_ms\": %.4f, \"audit_server_msm_ms\": %.4f, \"audit_server_aggregate_align_ms\": %.4f, "
           "\"audit_server_proof_ms\": %.4f, \"audit_client_ms\": %.4f, \"proofs_verified\": %s}%s\n",
           p.name.c_str(), p.updates, upd, upd_ex, p.crebuild_ms, p.butterfly_ms, p.align_ms, p.butterflies, p.audits,
           p.audits ? (p.audit_msm_ms + p.audit_agg_ms + p.audit_proof_ms + p.client_ms) / p.audits : 0.0,
           p.audits ? p.audit_msm_ms / p.audits : 0.0, p.audits ? p.audit_agg_ms / p.audits : 0.0,
           p.audits ? p.audit_proof_ms / p.audits : 0.0, p.audits ? p.client_ms / p.audits : 0.0, p.verified ? "true" : "false",
           last ? "" : ",");
}

}  // namespace

int main(int argc, char** argv) {
    int n_blocks = 1024, n_audits = 100, cpu_updates = -1, cpu_audits = 5;
    const char* oracle_path = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--blocks") && i + 1 < argc) n_blocks = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--audits") && i + 1 < argc) n_audits = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--oracle") && i + 1 < argc) oracle_path = argv[++i];
        else if (!strcmp(argv[i], "--cpu-updates") && i + 1 < argc) cpu_updates = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--cpu-audits") && i + 1 < argc) cpu_audits = atoi(argv[++i]);
        else {
            fprintf(stderr, "usage: %s [--blocks N (power of two)] [--audits K] [--oracle liboracle_bn254.so] [--cpu-updates U] [--cpu-audits A]\n", argv[0]);
            return 2;
        }
    }
    if (n_blocks < 2 || (n_blocks & (n_blocks - 1))) return 2;
    Inputs in;
    in.n_blocks = n_blocks;
    in.height = 1;
    while ((1 << (in.height - 1)) < n_blocks) in.height++;   // height = ceil(log2 n) + 1 (Server.hpp:219)

    // ---- Client::initialize's library calls: keys, SRS (Client.hpp:159-167, 348-354)
    uint8_t tau[16] = {0xff, 0xee, 0xdd, 0xcc, 0xbb, 0xaa, 0x99, 0x88, 0x77, 0x66, 0x55, 0x44, 0x33, 0x22, 0x11, 0x00};   // TAU_KEY
    uint8_t alpha[32] = {0};
    const uint8_t secret[16] = {0x00, 0x11, 0x22, 0x33, 0x44, 0x55, 0x66, 0x77, 0x88, 0x99, 0xaa, 0xbb, 0xcc, 0xdd, 0xee, 0xff};
    memcpy(alpha + 16, secret, 16);   // Client.hpp:851-853
    GoSlice gt = slice(tau, 16), ga = slice(alpha, 32);
    init_key(&gt, &ga);
    std::vector<uint8_t> blob(kChunks * 32 + 132);
    GoSlice gblob = slice(blob.data(), (long long)blob.size());
    GoInt64 blob_len = 0;
    auto t_init = Clock::now();
    init_SRS(kChunks, &gblob, &blob_len);
    const double init_srs_ms = ms_since(t_init);

    Ops legacy = library_ops(false), batched = library_ops(true);
    std::mt19937_64 rng(20220615);
    // synthetic MACs: a chain P_i = P_0 + i Q built with the library's own single-point calls
    Pt g;
    memset(g.b, 0, 64);
    g.b[31] = 1;
    g.b[63] = 2;
    Pt q = g;
    legacy.mult(q, scalar_from_int(0x12345677u));
    in.mac_u.resize((size_t)n_blocks);
    Pt cur = g;
    legacy.mult(cur, scalar_from_int(0x0badcafeu));
    for (int i = 0; i < n_blocks; i++) {
        in.mac_u[(size_t)i] = cur;
        legacy.add(cur, q);
    }
    in.complements.resize(257);
    for (auto& c : in.complements) {
        c = cur;
        legacy.add(cur, q);
    }
    in.align_sc.resize((size_t)n_blocks);
    for (auto& v : in.align_sc) {
        v.resize(kChunks);
        for (auto& s : v)
            for (int k = 0; k < 32; k++) s.b[k] = (uint8_t)rng();
    }
    in.blocks.resize((size_t)n_blocks * kChunks * 64);
    for (size_t i = 0; i < in.blocks.size(); i++) in.blocks[i] = (uint8_t)rng();
    for (size_t i = 63; i < in.blocks.size(); i += 64) in.blocks[i] &= 0x07;   // chunks below 2^507 < LCM (utils.h:42)
    // w: an element of order 2 * n_blocks ... the reference takes a root of unity of the FFT length (Server.hpp:205-217);
    // PRIME_MODULUS - 1 = 207 * 2^248, so g^(207 * 2^248 / 2^k) has order dividing 2^k; pick the first base that gives full order
    {
        const int k = in.height;   // order 2^height = 2 * n_blocks covers every exponent used above
        for (uint32_t base = 3;; base += 2) {
            U256 b = {{base, 0, 0, 0, 0, 0, 0, 0}};
            U256 cand = powmod(b, 207, 248 - k);
            U256 chk = cand;
            for (int i = 0; i < k - 1; i++) chk = mulmod(chk, chk);   // cand^(2^(k-1)) must be -1, not 1
            U256 one = {{1, 0, 0, 0, 0, 0, 0, 0}};
            if (memcmp(chk.w, one.w, 32) != 0) {
                in.w = cand;
                break;
            }
        }
    }
    // SRS points for the CPU pass: [tau^i] G by repeated mult_point
    {
        Sc ts;
        memset(ts.b, 0, 32);
        memcpy(ts.b + 16, tau, 16);
        in.srs.resize(kChunks);
        Pt p = g;
        for (int i = 0; i < kChunks; i++) {
            in.srs[(size_t)i] = p;
            legacy.mult(p, ts);
        }
    }

    std::vector<PassResult> passes;
    passes.push_back(run_pass(in, legacy, n_blocks, n_audits));
    passes.push_back(run_pass(in, batched, n_blocks, n_audits));
    bool have_cpu = false;
    Ops cpu;
    if (oracle_path && cpu_ops(oracle_path, in.srs, &cpu)) {
        have_cpu = true;
        // the CPU pass may be bounded (bench.py's cpu_baseline leg): a prefix of the updates and a few audits
        passes.push_back(run_pass(in, cpu, cpu_updates < 0 ? n_blocks : cpu_updates, cpu_audits));
    }
    // ---- parity between the passes
    bool state_equal = passes[0].final_state == passes[1].final_state;
    bool audits_equal = passes[0].replies.size() == passes[1].replies.size();
    for (size_t j = 0; audits_equal && j < passes[0].replies.size(); j++) {
        const AuditReply &a = passes[0].replies[j], &b = passes[1].replies[j];
        audits_equal = a.combined_mac == b.combined_mac && a.combined_align == b.combined_align && a.commitment == b.commitment &&
                       a.proof_h == b.proof_h && a.b_mod == b.b_mod && !memcmp(a.claim, b.claim, 32);
    }
    bool cpu_equal = true;
    if (have_cpu && passes[2].updates == n_blocks) {
        cpu_equal = passes[2].final_state == passes[0].final_state;
        if (!cpu_equal) {
            size_t bad = 0, first = 0;
            for (size_t i = 0; i < passes[0].final_state.size(); i++)
                if (!(passes[0].final_state[i] == passes[2].final_state[i])) {
                    if (!bad) first = i;
                    bad++;
                }
            fprintf(stderr, "cpu pass: %zu of %zu state entries differ, first at %zu\n", bad, passes[0].final_state.size(), first);
        }
        for (size_t j = 0; cpu_equal && j < passes[2].replies.size() && j < passes[0].replies.size(); j++)
        {
            const bool m = passes[2].replies[j].combined_mac == passes[0].replies[j].combined_mac;
            const bool a = passes[2].replies[j].combined_align == passes[0].replies[j].combined_align;
            if (!m || !a) fprintf(stderr, "cpu pass: audit %zu differs (combined_mac %d, combined_align %d)\n", j, (int)m, (int)a);
            cpu_equal = m && a;
        }
    }
    printf("{\n  \"workload\": \"Porla KZG mode, %d data blocks, NUM_CHUNKS = 128, TOP_CACHING_LEVEL = 10: %d updates (the last one "
           "rebuilds C) then %d audits; MAC-side C-ABI call census of SURVEY.md Appendix C\",\n",
           n_blocks, n_blocks, n_audits);
#ifdef PORLA_USE_REFERENCE_HEADER
    printf("  \"header\": \"reference libmultiexp.h\",\n");
#else
    printf("  \"header\": \"include/porla_multiexp.h\",\n");
#endif
    printf("  \"init_srs_ms\": %.3f,\n", init_srs_ms);
    for (size_t i = 0; i < passes.size(); i++) print_pass(passes[i], false);
    printf("  \"legacy_equals_batched_state\": %s, \"legacy_equals_batched_audits\": %s, \"cpu_equals_legacy\": %s\n}\n",
           state_equal ? "true" : "false", audits_equal ? "true" : "false",
           !have_cpu ? "null" : (passes[2].updates == n_blocks ? (cpu_equal ? "true" : "false") : "\"bounded pass: not compared\""));
    return state_equal && audits_equal && cpu_equal ? 0 : 1;
}
