/* bn254_oracle.c -- CPU restatement of the BN254 G1 path behind Porla's libmultiexp.so.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under porla_b200/ links, loads or calls this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in gnark-crypto v0.6.0 (Go module pinned by
 * /root/reference/auto_setup.sh:44-51 and README.md:16), which is NOT under /root/reference and
 * cannot be built here (no Go toolchain).  This file restates the published algorithm that
 * /root/reference/porla/main.go reaches:
 *   - fr.Element.SetBytes / fp.Element.SetBytes : big-endian, reduced mod r / p      (main.go:127,130)
 *   - G1Affine.Unmarshal / Marshal               : 64-byte X||Y big-endian, inf = 0  (main.go:130,137)
 *   - G1Affine.MultiExp                          : 4x64 Montgomery field, signed c-bit windows,
 *       2^(c-1) extended-Jacobian (XYZZ) buckets per window, mixed additions, running-sum
 *       bucket reduction, c doublings between windows, windows processed by parallel workers
 *                                                               (main.go:134-136)
 *   - G1Affine.Add / ScalarMultiplication / Neg                                      (main.go:196-222)
 * It is checked against the independent big-integer oracle oracle/curves_py.py and the public
 * vectors of SURVEY.md 8(c) in tests/test_oracle.py.  Every result is a canonical affine point,
 * so any correct MSM is bit-exact after Marshal.
 *
 * Build: make -C oracle   (gcc -O3 -march=native -shared -fPIC -pthread)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;            /* Montgomery form, canonical (< p) */
typedef struct { fe x, y; } g1a;                  /* affine; infinity = (0,0)            */
typedef struct { fe x, y, zz, zzz; } g1x;         /* extended Jacobian; infinity: zz = 0 */

static const uint64_t P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
static const uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
static const uint64_t ORDER[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t INV = 0x87d20782e4866389ull;

/* ------------------------------------------------------------------------------------ field */
static inline int ge4(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - (uint64_t)br;
        r[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
    return (uint64_t)br;
}
static inline uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, sizeof(fe)) == 0; }
static inline void fe_add(fe* r, const fe* a, const fe* b) {
    uint64_t c = add4(r->l, a->l, b->l);
    if (c || ge4(r->l, P)) sub4(r->l, r->l, P);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b) {
    if (sub4(r->l, a->l, b->l)) add4(r->l, r->l, P);
}
static inline void fe_neg(fe* r, const fe* a) {
    if (fe_is_zero(a)) *r = *a;
    else sub4(r->l, P, a->l);
}
/* CIOS Montgomery product; p < 2^254 so the running value fits 4 limbs + 1 */
static inline void fe_mul(fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    memcpy(r->l, t, 32);
    if (t[4] || ge4(r->l, P)) sub4(r->l, r->l, P);
}
static inline void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static inline void fe_dbl(fe* r, const fe* a) { fe_add(r, a, a); }
static void fe_inv(fe* r, const fe* a) { /* a^(p-2) */
    uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
    fe acc;
    memcpy(acc.l, ONE, 32);
    for (int i = 255; i >= 0; i--) {
        fe_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a);
    }
    *r = acc;
}
/* big-endian 32 bytes -> Montgomery element, reduced mod p (fp.Element.SetBytes) */
static void fe_from_be(fe* r, const uint8_t* b, int mask_flags) {
    fe t;
    for (int i = 0; i < 4; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (3 - i) + j];
        t.l[i] = v;
    }
    if (mask_flags) t.l[3] &= 0x3fffffffffffffffull;
    while (ge4(t.l, P)) sub4(t.l, t.l, P);
    fe r2;
    memcpy(r2.l, R2, 32);
    fe_mul(r, &t, &r2);
}
static void fe_to_be(uint8_t* b, const fe* a) {
    fe one = {{1, 0, 0, 0}}, t;
    fe_mul(&t, a, &one);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (uint8_t)(t.l[i] >> (8 * (7 - j)));
}

/* ------------------------------------------------------------------------------------ group */
static inline int g1x_is_inf(const g1x* p) { return fe_is_zero(&p->zz); }
static inline int g1a_is_inf(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void g1x_set_inf(g1x* p) { memset(p, 0, sizeof(*p)); }

static void g1x_dbl_affine(g1x* r, const g1a* p) { /* mdbl-2008-s-1, a = 0 */
    fe u, v, w, s, xx, m, t;
    fe_dbl(&u, &p->y);
    fe_sqr(&v, &u);
    fe_mul(&w, &u, &v);
    fe_mul(&s, &p->x, &v);
    fe_sqr(&xx, &p->x);
    fe_dbl(&m, &xx);
    fe_add(&m, &m, &xx);
    fe_sqr(&r->x, &m);
    fe_dbl(&t, &s);
    fe_sub(&r->x, &r->x, &t);
    fe_sub(&t, &s, &r->x);
    fe_mul(&t, &m, &t);
    fe_mul(&u, &w, &p->y);
    fe_sub(&r->y, &t, &u);
    r->zz = v;
    r->zzz = w;
}
static void g1x_dbl(g1x* r, const g1x* p) { /* dbl-2008-s-1, a = 0 */
    if (g1x_is_inf(p)) { *r = *p; return; }
    fe u, v, w, s, xx, m, t, x3, y3;
    fe_dbl(&u, &p->y);
    fe_sqr(&v, &u);
    fe_mul(&w, &u, &v);
    fe_mul(&s, &p->x, &v);
    fe_sqr(&xx, &p->x);
    fe_dbl(&m, &xx);
    fe_add(&m, &m, &xx);
    fe_sqr(&x3, &m);
    fe_dbl(&t, &s);
    fe_sub(&x3, &x3, &t);
    fe_sub(&t, &s, &x3);
    fe_mul(&t, &m, &t);
    fe_mul(&u, &w, &p->y);
    fe_sub(&y3, &t, &u);
    fe_mul(&r->zz, &v, &p->zz);
    fe_mul(&r->zzz, &w, &p->zzz);
    r->x = x3;
    r->y = y3;
}
/* r += q (affine), madd-2008-s; neg != 0 adds -q */
static void g1x_madd(g1x* r, const g1a* q, int neg) {
    if (g1a_is_inf(q)) return;
    g1a qq = *q;
    if (neg) fe_neg(&qq.y, &q->y);
    if (g1x_is_inf(r)) {
        r->x = qq.x;
        r->y = qq.y;
        memcpy(r->zz.l, ONE, 32);
        memcpy(r->zzz.l, ONE, 32);
        return;
    }
    fe p, rr, pp, ppp, q2, t, x3;
    fe_mul(&p, &qq.x, &r->zz);
    fe_sub(&p, &p, &r->x);
    fe_mul(&rr, &qq.y, &r->zzz);
    fe_sub(&rr, &rr, &r->y);
    if (fe_is_zero(&p)) {
        if (fe_is_zero(&rr)) g1x_dbl_affine(r, &qq);
        else g1x_set_inf(r);
        return;
    }
    fe_sqr(&pp, &p);
    fe_mul(&ppp, &p, &pp);
    fe_mul(&q2, &r->x, &pp);
    fe_sqr(&x3, &rr);
    fe_sub(&x3, &x3, &ppp);
    fe_dbl(&t, &q2);
    fe_sub(&x3, &x3, &t);
    fe_sub(&t, &q2, &x3);
    fe_mul(&t, &rr, &t);
    fe_mul(&q2, &r->y, &ppp);
    fe_sub(&r->y, &t, &q2);
    r->x = x3;
    fe_mul(&r->zz, &r->zz, &pp);
    fe_mul(&r->zzz, &r->zzz, &ppp);
}
/* r += o, add-2008-s */
static void g1x_add(g1x* r, const g1x* o) {
    if (g1x_is_inf(o)) return;
    if (g1x_is_inf(r)) { *r = *o; return; }
    fe u1, u2, s1, s2, p, rr, pp, ppp, q, t, x3;
    fe_mul(&u1, &r->x, &o->zz);
    fe_mul(&u2, &o->x, &r->zz);
    fe_mul(&s1, &r->y, &o->zzz);
    fe_mul(&s2, &o->y, &r->zzz);
    fe_sub(&p, &u2, &u1);
    fe_sub(&rr, &s2, &s1);
    if (fe_is_zero(&p)) {
        if (fe_is_zero(&rr)) { g1x d; g1x_dbl(&d, r); *r = d; }
        else g1x_set_inf(r);
        return;
    }
    fe_sqr(&pp, &p);
    fe_mul(&ppp, &p, &pp);
    fe_mul(&q, &u1, &pp);
    fe_sqr(&x3, &rr);
    fe_sub(&x3, &x3, &ppp);
    fe_dbl(&t, &q);
    fe_sub(&x3, &x3, &t);
    fe_sub(&t, &q, &x3);
    fe_mul(&t, &rr, &t);
    fe_mul(&q, &s1, &ppp);
    fe_sub(&r->y, &t, &q);
    r->x = x3;
    fe_mul(&t, &r->zz, &o->zz);
    fe_mul(&r->zz, &t, &pp);
    fe_mul(&t, &r->zzz, &o->zzz);
    fe_mul(&r->zzz, &t, &ppp);
}
static void g1x_to_affine(g1a* r, const g1x* p) {
    if (g1x_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
    fe i3, t;
    fe_inv(&i3, &p->zzz);
    fe_mul(&t, &p->zz, &i3);
    fe_sqr(&t, &t);
    fe_mul(&r->x, &p->x, &t);
    fe_mul(&r->y, &p->y, &i3);
}
static void g1a_from_bytes(g1a* r, const uint8_t* b) { /* uncompressed form only (what Porla passes) */
    fe_from_be(&r->x, b, 1);
    fe_from_be(&r->y, b + 32, 0);
}
static void g1a_to_bytes(uint8_t* b, const g1a* p) {
    fe_to_be(b, &p->x);
    fe_to_be(b + 32, &p->y);
}
/* big-endian 32 bytes -> canonical scalar limbs mod r (fr.Element.SetBytes then ToRegular) */
static void scalar_from_be(uint64_t* s, const uint8_t* b) {
    for (int i = 0; i < 4; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (3 - i) + j];
        s[i] = v;
    }
    while (ge4(s, ORDER)) sub4(s, s, ORDER);
}

/* ------------------------------------------------------------------------------------ MSM */
typedef struct {
    const g1a* pts;
    const int32_t* digits; /* [nwin][n] signed digits */
    size_t n;
    int c, nwin;
    int next_window;       /* work queue */
    pthread_mutex_t mu;
    g1x* wsum;             /* per-window sums */
} msm_job;

static void window_sum(const msm_job* j, int w, g1x* out) {
    size_t nb = (size_t)1 << (j->c - 1);
    g1x* buckets = (g1x*)calloc(nb, sizeof(g1x));
    const int32_t* d = j->digits + (size_t)w * j->n;
    for (size_t i = 0; i < j->n; i++) {
        int32_t v = d[i];
        if (v > 0) g1x_madd(&buckets[v - 1], &j->pts[i], 0);
        else if (v < 0) g1x_madd(&buckets[-v - 1], &j->pts[i], 1);
    }
    g1x run, sum;
    g1x_set_inf(&run);
    g1x_set_inf(&sum);
    for (size_t k = nb; k-- > 0;) {
        g1x_add(&run, &buckets[k]);
        g1x_add(&sum, &run);
    }
    free(buckets);
    *out = sum;
}
static void* msm_worker(void* arg) {
    msm_job* j = (msm_job*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int w = j->next_window++;
        pthread_mutex_unlock(&j->mu);
        if (w >= j->nwin) break;
        window_sum(j, w, &j->wsum[w]);
    }
    return NULL;
}
static int pick_window(size_t n) { /* minimise nwin * (n + 2^c) */
    int best = 4;
    double bc = 1e300;
    for (int c = 2; c <= 20; c++) {
        int nwin = (254 + 1 + c - 1) / c;
        double cost = (double)nwin * ((double)n + 2.0 * (double)((size_t)1 << (c - 1)));
        if (cost < bc) { bc = cost; best = c; }
    }
    return best;
}

/* sum_i s_i * P_i.  scalars: n x 32 B big-endian (any value < 2^256), points: n x 64 B Marshal
 * layout, out: 64 B.  nthreads <= 0 selects 1. */
void oracle_bn254_msm(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out, int nthreads) {
    if (n == 0) { memset(out, 0, 64); return; }
    int c = pick_window(n);
    int nwin = (254 + 1 + c - 1) / c;
    g1a* pts = (g1a*)malloc(n * sizeof(g1a));
    int32_t* digits = (int32_t*)malloc((size_t)nwin * n * sizeof(int32_t));
    for (size_t i = 0; i < n; i++) {
        g1a_from_bytes(&pts[i], points + 64 * i);
        uint64_t s[5];
        scalar_from_be(s, scalars + 32 * i);
        s[4] = 0;
        int carry = 0;
        for (int w = 0; w < nwin; w++) {
            int pos = w * c, word = pos >> 6, sh = pos & 63;
            uint64_t v = word < 4 ? s[word] >> sh : 0;
            if (sh && word < 3) v |= s[word + 1] << (64 - sh);
            int64_t dgt = (int64_t)(v & (((uint64_t)1 << c) - 1)) + carry;
            if (dgt > ((int64_t)1 << (c - 1))) { dgt -= (int64_t)1 << c; carry = 1; } else carry = 0;
            digits[(size_t)w * n + i] = (int32_t)dgt;
        }
    }
    msm_job job;
    job.pts = pts; job.digits = digits; job.n = n; job.c = c; job.nwin = nwin; job.next_window = 0;
    job.wsum = (g1x*)malloc((size_t)nwin * sizeof(g1x));
    pthread_mutex_init(&job.mu, NULL);
    if (nthreads <= 1) {
        msm_worker(&job);
    } else {
        if (nthreads > nwin) nthreads = nwin;
        pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, msm_worker, &job);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    g1x acc;
    g1x_set_inf(&acc);
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) { g1x d; g1x_dbl(&d, &acc); acc = d; }
        g1x_add(&acc, &job.wsum[w]);
    }
    g1a res;
    g1x_to_affine(&res, &acc);
    g1a_to_bytes(out, &res);
    pthread_mutex_destroy(&job.mu);
    free(job.wsum); free(digits); free(pts);
}

/* a + b on 64-byte buffers (G1Affine.Add, main.go:196-203) */
void oracle_bn254_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    g1a pa, pb, r;
    g1a_from_bytes(&pa, a);
    g1a_from_bytes(&pb, b);
    g1x acc;
    g1x_set_inf(&acc);
    g1x_madd(&acc, &pa, 0);
    g1x_madd(&acc, &pb, 0);
    g1x_to_affine(&r, &acc);
    g1a_to_bytes(out, &r);
}
/* k * p, k 32 B big-endian reduced mod r (ScalarMultiplication, main.go:205-215) */
void oracle_bn254_mul(const uint8_t* p, const uint8_t* k, uint8_t* out) {
    g1a pa, r;
    g1a_from_bytes(&pa, p);
    uint64_t s[4];
    scalar_from_be(s, k);
    g1x acc;
    g1x_set_inf(&acc);
    for (int i = 255; i >= 0; i--) {
        g1x d;
        g1x_dbl(&d, &acc);
        acc = d;
        if ((s[i >> 6] >> (i & 63)) & 1) g1x_madd(&acc, &pa, 0);
    }
    g1x_to_affine(&r, &acc);
    g1a_to_bytes(out, &r);
}
/* deterministic synthetic points for the CPU baseline: P_0 = base, P_{i+1} = P_i + step (affine,
 * 64-byte Marshal each); cheap on one core, on the curve by construction */
void oracle_bn254_point_chain(const uint8_t* base, const uint8_t* step, size_t n, uint8_t* out) {
    enum { BLK = 1024 };
    g1a b, s, cur;
    g1a_from_bytes(&b, base);
    g1a_from_bytes(&s, step);
    g1x acc;
    g1x_set_inf(&acc);
    g1x_madd(&acc, &b, 0);
    g1x* blk = (g1x*)malloc(BLK * sizeof(g1x));
    fe* pre = (fe*)malloc(BLK * sizeof(fe));
    for (size_t i0 = 0; i0 < n; i0 += BLK) {
        size_t m = n - i0 < BLK ? n - i0 : BLK;
        for (size_t k = 0; k < m; k++) {   /* no point of the chain is infinity for sane inputs */
            blk[k] = acc;
            g1x_madd(&acc, &s, 0);
        }
        /* Montgomery batch inversion of the zzz coordinates */
        fe run;
        memcpy(run.l, ONE, 32);
        for (size_t k = 0; k < m; k++) {
            pre[k] = run;
            fe_mul(&run, &run, &blk[k].zzz);
        }
        fe inv;
        fe_inv(&inv, &run);
        for (size_t k = m; k-- > 0;) {
            fe i3, t;
            fe_mul(&i3, &inv, &pre[k]);
            fe_mul(&inv, &inv, &blk[k].zzz);
            fe_mul(&t, &blk[k].zz, &i3);
            fe_sqr(&t, &t);
            fe_mul(&cur.x, &blk[k].x, &t);
            fe_mul(&cur.y, &blk[k].y, &i3);
            g1a_to_bytes(out + 64 * (i0 + k), &cur);
        }
    }
    free(pre);
    free(blk);
}
