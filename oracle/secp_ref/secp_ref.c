/* secp_ref.c -- thin C entry points over the UNMODIFIED vendored secp256k1 of the reference
 * (/root/reference/porla/Utils/secp256k1_lib), compiled from where the sources lie into
 * oracle/_ref/libsecp_ref.so.  No reference source is copied into this repository.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (parity of the CUDA secp256k1 path), by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * The include order follows secp256k1.c:12-23 (minus modules the path does not touch).
 */
#include "libsecp256k1-config.h"
#include "../include/secp256k1.h"
#include "assumptions.h"
#include "util.h"
#include "field_impl.h"
#include "scalar_impl.h"
#include "group_impl.h"
#include "ecmult_impl.h"
#include "ecmult_const_impl.h"
#include "eckey_impl.h"
#include "hash_impl.h"
#include "scratch_impl.h"

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void ref_default_error(const char* text, void* data) {
    (void)data;
    fprintf(stderr, "[secp_ref] internal error: %s\n", text);
    abort();
}
static const secp256k1_callback ref_error_cb = {ref_default_error, NULL};

typedef struct {
    secp256k1_scalar* sc;
    secp256k1_ge* pt;
} ref_cb_data;

/* same shape as ecmult_multi_callback, /root/reference/porla/Utils/utils.h:166-171 */
static int ref_cb(secp256k1_scalar* sc, secp256k1_ge* pt, size_t idx, void* cbdata) {
    ref_cb_data* d = (ref_cb_data*)cbdata;
    *sc = d->sc[idx];
    *pt = d->pt[idx];
    return 1;
}

size_t ref_secp_sizeof_ge(void) { return sizeof(secp256k1_ge); }
size_t ref_secp_sizeof_gej(void) { return sizeof(secp256k1_gej); }
size_t ref_secp_sizeof_scalar(void) { return sizeof(secp256k1_scalar); }

/* external formats -> reference structs.  scalars: 32 B little-endian limbs exactly as
 * convert_ZZ_to_scalar writes them (utils.h:180-192, NOT reduced); points: X||Y big-endian,
 * 64 zero bytes = infinity. */
static void load_inputs(const uint8_t* scalars, const uint8_t* points, size_t n, secp256k1_scalar* sc, secp256k1_ge* pt) {
    size_t i;
    for (i = 0; i < n; i++) {
        secp256k1_fe x, y;
        int allzero = 1, k;
        memcpy(sc[i].d, scalars + 32 * i, 32);
        for (k = 0; k < 64; k++) if (points[64 * i + k]) { allzero = 0; break; }
        if (allzero) {
            secp256k1_ge_set_infinity(&pt[i]);
        } else {
            secp256k1_fe_set_b32(&x, points + 64 * i);
            secp256k1_fe_set_b32(&y, points + 64 * i + 32);
            secp256k1_ge_set_xy(&pt[i], &x, &y);
        }
    }
}

static void store_result(const secp256k1_gej* r, uint8_t* out64, uint8_t* out33) {
    secp256k1_gej t = *r;
    secp256k1_ge a;
    if (secp256k1_gej_is_infinity(&t)) {
        if (out64) memset(out64, 0, 64);
        if (out33) memset(out33, 0, 33);
        return;
    }
    secp256k1_ge_set_gej(&a, &t);
    if (out64) {
        secp256k1_fe_normalize_var(&a.x);
        secp256k1_fe_normalize_var(&a.y);
        secp256k1_fe_get_b32(out64, &a.x);
        secp256k1_fe_get_b32(out64 + 32, &a.y);
    }
    if (out33) {
        size_t sz = 33;
        secp256k1_eckey_pubkey_serialize(&a, out33, &sz, 1); /* eckey_impl.h:36-52 */
    }
}

/* scratch sized the way Porla sizes it (Client.hpp:119-123, :755-758) */
static secp256k1_scratch* make_scratch(size_t n) {
    int bucket_window = secp256k1_pippenger_bucket_window(n ? n : 1);
    size_t sz = secp256k1_pippenger_scratch_size(n ? n : 1, bucket_window) + PIPPENGER_SCRATCH_OBJECTS * ALIGNMENT;
    return secp256k1_scratch_create(&ref_error_cb, sz);
}

/* r = sum sc_i * pt_i through secp256k1_ecmult_multi_var (ecmult_impl.h:814), zero G scalar as
 * Porla passes it (&szero).  Returns the function's return value. */
int ref_secp_msm(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out64, uint8_t* out33) {
    secp256k1_scalar* sc = (secp256k1_scalar*)malloc((n ? n : 1) * sizeof(secp256k1_scalar));
    secp256k1_ge* pt = (secp256k1_ge*)malloc((n ? n : 1) * sizeof(secp256k1_ge));
    secp256k1_scalar szero;
    secp256k1_gej r;
    secp256k1_scratch* scratch = make_scratch(n);
    ref_cb_data d;
    int ok;
    load_inputs(scalars, points, n, sc, pt);
    secp256k1_scalar_set_int(&szero, 0);
    d.sc = sc;
    d.pt = pt;
    ok = secp256k1_ecmult_multi_var(&ref_error_cb, scratch, &r, &szero, ref_cb, &d, n);
    store_result(&r, out64, out33);
    secp256k1_scratch_destroy(&ref_error_cb, scratch);
    free(sc);
    free(pt);
    return ok;
}

/* The reference's own parallel shape: contiguous range partition over `nthreads` workers, one
 * scratch each, partial gej sums added serially (Client.hpp:747-787). */
typedef struct {
    ref_cb_data d;
    size_t n;
    secp256k1_gej r;
    int ok;
} ref_part;

static void* ref_part_run(void* arg) {
    ref_part* p = (ref_part*)arg;
    secp256k1_scalar szero;
    secp256k1_scratch* scratch = make_scratch(p->n);
    secp256k1_scalar_set_int(&szero, 0);
    p->ok = secp256k1_ecmult_multi_var(&ref_error_cb, scratch, &p->r, &szero, ref_cb, &p->d, p->n);
    secp256k1_scratch_destroy(&ref_error_cb, scratch);
    return NULL;
}

typedef struct {
    secp256k1_scalar* sc;
    secp256k1_ge* pt;
    size_t n;
} ref_prepared;

/* parse once so that timing loops measure ecmult_multi only */
void* ref_secp_prepare(const uint8_t* scalars, const uint8_t* points, size_t n) {
    ref_prepared* p = (ref_prepared*)malloc(sizeof(ref_prepared));
    p->sc = (secp256k1_scalar*)malloc((n ? n : 1) * sizeof(secp256k1_scalar));
    p->pt = (secp256k1_ge*)malloc((n ? n : 1) * sizeof(secp256k1_ge));
    p->n = n;
    load_inputs(scalars, points, n, p->sc, p->pt);
    return p;
}
void ref_secp_release(void* h) {
    ref_prepared* p = (ref_prepared*)h;
    free(p->sc);
    free(p->pt);
    free(p);
}
int ref_secp_msm_prepared(void* h, size_t n, int nthreads, uint8_t* out64, uint8_t* out33) {
    ref_prepared* p = (ref_prepared*)h;
    ref_part* parts;
    pthread_t* th;
    secp256k1_gej total;
    size_t each, start = 0;
    int t, ok = 1;
    if (n > p->n) n = p->n;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    parts = (ref_part*)malloc((size_t)nthreads * sizeof(ref_part));
    th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    each = n / (size_t)nthreads;
    for (t = 0; t < nthreads; t++) {
        size_t cnt = t == nthreads - 1 ? n - each * (size_t)t : each;
        parts[t].d.sc = p->sc + start;
        parts[t].d.pt = p->pt + start;
        parts[t].n = cnt;
        start += cnt;
        if (nthreads > 1) pthread_create(&th[t], NULL, ref_part_run, &parts[t]);
        else ref_part_run(&parts[t]);
    }
    secp256k1_gej_set_infinity(&total);
    for (t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        ok &= parts[t].ok;
        secp256k1_gej_add_var(&total, &total, &parts[t].r, NULL);
    }
    store_result(&total, out64, out33);
    free(parts);
    free(th);
    return ok;
}

/* x*G through ecmult_multi_var with ONE point (tests.c:4695), serialised as tests.c:4702-4711 does:
 * 65-byte uncompressed, or a single 0x00 for infinity.  Returns the number of bytes written. */
typedef struct { const secp256k1_scalar* x; } ref_acc_data;
static int ref_acc_cb(secp256k1_scalar* sc, secp256k1_ge* pt, size_t idx, void* data) {
    (void)idx;
    *sc = *((ref_acc_data*)data)->x;
    *pt = secp256k1_ge_const_g;
    return 1;
}
/* Known-answer test of the reference: tests.c:4715-4757 (test_ecmult_constants), computed through
 * ecmult_multi_var only.  out32 receives the SHA-256; expected e4711b4d...859ab7b4 (tests.c:4732-4737).
 * If scalars_out != NULL it also receives the 32842 scalars (32 B little-endian limbs each). */
size_t ref_secp_kat(uint8_t* out32, uint8_t* scalars_out) {
    secp256k1_scalar x, zero;
    secp256k1_sha256 acc;
    secp256k1_scratch* scratch = secp256k1_scratch_create(&ref_error_cb, 65536);
    size_t count = 0;
    int i, j, pass;
    secp256k1_scalar_set_int(&zero, 0);
    secp256k1_sha256_initialize(&acc);
    for (pass = 0; pass < 2; pass++) {
        int limit = pass == 0 ? 36 : 255;
        for (i = 0; i <= limit; ++i) {
            int jmax = pass == 0 ? 2 : 256;
            for (j = (pass == 0 ? 0 : 1); j < jmax; j += (pass == 0 ? 1 : 2)) {
                secp256k1_gej rj;
                secp256k1_ge r;
                ref_acc_data d;
                if (pass == 0) {
                    secp256k1_scalar_set_int(&x, (unsigned)i);
                    if (j == 1) secp256k1_scalar_negate(&x, &x);
                } else {
                    int k;
                    secp256k1_scalar_set_int(&x, (unsigned)j);
                    for (k = 0; k < i; ++k) secp256k1_scalar_add(&x, &x, &x);
                }
                d.x = &x;
                if (scalars_out) memcpy(scalars_out + 32 * count, x.d, 32);
                count++;
                secp256k1_ecmult_multi_var(&ref_error_cb, scratch, &rj, &zero, ref_acc_cb, &d, 1);
                if (secp256k1_gej_is_infinity(&rj)) {
                    const unsigned char zerobyte[1] = {0};
                    secp256k1_sha256_write(&acc, zerobyte, 1);
                } else {
                    unsigned char bytes[65];
                    size_t size = 65;
                    secp256k1_ge_set_gej_var(&r, &rj);
                    secp256k1_eckey_pubkey_serialize(&r, bytes, &size, 0);
                    secp256k1_sha256_write(&acc, bytes, size);
                }
            }
        }
    }
    secp256k1_sha256_finalize(&acc, out32);
    secp256k1_scratch_destroy(&ref_error_cb, scratch);
    return count;
}

/* Synthetic points as in BASELINE.md section 2: P_0 = G, P_{i+1} = P_i + Q with Q = q*G,
 * batch-normalised (group_impl.h:122).  out: n x 64 B big-endian X||Y. */
void ref_secp_point_chain(const uint8_t* q_le32, size_t n, uint8_t* out) {
    secp256k1_scalar q, zero;
    secp256k1_gej qj, cur, gj;
    secp256k1_ge qa;
    secp256k1_gej* js = (secp256k1_gej*)malloc((n ? n : 1) * sizeof(secp256k1_gej));
    secp256k1_ge* as = (secp256k1_ge*)malloc((n ? n : 1) * sizeof(secp256k1_ge));
    size_t i;
    memcpy(q.d, q_le32, 32);
    secp256k1_scalar_set_int(&zero, 0);
    secp256k1_gej_set_ge(&gj, &secp256k1_ge_const_g);
    secp256k1_ecmult(&qj, &gj, &q, &zero);
    secp256k1_ge_set_gej(&qa, &qj);
    cur = gj;
    for (i = 0; i < n; i++) {
        js[i] = cur;
        secp256k1_gej_add_ge_var(&cur, &cur, &qa, NULL);
    }
    secp256k1_ge_set_all_gej_var(as, js, n);
    for (i = 0; i < n; i++) {
        secp256k1_fe_normalize_var(&as[i].x);
        secp256k1_fe_normalize_var(&as[i].y);
        secp256k1_fe_get_b32(out + 64 * i, &as[i].x);
        secp256k1_fe_get_b32(out + 64 * i + 32, &as[i].y);
    }
    free(js);
    free(as);
}

/* The reference's SHA-256 object as inner_product_prove / inner_product_verify drive it (Server.hpp:2306-2310, 2386-2387,
 * 2429-2430): ONE secp256k1_sha256 that is written to and finalized again and again without re-initialisation.
 * finalize() zeroes the state words but keeps the byte counter, so the digests are not plain SHA-256 of anything.
 * data = the concatenated segments, lens[i] = length of segment i; after each segment the object is finalized and the
 * 32-byte digest stored at outs + 32 i. */
void ref_sha256_sequence(const uint8_t* data, const uint32_t* lens, int nseg, uint8_t* outs) {
    secp256k1_sha256 h;
    int i;
    secp256k1_sha256_initialize(&h);
    for (i = 0; i < nseg; i++) {
        secp256k1_sha256_write(&h, data, lens[i]);
        data += lens[i];
        secp256k1_sha256_finalize(&h, outs + 32 * i);
    }
}
