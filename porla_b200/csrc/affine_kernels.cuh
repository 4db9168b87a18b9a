// Bucket accumulation in AFFINE coordinates with shared inversions (the "batched affine" variant of k_accumulate).
//
// A mixed addition into an XYZZ bucket costs 8M + 2S = 10 field products (ec.cuh, madd-2008-s); an affine + affine addition
// costs one inversion plus 2M + 1S, and Montgomery's trick turns k inversions into one inversion plus 3(k - 1) products:
// 6 products per addition when the inversion is shared by many.  The inversion itself is the binary GCD of fp_inv.cuh,
// which runs on the integer ALU, not on the multiplier pipe that bounds the kernel.
//
// Same slice structure as k_accumulate (msm_kernels.cuh): thread t owns the pairs [t L, (t+1) L) of the bucket-sorted list,
// L = 64.  Instead of a running sum per bucket the thread reduces its slice as a TREE, so that the additions of one round
// are independent and can share an inversion, without any communication between threads:
//   round 0   neighbours (2i, 2i+1) with the same bucket id are added:  table points -> p1[] (<= 44 entries, local memory)
//   round 1   the same on p1[] -> p2[] (<= 24 entries)
//   tail      what is left of every bucket (about an eighth of its points) is summed with XYZZ mixed additions
// For 61 additions of a typical slice: 46 at 6 products + 15 at 10 products + 2 inversions, against 610 products.
// Exceptional cases are exact: P + P uses the tangent (denominator 2y), P - P yields infinity (tracked in a bit mask, never
// encoded in coordinates), infinity + P copies P.  Results leave the kernel exactly as k_accumulate's do -- buckets[],
// part_head[], part_tail[] as XYZZ records -- so the stitching and reduction kernels are unchanged.  A slice whose keys
// are mostly distinct (sparse buckets) exceeds p1[] / p2[] and simply skips the rounds it cannot run.
#pragma once
#include "msm_kernels.cuh"

namespace porla {

constexpr int kAffL = 64;          // pairs per slice
constexpr int kAffCap1 = 44;       // outputs of round 0 a thread can hold
constexpr int kAffCap2 = 24;       // outputs of round 1
constexpr int kAffMaxAdds = 32;    // additions per round and thread (L / 2)
constexpr int kAffThreads = 128;
constexpr int kAffDefaultRounds = 0;   // rounds used when PORLA_ACC_AFFINE is unset and the MSM is large (0: XYZZ kernel)
#ifndef PORLA_AFF_MIN_BLOCKS
#define PORLA_AFF_MIN_BLOCKS 3
#endif

enum : uint32_t { kOpCopy = 0, kOpAdd = 1, kOpDbl = 2, kOpInf = 3 };

// Where a round reads its inputs: the bucket-sorted pair list + point table (round 0), or a local array (later rounds).
template <class C>
struct AffSrc {
    using F = typename C::F;
    const Affine<F>* points;
    const F* phi_x;
    uint32_t phi_off;
    const uint2* pairs;        // non-null: input i is the table point pairs[i].y, its bucket pairs[i].x
    const Affine<F>* arr;      // else: arr[i], keys[i]
    const uint32_t* keys;

    PORLA_D uint32_t key(int i) const { return pairs ? __ldg(&pairs[i].x) : keys[i]; }
    PORLA_D Affine<F> pt(int i) const {
        if (pairs) return load_signed_point<C>(points, phi_x, phi_off, __ldg(&pairs[i].y));
        return arr[i];
    }
    PORLA_D F ptx(int i) const {
        if (pairs) {
            const uint32_t idx = __ldg(&pairs[i].y) & 0x7fffffffu;
            const bool image = C::kGlv && phi_off != 0 && idx >= phi_off;
            const uint32_t j = image ? idx - phi_off : idx;
            const uint4* sx = image ? reinterpret_cast<const uint4*>(phi_x + j) : reinterpret_cast<const uint4*>(points + j);
            F x;
            uint4* d = reinterpret_cast<uint4*>(&x);
            d[0] = __ldg(sx);
            d[1] = __ldg(sx + 1);
            return x;
        }
        return arr[i].x;
    }
};

// number of outputs of a pairing round over n inputs: sum over the runs of equal keys of ceil(length / 2)
template <class C>
PORLA_D int aff_count_outputs(const AffSrc<C>& src, int n) {
    int outs = 0, i = 0;
    while (i < n) {
        const uint32_t k = src.key(i);
        i += (i + 1 < n && src.key(i + 1) == k) ? 2 : 1;
        outs++;
    }
    return outs;
}

// One pairing round.  Returns the number of outputs; out_inf receives the mask of outputs that are the point at infinity.
template <class C>
PORLA_D int aff_round(const AffSrc<C>& src, int n, uint64_t in_inf, Affine<typename C::F>* out, uint32_t* out_keys,
                      typename C::F* pre, uint8_t* opi, uint64_t* out_inf) {
    using F = typename C::F;
    // ---- forward: classify the operations, running product of the denominators
    int i = 0, j = 0, na = 0;
    F acc = F::one();
    while (i < n) {
        const uint32_t k = src.key(i);
        const bool pair = i + 1 < n && src.key(i + 1) == k;
        uint32_t type = kOpCopy, idx = (uint32_t)i;
        if (pair) {
            const bool inf0 = (in_inf >> i) & 1u, inf1 = (in_inf >> (i + 1)) & 1u;
            if (inf0 || inf1) {
                if (inf0 && inf1) type = kOpInf;
                else idx = (uint32_t)(inf0 ? i + 1 : i);
            } else {
                F den = src.ptx(i + 1) - src.ptx(i);
                type = kOpAdd;
                if (den.is_zero()) {                       // same x: the same point (tangent) or opposite points
                    const Affine<F> p = src.pt(i), q = src.pt(i + 1);
                    if (p.y == q.y && !p.y.is_zero()) {
                        den = p.y.dbl();
                        type = kOpDbl;
                    } else {
                        type = kOpInf;
                    }
                }
                if (type != kOpInf) {
                    pre[na++] = acc;
                    acc = acc * den;
                }
            }
        } else if ((in_inf >> i) & 1u) {
            type = kOpInf;
        }
        opi[j] = (uint8_t)(idx | (type << 6));
        out_keys[j] = k;
        i += pair ? 2 : 1;
        j++;
    }
    const int outs = j;
    // ---- one inversion for the whole round (a warp runs it when any of its lanes has additions)
    F inv = F::one();
    if (na > 0) inv = acc.inverse();
    // ---- backward: peel the individual inverses off, evaluate the additions
    uint64_t infm = 0;
    for (j = outs - 1; j >= 0; j--) {
        const uint32_t idx = opi[j] & 63u, type = opi[j] >> 6;
        if (type == kOpCopy) {
            out[j] = src.pt((int)idx);
        } else if (type == kOpInf) {
            infm |= 1ull << j;
        } else {
            const Affine<F> p = src.pt((int)idx), q = src.pt((int)idx + 1);
            const F den = type == kOpAdd ? q.x - p.x : p.y.dbl();
            const F inv_j = inv * pre[--na];
            inv = inv * den;
            F lam, x3;
            if (type == kOpAdd) {
                lam = (q.y - p.y) * inv_j;
                x3 = lam.sqr() - p.x - q.x;
            } else {
                const F xx = p.x.sqr();
                lam = (xx.dbl() + xx) * inv_j;
                x3 = lam.sqr() - p.x.dbl();
            }
            Affine<F> r;
            r.x = x3;
            r.y = lam * (p.x - x3) - p.y;
            out[j] = r;
        }
    }
    *out_inf = infm;
    return outs;
}

template <class C>
__global__ void __launch_bounds__(kAffThreads, PORLA_AFF_MIN_BLOCKS)
k_accumulate_affine(const Affine<typename C::F>* __restrict__ points, const typename C::F* __restrict__ phi_x, uint32_t phi_off,
                    const uint2* __restrict__ sorted, const uint32_t* __restrict__ total_pairs,
                    XYZZ<typename C::F>* __restrict__ buckets, XYZZ<typename C::F>* __restrict__ part_head,
                    XYZZ<typename C::F>* __restrict__ part_tail, int rounds) {
    using F = typename C::F;
    const uint32_t M = *total_pairs;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t start64 = (uint64_t)t * kAffL;
    if (start64 >= M) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (M - start > (uint32_t)kAffL) ? start + kAffL : M;
    const uint32_t prev_key = start > 0 ? __ldg(&sorted[start - 1].x) : 0xffffffffu;
    const uint32_t next_key = end < M ? __ldg(&sorted[end].x) : 0xffffffffu;

    Affine<F> p1[kAffCap1], p2[kAffCap2];
    F pre[kAffMaxAdds];
    uint32_t k1[kAffCap1], k2[kAffCap2];
    uint8_t opi[kAffCap1];

    AffSrc<C> cur;
    cur.points = points;
    cur.phi_x = phi_x;
    cur.phi_off = phi_off;
    cur.pairs = sorted + start;
    cur.arr = nullptr;
    cur.keys = nullptr;
    int n = (int)(end - start);
    uint64_t inf = 0;

    if (rounds >= 1) {
        const int outs0 = aff_count_outputs<C>(cur, n);
        if (outs0 <= kAffCap1 && outs0 < n) {
            n = aff_round<C>(cur, n, inf, p1, k1, pre, opi, &inf);
            cur.pairs = nullptr;
            cur.arr = p1;
            cur.keys = k1;
            if (rounds >= 2) {
                const int outs1 = aff_count_outputs<C>(cur, n);
                if (outs1 <= kAffCap2 && outs1 < n) {
                    n = aff_round<C>(cur, n, inf, p2, k2, pre, opi, &inf);
                    cur.arr = p2;
                    cur.keys = k2;
                }
            }
        }
    }

    // ---- tail: the remaining elements of every bucket, summed with mixed additions; stored like k_accumulate stores them
    int i = 0;
    while (i < n) {
        const uint32_t key = cur.key(i);
        const bool first_run = i == 0;
        XYZZ<F> acc = XYZZ<F>::inf();
        if (!((inf >> i) & 1u)) {
            const Affine<F> p = cur.pt(i);
            acc = XYZZ<F>{p.x, p.y, F::one(), F::one()};
        }
        i++;
        while (i < n && cur.key(i) == key) {
            if (!((inf >> i) & 1u)) acc.madd(cur.pt(i));
            i++;
        }
        const bool last_run = i == n;
        if (first_run && prev_key == key) st16(part_head + t, acc);          // continues from the left (maybe also to the right)
        else if (last_run && next_key == key) st16(part_tail + t, acc);      // starts here, continues to the right
        else st16(buckets + key, acc);
    }
}

}  // namespace porla
