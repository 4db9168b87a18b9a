// Four lanes per point operation (round 2; VERDICT r1 "what's weak" 5, 7, 11: the latency of DEPENDENT point additions).
//
// A lone warp needs ~838 cycles per field product whatever the instruction-level parallelism (lat.cu), and the 14 products of
// an XYZZ addition run one after the other in a thread: 12.4 k cycles = 6.3 us per addition, which is the whole cost of the
// tree sums behind Porla-sized calls (12 dependent additions of a 128-term commitment), of the bucket reduction below 2^18
// terms and of the window combine.  The products of one addition are not all dependent, though: they form FOUR levels of up
// to four independent products.  Here the four lanes of a quad (lane & 3 = role) hold one coordinate each
//     role 0: X      role 1: Y      role 2: ZZ      role 3: ZZZ
// and every level is ONE field product per lane on role-selected operands, with warp shuffles in between:
//     addition  (add-2008-s)    L1  U1 = X1 ZZ2 | S1 = Y1 ZZZ2 | U2 = X2 ZZ1 | S2 = Y2 ZZZ1      P = U2 - U1, R = S2 - S1
//                               L2  PP = P^2    | RR = R^2     | T1 = ZZ1 ZZ2 | T2 = ZZZ1 ZZZ2
//                               L3  Q = U1 PP   | PPP = P PP   | ZZ3 = T1 PP  | -                 X3 = RR - PPP - 2Q
//                               L4  S1 PPP      | R (Q - X3)   | -            | ZZZ3 = T2 PPP      Y3 = R (Q - X3) - S1 PPP
//     doubling  (dbl-2008-s-1)  L1  XX = X^2    | V = U^2 (U = 2Y)                                M = 3 XX
//                               L2  S = X V     | W = U V      | ZZ3 = V ZZ   | MM = M^2           X3 = MM - 2S
//                               L3  M (S - X3)  | W Y          | -            | ZZZ3 = W ZZZ       Y3 = M (S - X3) - W Y
// A warp then carries 8 point operations instead of 32 and executes 4 (3) products instead of 14 (9): the latency of one
// addition drops ~3x for 1.4x the pipe work per addition -- the right trade wherever the dependent chain, not the pipe, is the
// bound.  All 32 lanes of the warp must call these functions together (full-mask shuffles); quads without work pass infinity.
// Exceptional cases are exact as in ec.cuh: infinity operands, P + P (falls into quad_dbl) and P - P.
#pragma once
#include "ec.cuh"

namespace porla {

#ifdef __CUDACC__
constexpr unsigned kFullMask = 0xffffffffu;

template <class F>
PORLA_D F quad_bcast(const F& a, int role) {      // the value held by lane `role` of this quad, to all four lanes
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(kFullMask, a.v[i], role, 4);
    return r;
}
template <class F>
PORLA_D F quad_xor(const F& a, int m) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(kFullMask, a.v[i], m, 4);
    return r;
}
template <class F>
PORLA_D F fsel(bool take_a, const F& a, const F& b) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = take_a ? a.v[i] : b.v[i];
    return r;
}
PORLA_D bool quad_flag(bool f, int role) { return __shfl_sync(kFullMask, (int)f, role, 4) != 0; }

// One point per quad: this lane's coordinate (X, Y, ZZ or ZZZ by role).  Infinity: all four coordinates zero.
template <class F>
struct QuadPoint {
    F c;
    PORLA_D static QuadPoint inf() { return QuadPoint{F::zero()}; }
    PORLA_D bool is_inf() const { return quad_flag(c.is_zero(), 2); }          // ZZ == 0
    // the quad's four lanes read the 128-byte record (coalesced: 32 bytes per lane)
    PORLA_D static QuadPoint load(const XYZZ<F>* p) {
        const int role = threadIdx.x & 3;
        return QuadPoint{ld16(reinterpret_cast<const F*>(p) + role)};
    }
    // the same through the coherent path: for records the running kernel itself has written (k_reduce_scan re-reads the suffix
    // sums other quads of its block stored; the non-coherent path of ld16 is only for data that is read-only during the launch)
    PORLA_D static QuadPoint load_coherent(const XYZZ<F>* p) {
        const int role = threadIdx.x & 3;
        const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const F*>(p) + role);
        QuadPoint r;
        uint4* d = reinterpret_cast<uint4*>(&r.c);
        d[0] = s[0];
        d[1] = s[1];
        return r;
    }
    PORLA_D void store(XYZZ<F>* p) const {
        const int role = threadIdx.x & 3;
        st16(reinterpret_cast<F*>(p) + role, c);
    }
    // affine table entry (x = y = 0: infinity), optionally negated
    PORLA_D static QuadPoint load_affine(const Affine<F>* p, bool negate) {
        const int role = threadIdx.x & 3;
        F v = F::one();
        if (role < 2) v = ld16(reinterpret_cast<const F*>(p) + role);
        const bool z = v.is_zero();
        const bool zx = quad_flag(z, 0), zy = quad_flag(z, 1);     // both shuffles unconditionally (no short-circuit: every lane takes part)
        const bool inf = zx & zy;
        if (role == 1 && negate) v = v.neg();
        return QuadPoint{inf ? F::zero() : v};
    }
    // from / to a point held whole by ONE lane of the quad (`owner` = its role)
    PORLA_D static QuadPoint scatter(const XYZZ<F>& p, int owner) {
        const int role = threadIdx.x & 3;
        // every lane picks coordinate `role` of the owner's point: shuffle each coordinate from the owner, keep one
        F r;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t x = __shfl_sync(kFullMask, p.x.v[i], owner, 4), y = __shfl_sync(kFullMask, p.y.v[i], owner, 4);
            const uint32_t zz = __shfl_sync(kFullMask, p.zz.v[i], owner, 4), zzz = __shfl_sync(kFullMask, p.zzz.v[i], owner, 4);
            r.v[i] = role == 0 ? x : role == 1 ? y : role == 2 ? zz : zzz;
        }
        return QuadPoint{r};
    }
    PORLA_D XYZZ<F> gather() const {              // the whole point, in every lane of the quad
        XYZZ<F> p;
        p.x = quad_bcast(c, 0);
        p.y = quad_bcast(c, 1);
        p.zz = quad_bcast(c, 2);
        p.zzz = quad_bcast(c, 3);
        return p;
    }
};

// 2 a
template <class F>
PORLA_D QuadPoint<F> quad_dbl(const QuadPoint<F>& a) {
    const int role = threadIdx.x & 3;
    const bool a_inf = a.is_inf();
    const F x = quad_bcast(a.c, 0), y = quad_bcast(a.c, 1);
    const F u = y.dbl();
    // L1: XX | V
    const F m1 = fsel(role == 0, x, u).sqr();
    const F xx = quad_bcast(m1, 0), v = quad_bcast(m1, 1);
    const F m = xx.dbl() + xx;
    // L2: S = X V | W = U V | ZZ3 = ZZ V | MM = M M
    const F l2a = role == 0 ? x : role == 1 ? u : role == 2 ? a.c : m;
    const F m2 = l2a * fsel(role == 3, m, v);
    const F s = quad_bcast(m2, 0), w = quad_bcast(m2, 1), mm = quad_bcast(m2, 3);
    const F x3 = mm - s.dbl();
    // L3: M (S - X3) | W Y | - | ZZZ3 = W ZZZ
    const F l3a = role == 0 ? m : w;
    const F l3b = role == 0 ? s - x3 : role == 1 ? y : a.c;
    const F m3 = l3a * l3b;
    const F md = quad_bcast(m3, 0);
    F r = role == 0 ? x3 : role == 1 ? md - m3 : role == 2 ? m2 : m3;
    // y == 0 cannot occur on a prime-order curve; infinity doubles to infinity
    return QuadPoint<F>{a_inf ? F::zero() : r};
}

// a + b
template <class F>
PORLA_D QuadPoint<F> quad_add(const QuadPoint<F>& a, const QuadPoint<F>& b) {
    const int role = threadIdx.x & 3;
    const bool a_inf = a.is_inf(), b_inf = b.is_inf();
    // L1: lane r multiplies its coordinate of a by the coordinate role ^ 2 of b
    const F m1 = a.c * quad_xor(b.c, 2);                     // U1 | S1 | U2 | S2
    const F u1 = quad_bcast(m1, 0), s1 = quad_bcast(m1, 1), u2 = quad_bcast(m1, 2), s2 = quad_bcast(m1, 3);
    const F p = u2 - u1, r = s2 - s1;
    const bool p_zero = p.is_zero(), r_zero = r.is_zero();   // identical in the four lanes
    // L2: PP | RR | T1 = ZZ1 ZZ2 | T2 = ZZZ1 ZZZ2
    const F l2a = role == 0 ? p : role == 1 ? r : a.c;
    const F l2b = role == 0 ? p : role == 1 ? r : b.c;
    const F m2 = l2a * l2b;
    const F pp = quad_bcast(m2, 0), rr = quad_bcast(m2, 1);
    // L3: Q = U1 PP | PPP = P PP | ZZ3 = T1 PP | -
    const F l3a = role == 0 ? u1 : role == 1 ? p : m2;
    const F m3 = l3a * pp;
    const F q = quad_bcast(m3, 0), ppp = quad_bcast(m3, 1);
    const F x3 = rr - ppp - q.dbl();
    // L4: S1 PPP | R (Q - X3) | - | ZZZ3 = T2 PPP
    const F l4a = role == 0 ? s1 : role == 1 ? r : m2;
    const F l4b = role == 1 ? q - x3 : ppp;
    const F m4 = l4a * l4b;
    const F sp = quad_bcast(m4, 0);
    F out = role == 0 ? x3 : role == 1 ? m4 - sp : role == 2 ? m3 : m4;
    const bool both = !a_inf && !b_inf;
    if (both && p_zero) out = F::zero();                     // P - P (P + P is patched below)
    if (b_inf) out = a.c;
    else if (a_inf) out = b.c;
    if (__any_sync(kFullMask, both && p_zero && r_zero)) {   // P + P somewhere in the warp: rare, warp-uniform branch
        const QuadPoint<F> d = quad_dbl(a);
        if (both && p_zero && r_zero) out = d.c;
    }
    return QuadPoint<F>{out};
}
#endif  // __CUDACC__

}  // namespace porla
