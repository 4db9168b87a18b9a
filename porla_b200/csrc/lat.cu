// Latency probe of the field / point operations for ONE block on ONE SM (development aid behind
// porla_debug_latency): how long a dependent chain of operations takes per operation when a warp has a
// scheduler to itself, and how much instruction-level parallelism between independent chains recovers.
// The small-MSM kernels (small_kernels.cuh), the bucket reduction below 2^18 points and the butterfly
// kernel all live in this regime.
//   mode 0  x = x * y                      one chain of Montgomery / special-form products
//   mode 1  two independent product chains in one thread
//   mode 2  four independent product chains in one thread
//   mode 3  x = x^2
//   mode 4  acc.add(q)    XYZZ + XYZZ, inlined field type
//   mode 5  acc.add(q)    XYZZ + XYZZ, outlined multiplier (FC)
//   mode 6  acc.madd(p)   XYZZ + affine, inlined
//   mode 7  acc = acc.dbl(), inlined
//   mode 8  x = 1/x + y    binary-GCD inversion (fp_inv.cuh);   mode 9: the same with the Fermat inversion
//   mode + 100: the same chain on the whole device (4 blocks per SM): ns_per_op is then the time per operation of a thread under load
// Measured (B200, BN254): 838 cycles per product with one warp per scheduler whether the thread runs one,
// two or four independent chains -- also with the two products interleaved round by round at source level
// (tried and removed) -- against ~600 cycles per product when two warps share a scheduler: the lone warp is
// bound by its own issue cadence on the multiplier pipe, not by the carry chains.
#include <cstdio>

#include "../../include/porla_multiexp.h"
#include "msm.h"
#include "msm_kernels.cuh"

struct porla_table {   // as in abi.cu
    porla::PointTable t;
};

namespace porla {

template <class C>
__global__ void k_latency(int mode, int iters, const Affine<typename C::F>* __restrict__ seed, unsigned long long* cycles,
                          XYZZ<typename C::F>* sink) {
    using F = typename C::F;
    using FC = typename C::FC;
    Affine<F> p = seed[threadIdx.x & 1];
    Affine<F> q = seed[2 + (threadIdx.x & 1)];
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    XYZZ<F> other = XYZZ<F>::from_affine(q);
    other = other.dbl();
    F x0 = p.x, x1 = p.y, x2 = q.x, x3 = q.y;
    const F y = q.y;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) x0 = x0 * y;
    } else if (mode == 1) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            x0 = x0 * y;
            x1 = x1 * y;
        }
    } else if (mode == 2) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            x0 = x0 * y;
            x1 = x1 * y;
            x2 = x2 * y;
            x3 = x3 * y;
        }
    } else if (mode == 3) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) x0 = x0.sqr();
    } else if (mode == 4) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) acc.add(other);
    } else if (mode == 5) {
        XYZZ<FC> a2 = *reinterpret_cast<XYZZ<FC>*>(&acc), o2 = *reinterpret_cast<XYZZ<FC>*>(&other);
#pragma unroll 1
        for (int i = 0; i < iters; i++) a2.add(o2);
        acc = *reinterpret_cast<XYZZ<F>*>(&a2);
    } else if (mode == 6) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) acc.madd_finite(q);
    } else if (mode == 7) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) acc = acc.dbl();
    } else if (mode == 8) {   // binary-GCD inversion (fp_inv.cuh), a dependent chain
#pragma unroll 1
        for (int i = 0; i < iters; i++) x0 = x0.inverse() + y;
    } else {                  // Fermat inversion, for comparison
#pragma unroll 1
        for (int i = 0; i < iters; i++) x0 = x0.inverse_fermat() + y;
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0 && blockIdx.x == 0) cycles[threadIdx.x / 32] = (unsigned long long)(t1 - t0);
    acc.x = acc.x + x0 + x1 + x2 + x3;
    if (acc.x.v[0] == 0x12345678u && acc.y.v[1] == 0x9abcdef0u) *sink = acc;   // keep the work alive
}

}  // namespace porla

extern "C" int porla_debug_latency(int curve, int mode, int warps, int iters, double* cycles_per_op, double* ns_per_op) {
    using namespace porla;
    device_init();
    // mode + 100: the same chain in `warps` warps per block on 4 blocks per SM of the whole device (throughput under load)
    int blocks = 1;
    if (mode >= 100) {
        mode -= 100;
        cudaDeviceProp prop;
        int dev = 0;
        PORLA_CUDA(cudaGetDevice(&dev));
        PORLA_CUDA(cudaGetDeviceProperties(&prop, dev));
        blocks = prop.multiProcessorCount * 4;
    }
    if (warps < 1) warps = 1;
    if (warps > 32) warps = 32;
    // four affine points: small multiples of the generator, built on the host side of the library
    porla_table* t = nullptr;
    uint32_t ks[4][8] = {{3}, {5}, {7}, {11}};
    t = porla_table_create_multiples(curve, ks, 4, PORLA_SCALAR_LE32, 0, nullptr);
    unsigned long long* d_cycles = nullptr;
    void* d_sink = nullptr;
    PORLA_CUDA(cudaMalloc(&d_cycles, 32 * 8));
    PORLA_CUDA(cudaMalloc(&d_sink, 256));
    cudaEvent_t e0, e1;
    PORLA_CUDA(cudaEventCreate(&e0));
    PORLA_CUDA(cudaEventCreate(&e1));
    const void* pts = t->t.d_points;
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        PORLA_CUDA(cudaEventRecord(e0));
        if (curve == kCurveBn254)
            k_latency<Bn254><<<blocks, warps * 32>>>(mode, iters, reinterpret_cast<const Affine<Bn254::F>*>(pts), d_cycles,
                                               reinterpret_cast<XYZZ<Bn254::F>*>(d_sink));
        else
            k_latency<Secp256k1><<<blocks, warps * 32>>>(mode, iters, reinterpret_cast<const Affine<Secp256k1::F>*>(pts), d_cycles,
                                                   reinterpret_cast<XYZZ<Secp256k1::F>*>(d_sink));
        PORLA_CUDA(cudaEventRecord(e1));
        PORLA_CUDA(cudaEventSynchronize(e1));
        PORLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    unsigned long long h[32] = {};
    PORLA_CUDA(cudaMemcpy(h, d_cycles, (size_t)warps * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
    const double ops = (double)iters * (mode == 1 ? 2.0 : mode == 2 ? 4.0 : 1.0);
    *cycles_per_op = (double)mx / ops;
    *ns_per_op = (double)ms * 1e6 / ops;
    PORLA_CUDA(cudaEventDestroy(e0));
    PORLA_CUDA(cudaEventDestroy(e1));
    PORLA_CUDA(cudaFree(d_cycles));
    PORLA_CUDA(cudaFree(d_sink));
    porla_table_destroy(t);
    return 0;
}
