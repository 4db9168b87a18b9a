// Host build of the binary-GCD inversion (porla_b200/csrc/fp_inv.cuh is plain C++): test hook for tests/test_fp_inv_host.py.
#include "../../porla_b200/csrc/ec.cuh"

using namespace porla;

extern "C" void fp_inv_plain(int curve, const uint32_t* x, uint32_t* out, int n) {
    for (int i = 0; i < n; i++) {
        if (curve == 0) fp_inverse_plain<Bn254FpParams, 254>(x + 8 * i, out + 8 * i);
        else fp_inverse_plain<Secp256k1FpParams, 256>(x + 8 * i, out + 8 * i);
    }
}

// internal form in, internal form out (Montgomery for BN254)
extern "C" void fp_inv_internal(int curve, const uint32_t* x, uint32_t* out, int n) {
    for (int i = 0; i < n; i++) {
        if (curve == 0) {
            Bn254Fp a;
            for (int k = 0; k < 8; k++) a.v[k] = x[8 * i + k];
            Bn254Fp r = a.inverse();
            for (int k = 0; k < 8; k++) out[8 * i + k] = r.v[k];
        } else {
            SecpFp a;
            for (int k = 0; k < 8; k++) a.v[k] = x[8 * i + k];
            SecpFp r = a.inverse();
            for (int k = 0; k < 8; k++) out[8 * i + k] = r.v[k];
        }
    }
}
