#include "msm_impl.cuh"
namespace porla {
PORLA_INSTANTIATE_CURVE(Bn254)

// curve-independent: the data-side FFT butterfly (msm.cu computes the Barrett constant and calls this)
void data_butterfly_launch(uint32_t* d_blocks, uint32_t n_blocks, uint32_t chunks, uint32_t m, const uint8_t* d_twiddles,
                           const uint32_t* lcm16, const uint32_t* mu17, cudaStream_t stream) {
    DataFftParams prm;
    for (int i = 0; i < 16; i++) prm.lcm[i] = lcm16[i];
    for (int i = 0; i < 17; i++) prm.mu[i] = mu17[i];
    data_butterfly_impl(d_blocks, n_blocks, chunks, m, d_twiddles, prm, stream);
}
}
