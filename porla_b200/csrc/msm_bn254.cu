#include "msm_impl.cuh"
namespace porla {
PORLA_INSTANTIATE_CURVE(Bn254)
}
