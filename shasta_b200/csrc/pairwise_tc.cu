// Tensor-core (tcgen05 / TMEM) variants of the pairwise stage. Placeholder until the tcgen05 tile kernel lands:
// the entry point exists so the C ABI is stable, and reports SHASTA_ERR_UNSUPPORTED instead of silently falling back.
#include "common.cuh"

namespace shasta {

int launch_pairwise_tc(const float* packed, int B, int M, float* ws, const WsLayout& L, int variant,
                       cudaStream_t s) {
  (void)packed, (void)B, (void)M, (void)ws, (void)L, (void)s;
  set_error("pairwise variant %d (tcgen05) is not built in this revision", variant);
  return SHASTA_ERR_UNSUPPORTED;
}

}  // namespace shasta
