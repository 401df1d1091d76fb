// K2+K3, tcgen05/TMEM/TMA implementation (placeholder until the tensor-core kernel lands).
#include "infonce.cuh"

namespace avssl {

bool infonce_tc_supported(int, int, int) { return false; }

int launch_infonce_tc(const InfoNceParams&, int, cudaStream_t) {
  set_error("moco_infonce: the tcgen05 implementation is not available in this build");
  return AVSSL_ERR_UNSUPPORTED;
}

}  // namespace avssl
