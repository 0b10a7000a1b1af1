// tcgen05/TMEM/TMA stacked-GEMM likelihood path — placeholder until the kernel lands.
#include "sapt_internal.h"
namespace sapt {
bool tc_topology_supported(int, int, int) { return false; }
int tc_prepare_dataset(sapt_problem*, int, const float*, int64_t) { return SAPT_OK; }
int launch_loglik_tc(sapt_problem*, int, const float*, int64_t, int, const int32_t*, const int32_t*, double*, int32_t*, double*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return SAPT_ERR_UNSUPPORTED;
}
}  // namespace sapt
