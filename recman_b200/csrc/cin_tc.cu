// K5 tensor-core path (tcgen05) - placeholder until the kernel lands: reports "unsupported" so the
// dispatcher uses the CUDA-core path.
#include "cin.cuh"
namespace rm {
bool cin_tc_supported(int64_t, int, int, int, int) { return false; }
size_t cin_tc_fwd_workspace(int64_t, int, int, int, int, int) { return 0; }
int cin_fwd_tc(const float*, int64_t, const float*, int64_t, const float*, const float*, int64_t, int, int, int, int,
               int, int, float*, float*, void*, size_t, cudaStream_t) {
  set_error("cin_fwd_tc: not built");
  return RM_E_UNSUPPORTED;
}
}  // namespace rm
