// K5 placeholder - replaced by the SIMT + tcgen05 implementation.
#include "common.cuh"
extern "C" {
int rm_cin_layer_fwd(const float*, int64_t, const float*, int64_t, const float*, const float*, int64_t, int32_t, int32_t,
                     int32_t, int32_t, int32_t, int32_t, float*, float*, void*, size_t, void*) {
  rm::set_error("rm_cin_layer_fwd: not built yet");
  return RM_E_UNSUPPORTED;
}
size_t rm_cin_layer_workspace_bytes(int64_t, int32_t, int32_t, int32_t, int32_t, int32_t) { return 0; }
int rm_cin_layer_bwd(const float*, int64_t, const float*, int64_t, const float*, const float*, const float*, int64_t, int32_t,
                     int32_t, int32_t, int32_t, int32_t, int32_t, float*, float*, float*, float*, int64_t, void*,
                     size_t, void*) {
  rm::set_error("rm_cin_layer_bwd: not built yet");
  return RM_E_UNSUPPORTED;
}
size_t rm_cin_layer_bwd_workspace_bytes(int64_t, int32_t, int32_t, int32_t, int32_t, int32_t) { return 0; }
}
