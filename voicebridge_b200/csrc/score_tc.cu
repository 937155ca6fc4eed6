// score_tc.cu — tensor-core (tcgen05/TMEM) dense scoring.  Placeholder until the kernel lands: reports "unavailable",
// so every call takes the FP32 SIMT path.
#include "common.h"

namespace vb {
int score_tc_prepare(vbgpu_gmm_t, const float *, const float *, const float *, int32_t) { return 0; }
bool score_tc_available(vbgpu_gmm_t) { return false; }
int score_tc_launch(vbgpu_gmm_t, const float *, int64_t, int32_t, float *, int32_t, cudaStream_t) {
  return fail(VBGPU_ERR_INVALID, "tensor-core scorer not built");
}
void score_tc_release(vbgpu_gmm_t) {}
int score_tc_update_gconsts(vbgpu_gmm_t, const float *) { return 0; }
}  // namespace vb
