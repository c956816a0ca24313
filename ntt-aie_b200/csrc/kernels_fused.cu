// kernels_fused.cu -- placeholder until the register-radix kernels land.
#include "plan.h"

namespace nttb200 {
int fused_prepare(nttb200_plan *) { return NTTB200_ERR_UNSUPPORTED; }
void fused_release(nttb200_plan *) {}
int launch_fused_gs(nttb200_plan *, const int32_t *, int32_t *, size_t, bool, cudaStream_t) {
    return NTTB200_ERR_UNSUPPORTED;
}
}  // namespace nttb200
