// dfm_fast.cu -- fused per-panel EM fast path for small state dimension (sm_100a only).
// Placeholder until the fused kernel lands: reports "not supported" so dfm_em_kalman uses the
// general path.
#include "../../include/dfm_b200.h"
#include <cuda_runtime.h>
struct dfm_handle;
extern "C" int dfm_em_fused_supported(const dfm_em_opts* o) { (void)o; return 0; }
extern "C" int dfm_em_kalman_fused(dfm_handle*, const double*, const dfm_em_opts*, double*, double*, double*, double*,
                                   const double*, double*, double*, double*, int*, int*, long long*) {
  return DFM_ERR_UNSUPPORTED;
}
