// batch_kernels.cu -- batched multi-instance mode (one CTA per LP).  Not implemented yet:
// the entry points exist so the ABI is complete and fail loudly.
#include "../../include/mllp_b200.h"

extern "C" {
int mllp_batch_create(int32_t, int32_t, const int32_t*, const int32_t*, const int64_t*, const int64_t*,
                      const int32_t*, const int32_t*, const double*, int, uint32_t, mllp_batch_t*) { return MLLP_E_STATE; }
int mllp_batch_destroy(mllp_batch_t) { return 0; }
int mllp_batch_info(mllp_batch_t, int64_t*) { return MLLP_E_STATE; }
int mllp_batch_run(mllp_batch_t, double*, double*, const double*, const double*, const double*, const double*,
                   int32_t, double*, void*) { return MLLP_E_STATE; }
int mllp_batch_solve(mllp_batch_t, double*, double*, const double*, const double*, const double*, double, int32_t,
                     int32_t, double, double*, void*) { return MLLP_E_STATE; }
}
