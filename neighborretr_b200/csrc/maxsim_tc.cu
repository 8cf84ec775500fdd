// placeholder until the tcgen05 kernels land (next commit)
#include "common.cuh"
#include "nrhead_internal.h"
int nr_maxsim_fwd_tc(const void*, const void*, const float*, const int64_t*, const int64_t*, int64_t, int64_t,
                     int64_t, int64_t, int64_t, float, float*, int64_t, int64_t, float*, int64_t, int64_t, int,
                     float*, uint8_t*, cudaStream_t) {
  nr::set_error("nr_maxsim_fwd: NR_PREC_BF16 not built");
  return -2;
}
int nr_maxsim_bwd_tc(int, const void*, const float*, const int64_t*, const int64_t*, const uint8_t*, const float*,
                     int64_t, int64_t, float, int64_t, int64_t, int64_t, int64_t, int64_t, float*, cudaStream_t) {
  nr::set_error("nr_maxsim_bwd: NR_PREC_BF16 not built");
  return -2;
}
