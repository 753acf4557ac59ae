// placeholder: the tcgen05 int8-slice contraction lands here
#include "common.cuh"
using namespace b200w;
int b200w_tom_prepare_i8(const double*, int64_t, int64_t, int64_t, void*, int64_t, double*, double*, cudaStream_t) {
  return fail(B200W_E_UNSUPPORTED, "i8 TOM path not built yet");
}
int b200w_tom_compute_i8(const double*, const void*, int64_t, const double*, const double*, int64_t, int64_t, int64_t, int, int, int, double*, cudaStream_t) {
  return fail(B200W_E_UNSUPPORTED, "i8 TOM path not built yet");
}
