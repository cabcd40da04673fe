// placeholder until the tcgen05 kernel lands
#include "common.cuh"
#include "nn.cuh"
bool nn_tc_supported(int d) { (void)d; return false; }
int nn_db_norm_launch(const float*, int, int, float*, cudaStream_t) { return ST3R_OK; }
int nn_tc_launch(const float*, const int32_t*, const int32_t*, int, const float*, int, int, const float*,
                 unsigned long long*, cudaStream_t) {
  st3r_set_error("tcgen05 NN kernel not built");
  return ST3R_ERR_UNSUPPORTED;
}
