// kernels_tc.cu - tcgen05 / TMEM contractions (placeholder until the shifted-window kernels land in this file).
#include "common.cuh"

bool tc_supported(const Geom&) { return false; }
size_t tc_packed_bytes(const Geom&) { return 0; }
int tc_pack_weights(const LaunchCtx&, const Geom&, const float*, void*) { return -1; }
int tc_gather_gemm(const LaunchCtx&, const Geom&, View, const void*, View, double*) {
  svae_global_error() = "tcgen05 path not built";
  return -1;
}
