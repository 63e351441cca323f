// tcgen05 / TMEM / TMA implicit-GEMM conv, deconv and linear kernels (sm_100a).
// Stage 0: every entry reports DMV_E_UNSUPPORTED_SHAPE so DMV_ALGO_AUTO uses the SIMT path.
#include "common.cuh"
#include "conv_impl.h"

namespace dmv {
static int unsupported(const char* what) { return fail(DMV_E_UNSUPPORTED_SHAPE, what); }
size_t tc_wgrad_workspace(int, int, int, long long) { return 0; }
int tc_conv_fwd(const void*, int, const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, cudaStream_t) { return unsupported("tc_conv_fwd: shape not covered"); }
int tc_conv_dgrad(const void*, const void*, void*, int, int, int, int, int, int, int, int, cudaStream_t) { return unsupported("tc_conv_dgrad: shape not covered"); }
int tc_conv_wgrad(const void*, int, const void*, float*, float*, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t) { return unsupported("tc_conv_wgrad: shape not covered"); }
int tc_deconv_fwd(const void*, const void*, void*, int, int, int, int, int, int, int, int, int, int, cudaStream_t) { return unsupported("tc_deconv_fwd: shape not covered"); }
int tc_deconv_dgrad(const void*, int, const void*, void*, int, int, int, int, int, int, int, int, cudaStream_t) { return unsupported("tc_deconv_dgrad: shape not covered"); }
int tc_deconv_wgrad(const void*, const void*, int, float*, int, int, int, int, int, int, int, int, void*, size_t, cudaStream_t) { return unsupported("tc_deconv_wgrad: shape not covered"); }
int tc_linear_fwd(const void*, const void*, const float*, void*, int, int, int, int, cudaStream_t) { return unsupported("tc_linear_fwd: shape not covered"); }
int tc_linear_dgrad(const void*, const void*, void*, int, int, int, cudaStream_t) { return unsupported("tc_linear_dgrad: shape not covered"); }
int tc_linear_wgrad(const void*, const void*, float*, float*, int, int, int, void*, size_t, cudaStream_t) { return unsupported("tc_linear_wgrad: shape not covered"); }
}  // namespace dmv
