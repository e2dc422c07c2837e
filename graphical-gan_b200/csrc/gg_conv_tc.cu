// gg_conv_tc.cu — tcgen05 implicit-GEMM convolution (placeholder until the probes are verified on hardware).
#include "gg_tc_common.cuh"
namespace gg {
int conv_tc_fwd(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, int,
                float, void*, size_t, cudaStream_t, bool* handled) { *handled = false; return GG_OK; }
int conv_tc_dgrad(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, int,
                  float, void*, size_t, cudaStream_t, bool* handled) { *handled = false; return GG_OK; }
int conv_tc_wgrad(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, void*, size_t,
                  cudaStream_t, bool* handled) { *handled = false; return GG_OK; }
size_t conv_tc_wgrad_workspace(int, int, int, int, int, int, int, int, int) { return 0; }
}  // namespace gg
