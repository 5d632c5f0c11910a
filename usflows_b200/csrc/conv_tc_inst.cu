// Instantiations of the implicit-GEMM tcgen05 convolution (own translation unit: parallel build).
#include "conv_tc.cuh"

namespace usf {
int launch_conv_tc(const usf_linear_args* a, const ConvGeom& g, const Epilogue& ep, cudaStream_t st) {
  const int bn = a->N <= 32 ? 32 : 64;
  if (a->N > 64 || conv_tc_stages(bn, a->K) == 0)
    return fail(USF_ERR_UNSUPPORTED, "usf_conv2d_rows: N <= 64 and the whole weight resident in shared memory%s%s");
  if (bn == 32) return launch_conv_tc_cfg<32>(a, g, ep, st);
  return launch_conv_tc_cfg<64>(a, g, ep, st);
}
}  // namespace usf
