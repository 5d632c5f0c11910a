// Instantiations of the CTA-pair tcgen05 contraction for the fp16-split engine (own translation unit).
#include "gemm_tc2.cuh"

namespace usf {
int launch_gemm_tc2_3xf16(const usf_linear_args* a, const Epilogue& ep, cudaStream_t st, int bn) {
  return launch_gemm_tc2_terms<3, tc2::KIND_F16>(a, ep, st, bn);
}
}  // namespace usf
