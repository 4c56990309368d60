// tcgen05 tensor-core engine entry points (see tc_engine.cu).  Return DEFT_EINVAL when a shape is
// not covered so callers drop to the SIMT fp32 engine (same arithmetic, same GPU -- never a CPU path).
#pragma once
#include "igemm_simt.cuh"

namespace deft {
int tc_conv_dispatch(const ConvAP& ap, const float* Wp, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st);
int tc_dcn_dispatch(const DcnAP& ap, const float* Wp, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st);
}  // namespace deft
