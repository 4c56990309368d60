// tcgen05 engine -- placeholder until the UMMA kernels land: reports "shape not covered".
#include "tc_engine.cuh"
namespace deft {
int tc_conv_dispatch(const ConvAP&, const float*, int, int, int, const Epilogue&, unsigned, cudaStream_t) { return DEFT_EINVAL; }
int tc_dcn_dispatch(const DcnAP&, const float*, int, int, int, const Epilogue&, unsigned, cudaStream_t) { return DEFT_EINVAL; }
}  // namespace deft
