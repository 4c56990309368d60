// tcgen05 tensor-core engine entry points (see tc_engine.cu).  Dispatchers return DEFT_EINVAL when a
// shape is not covered (or no packed weight blob was supplied) so callers drop to the SIMT fp32 engine
// -- same arithmetic, same GPU; never a CPU path.
#pragma once
#include "igemm_simt.cuh"

namespace deft {
int tc_tile_n(int ldw);                                  // N tile the conv engine uses for a weight of width ldw
size_t tc_blob_floats(int ldw, int K, int bn);           // size of the packed hi/lo weight-tile blob
int tc_pack_weights(const float* wp, int ldw, int K, int bn, float* blob, cudaStream_t st);
int tc_conv_dispatch(const ConvAP& ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st);
int tc_dcn_dispatch(const DcnAP& ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st);
int tc_pair_dispatch(const PairAP& ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st);
}  // namespace deft
