// tcgen05 tensor-core engine entry points (see tc_engine.cu).  Dispatchers return DEFT_EINVAL when a
// shape is not covered (or no packed weight blob was supplied) so callers drop to the SIMT fp32 engine
// -- same arithmetic, same GPU; never a CPU path.
#pragma once
#include "igemm_simt.cuh"

namespace deft {
constexpr int TC_KIND_CONV = 0, TC_KIND_DCN = 1, TC_KIND_HEADS = 2;

struct HeadsArgs {
    const float* b1;
    const float* w2;
    const float* b2;
    float* out[DEFT_MAX_HEADS];
    int c[DEFT_MAX_HEADS];
    int row0[DEFT_MAX_HEADS];
    int nheads, HW;
};

int tc_pick_bn(int ldw, int M, int kind);                // N tile for a weight of width ldw on an M-row problem
size_t tc_blob_floats(int ldw, int K, int bn);           // size of the packed hi/lo weight-tile blob
size_t tc_splitk_scratch_bytes(int ldw, int K, int M, int kind);
int tc_k_perm_taps(int K, int taps);                     // K-order rule shared by packer and dispatchers
int tc_pack_weights(const float* wp, int ldw, int K, int bn, int taps, float* blob, cudaStream_t st);
int tc_conv_dispatch(const ConvAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                     size_t scratch_bytes, cudaStream_t st);
int tc_dcn_dispatch(const DcnAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                    size_t scratch_bytes, cudaStream_t st);
int tc_pair_dispatch(const PairAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st);
int tc_heads_dispatch(const ConvAP& ap, const float* blob, int K, int M, const HeadsArgs& ha, unsigned flags, cudaStream_t st);
}  // namespace deft
