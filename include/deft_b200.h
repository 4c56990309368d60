/* deft_b200 -- C ABI of the B200-native DEFT per-frame dense hot path (libdeft_b200.so).
 *
 * The reference (MedChaabane/DEFT) has no FFI of its own: its boundary is the Python surface of
 * src/lib/model (+ the external dcn_v2 extension + kalman_filter_lstm).  Every entry point below
 * names the reference interface it replaces (file:line relative to /root/reference/src/lib);
 * deft_b200/ (Python) re-exposes them under the reference's own class/function names and
 * INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, <0 = DEFT_E*; deft_last_error() gives the message of
 *     the last failure on the calling thread.
 *   - all pointers are DEVICE pointers unless the name ends in _host; fp32 everywhere, int64 indices.
 *   - activations are NHWC ("channels-last": element (n,y,x,c) at ((n*H+y)*W+x)*pitch + c, where
 *     pitch >= C lets a tensor live inside a wider concat buffer).  deft_nchw_to_nhwc /
 *     deft_nhwc_to_nchw convert at the reference's NCHW boundary; head outputs are plain NCHW.
 *   - the library is stream-ordered, never allocates or frees device memory, never synchronises
 *     the host, keeps no pointer after return (CUDA-graph capturable).  Scratch is caller-owned
 *     (`ws`, size from the matching *_workspace_bytes()).
 *   - weights come pre-packed by deft_b200/packing.py (layouts documented per function).
 */
#ifndef DEFT_B200_H
#define DEFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* deft_stream_t; /* cudaStream_t */

#define DEFT_OK 0
#define DEFT_EINVAL (-1)   /* bad argument / unsupported shape        */
#define DEFT_ECUDA (-2)    /* CUDA launch/runtime error                */
#define DEFT_EWORKSPACE (-3) /* workspace too small                    */

#define DEFT_FLAG_RELU 1u
#define DEFT_FLAG_TENSORCORE 2u   /* use the tcgen05 (3xTF32) engine where available */
#define DEFT_FLAG_TF32X1 4u       /* single-pass TF32 (faster, ~1e-3 per layer)      */
#define DEFT_FLAG_NO_DIRECT 8u    /* conv: skip the specialised direct kernels (testing the generic engines) */

#define DEFT_MAX_SRC 4
#define DEFT_MAX_HEADS 12
#define DEFT_AFE_MAPS 13

int deft_version(void);
const char* deft_last_error(void);
/* 1 if the current device is sm_100 (B200); the product path refuses to run elsewhere. */
int deft_device_ok(void);
/* Total CUDA kernels this library has launched in this process (diagnostics / bench.py's gpu_launches). */
unsigned long long deft_launch_count(void);
/* Programmatic dependent launch of the library's kernels (default on; environment DEFT_PDL overrides both).  It
 * shortens an isolated frame (a kernel's prologue overlaps its predecessor's tail) but costs throughput when several
 * frames are in flight on different streams (early-launched CTAs hold an SM while they wait), so FramePipeline
 * switches it off for depth > 1.  Applies to launches (and graph captures) made after the call. */
void deft_set_pdl(int enabled);
int deft_get_pdl(void);
/* Engine observability.  Every GEMM-shaped entry point tries the tcgen05 engine first (DEFT_FLAG_TENSORCORE) and drops
 * to the fp32 SIMT engine when the tensor-core dispatcher does not cover the shape.  deft_simt_fallback_count() counts
 * those drops (a DLA-34 frame must report 0: tests/test_gpu_frame.py), deft_debug_set_engine_log(1) or
 * DEFT_DEBUG_ENGINE=1 prints one stderr line per drop.  deft_debug_set_trace: bring-up only -- a device buffer of
 * 4*96*4 int64 into which CTA 0 of the tcgen05 kernels stamps clock64() per protocol point (tools/tc_trace.py);
 * NULL disarms it. */
unsigned long long deft_simt_fallback_count(void);
/* The second DCN gather kernel (TsDcnPatchAP: bilinear corners out of a TMA-loaded halo patch in shared memory).  Built and
 * parity-tested in round 2, measured slower than the global gather (profiles/r2_dcn_patch_experiment.md), so it is OFF unless
 * DEFT_DCN_PATCH / deft_debug_set_dcn_patch says otherwise: 0 off, 1 on for maps that tile into 16x8 patches, 2 on for every
 * map.  deft_dcn_patch_launch_count(): gather-GEMM launches that took it since load (lets a test prove which kernel ran). */
void deft_debug_set_dcn_patch(int mode);
unsigned long long deft_dcn_patch_launch_count(void);
void deft_debug_set_engine_log(int enabled);
void deft_debug_set_trace(void* device_buffer);

/* ---- tcgen05 weight images: a packed weight Wp[K][ldw] re-laid as ready-to-copy shared-memory tiles
 * ([n_tile][k_block][hi|lo][bn x 32] fp32, SWIZZLE_128B, split x = hi + lo for 3xTF32).
 * kind: 0 = conv / plain GEMM, 1 = deformable gather (bn <= 128), 2 = fused heads (bn = 256).
 * deft_tc_tile_n picks bn for an M-row problem; the same bn must be passed to every call that uses the blob.
 * deft_tc_scratch_bytes: split-K scratch the engine wants for that problem (0 = none; optional, without it
 * small-M layers simply run unsplit). */
int deft_tc_tile_n(int ldw, int M, int kind);
size_t deft_tc_blob_bytes(int ldw, int K, int bn);
int deft_tc_pack_weights(const float* w_packed, int ldw, int K, int bn, int taps /* KH*KW of the conv that will use it (1 for a plain GEMM) */,
                         float* blob, deft_stream_t s);
size_t deft_tc_scratch_bytes(int ldw, int K, int M, int kind);

/* ---- layout converters at the reference's NCHW boundary ------------------------------------ */
int deft_nchw_to_nhwc(const float* src, float* dst, int dst_pitch, int N, int C, int H, int W, deft_stream_t s);
int deft_nhwc_to_nchw(const float* src, int src_pitch, float* dst, int N, int C, int H, int W, deft_stream_t s);

/* ---- a2: nn.Conv2d + eval BatchNorm2d (+ residual) (+ ReLU) over a channel-concat of up to 4
 * sources.  Replaces BasicBlock/Root/project/level convs (dla.py:47-87,184-207,301-398).
 *   out[p,co] = act( scale[co]*sum_k A[p,k]*Wp[k,co] + shift[co] + residual[p,co] )
 * Wp packed [KH*KW*Cin_total][ldw], k = (ky*KW+kx)*Cin_total + ci, ldw = Cout rounded up to 16.
 * Cin of every source must be a multiple of 16.  residual may be NULL.
 * w_tc / tc_bn: the same weights as tcgen05 tile images (deft_tc_pack_weights, kind 0) or NULL; used when
 * DEFT_FLAG_TENSORCORE is set, otherwise (or for uncovered shapes) the fp32 SIMT engine runs.
 * ws: optional split-K scratch (deft_tc_scratch_bytes) or NULL. */
int deft_conv2d_bn_act(const float* const* srcs_host, const int* src_c_host, const int* src_pitch_host, int nsrc,
                       int N, int H, int W, const float* w_packed, int ldw, const float* w_tc, int tc_bn, const float* scale,
                       const float* shift, const float* residual, int res_pitch, float* out, int out_pitch, int Cout,
                       int KH, int KW, int stride, int pad, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- a2: the 7x7 stem on the raw NCHW image (dla.py:301-305): x[N,3,H,W] -> NHWC [N,H,W,16].
 * w_packed [7*7*3][16] (k = (ky*7+kx)*3+ci). */
int deft_stem_conv7x7(const float* x_nchw, int N, int H, int W, const float* w_packed, const float* scale,
                      const float* shift, float* out, int out_pitch, deft_stream_t s);

/* ---- a2: nn.MaxPool2d(2,2) (dla.py:262). H, W are the INPUT dims (even). */
int deft_maxpool2x2(const float* in, int in_pitch, float* out, int out_pitch, int N, int H, int W, int C, deft_stream_t s);

/* ---- a3: dcn_v2.DCN.forward (+ DeformConv's BN/ReLU, dla.py:646-665), i.e. upstream
 * dcn_v2_forward(input,weight,bias,offset,mask,3,3,1,1,1,1,1,1,1) fused with its conv_offset_mask.
 *   om = conv3x3(x; w_om)+b_om ; offset=cat(om[0:9],om[9:18]) (chunk order), mask=sigmoid(om[18:27])
 *   y[p,co] = act( scale[co]*sum_{k,ci} Wp[k*Cin+ci,co]*mask_k(p)*bilinear(x[:,ci]; p+tap_k+off_k(p)) + shift[co] )
 * (bias and BN are folded into scale/shift by the packer; scale=1, shift=bias gives the bare op).
 * w_om_packed [9*Cin][32] (27 real columns), b_om[32]; Wp [9*Cin][ldw].
 * w_om_tc,om_bn / w_tc,tc_bn: tcgen05 tile images of the two weights (kind 0 / kind 1) or NULL.
 * ws: >= deft_dcnv2_workspace_bytes() (holds the [P,32] offset/mask map). */
size_t deft_dcnv2_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int deft_dcnv2_forward(const float* x, int x_pitch, int N, int H, int W, int Cin,
                       const float* w_om_packed, const float* b_om, const float* w_om_tc, int om_bn, const float* w_packed, int ldw,
                       const float* w_tc, int tc_bn, const float* scale, const float* shift, float* y, int y_pitch, int Cout,
                       unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- a4: IDAUp's depthwise ConvTranspose2d(k=2f,s=f,p=f/2,groups=C) fused with the following
 * `layers[i] + layers[i-1]` (dla.py:693-699).  x [N,h,w,C] -> out [N,h*f,w*f,C] = up(x) + skip.
 * w_packed [(2f)*(2f)][C].  skip may be NULL. */
int deft_upsample_add(const float* x, int x_pitch, const float* w_packed, const float* skip, int skip_pitch,
                      float* out, int out_pitch, int N, int h, int w, int C, int f, deft_stream_t s);

/* ---- a5: all CenterNet heads at once (base_model.py:33-68,127-131):
 *   hidden = relu(conv3x3(feat; 64 -> nheads*256) + b1) ; out_h = conv1x1(hidden_h; 256 -> C_h) + b2_h
 * w1_packed [9*64][nheads*256]; w2 [sum C_h][256] row-major; b2 [sum C_h].
 * outs_host[h] -> plain NCHW [N,C_h,H,W].  ws holds the hidden map. */
size_t deft_heads_workspace_bytes(int N, int H, int W, int nheads, int head_conv);
int deft_heads_forward(const float* feat, int feat_pitch, int N, int H, int W, int Cin,
                       const float* w1_packed, const float* w1_tc /* kind 2, bn 256 */, const float* b1, int head_conv,
                       const float* w2, const float* b2, const int* head_c_host, int nheads,
                       float* const* outs_host, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- a6: Detector._sigmoid_output (detector.py:486-494), in place on NCHW heads. dep may be NULL. */
int deft_sigmoid_output(float* hm, size_t n_hm, float* dep, size_t n_dep, float depth_scale, deft_stream_t s);

/* ---- a6: generic_decode = _nms + _topk + gathers (decode.py:102-219, utils.py:69-104).
 * hm [B,C,H,W] already sigmoided.  reg/wh/ltrb_amodal NCHW or NULL.
 * Outputs: scores[B,K] f32, inds[B,K] i64 (y*W+x), clses[B,K] i32, ys/xs[B,K] f32,
 * bboxes[B,K,4] (from wh; or from ltrb_amodal when given, as the reference overwrites it),
 * bboxes_amodal[B,K,4] or NULL.  Tie rule among equal scores: lowest flat index first.
 * K <= 1024. */
size_t deft_decode_workspace_bytes(int B, int C, int H, int W, int K);
int deft_decode(const float* hm, int B, int C, int H, int W, int K, const float* reg, const float* wh,
                const float* ltrb_amodal, float* scores, int64_t* inds, int32_t* clses, float* ys, float* xs,
                float* bboxes, float* bboxes_amodal, void* ws, size_t ws_bytes, deft_stream_t s);
/* _tranpose_and_gather_feat (utils.py:32-36): feat [B,C,H,W], inds [B,K] -> out [B,K,C]. */
int deft_gather_feat(const float* feat, int B, int C, int HW, const int64_t* inds, int K, float* out, deft_stream_t s);

/* ---- boundary shim: utils/image.py:391-412 convert_detection on the device.
 * boxes [N,4] (x1,y1,x2,y2) in a W x H pixel frame -> centres [N,2] = (x1+x2)/W - 1, (y1+y2)/H - 1. */
int deft_boxes_to_centres(const float* boxes, int N, float W, float H, float* centres, deft_stream_t s);

/* ---- a7: AFE_module.forward_feature_extracter (AFE.py:88-92,162-188), evaluated sparsely: the
 * selector conv3x3+ReLU is computed only at the (up to) 4 bilinear corners of every centre.
 * maps_host[k] NHWC [1,H_k,W_k,C_k]; selw_host[k] packed [9*C_k][Co_k]; selb_host[k] [Co_k];
 * centres [N,2] (x,y) in [-1,1] (grid_sample, padding_mode=border); emb [N, sum Co_k].
 * ws: >= deft_afe_extract_workspace_bytes(N) (per-corner partial sums). */
size_t deft_afe_extract_workspace_bytes(int N);
int deft_afe_extract(const float* const* maps_host, const int* pitch_host, const int* C_host, const int* H_host,
                     const int* W_host, const float* const* selw_host, const float* const* selb_host,
                     const int* Co_host, const float* centres, int N, float* emb, int align_corners,
                     void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- a8: AFE_module.forward_stacker_features (AFE.py:110-160) for F memory frames in one call.
 * xp [F,P,Cemb] (rows >= n_f ignored), n_host[F]; xn [P,Cemb] (m live rows).
 * Packed MLP (packing.py: stacker2_bn and the three BNs folded):
 *   w1a,w1b [Cemb][512] + sh1a[512] (all of bias/BN shift), w2 [512][256]+sc2,sh2, w3 [256][128]+sc3,sh3,
 *   w4 [128][64]+b4, w5[64], b5.
 * out [F,P,P+1]: out[f,i,j<m] = max(softmax_row, softmax_col)[i,j]; out[f,i,m] = softmax_row[i,P];
 * the rest of each [P, P+1] block is written as 0 (a reused buffer never shows an earlier frame).  P = max_object (the softmax denominators contain the
 * (P-n)/(P-m) zero entries and the constant-1 row/col exactly as the reference). */
size_t deft_afe_affinity_workspace_bytes(int F, int P, int Cemb);
int deft_afe_affinity(const float* xp, const int* n_host, int F, const float* xn, int m, int P, int Cemb,
                      const float* w1a, const float* w1b, const float* sh1,
                      const float* w2, const float* sc2, const float* sh2,
                      const float* w3, const float* sc3, const float* sh3,
                      const float* w4, const float* b4, const float* w5, float b5,
                      const float* const* tc_host /* {w1a,w1b,w2,w3,w4} tile images or NULL */, const int* tc_bn_host,
                      float* out, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- a8, device-resident form: the first MLP layer is linear in the concatenated pair, so
 * Au' = bnfold(W1a bn(u) + b1) of a MEMORY frame never changes while the frame sits in the recorder's ring
 * (tracker.py:59-90).  deft_afe_project computes Au' and/or Bv' for `rows` embeddings ([rows,512] each, NULL = skip);
 * deft_afe_affinity_cached runs layers 2..5 + the dual softmax against a ring au_mem [F,P,512] whose live row counts
 * n_dev[F] (and optionally the live detection count *m_dev) stay on the device: the pair grid is the fixed
 * F x rows x m_max block (dead rows / columns are computed and ignored by the softmax), so the call is
 * graph-capturable with no host knowledge of the counts.  ws >= deft_afe_pairs_workspace_bytes(F, rows, m_max). */
size_t deft_afe_pairs_workspace_bytes(int F, int rows, int cols);
int deft_afe_project(const float* emb, int rows, int Cemb, const float* w1a, const float* w1b, const float* sh1,
                     const float* const* tc_host, const int* tc_bn_host, float* au, float* bv, unsigned flags, deft_stream_t s);
int deft_afe_affinity_cached(const float* au_mem, int F, int P, int rows, const int* n_dev, const float* bv, int m_max,
                             const int* m_dev, const float* w2, const float* sc2, const float* sh2, const float* w3,
                             const float* sc3, const float* sh3, const float* w4, const float* b4, const float* w5, float b5,
                             const float* const* tc_host, const int* tc_bn_host, float* out, unsigned flags, void* ws,
                             size_t ws_bytes, deft_stream_t s);

/* ---- a9: AFE_module.forward (AFE.py:74-86) after the extractor, eval-mode BN: au/bv [B,P,512] = projected embeddings
 * of the previous / next frame of every batch element (padding rows included, as the reference keeps them) ->
 * out [B,1,P+1,P+1] = relu-MLP scores with the false-object column and row of `false_constant` appended
 * (add_unmatched_dim, AFE.py:215-233).  ws >= deft_afe_pairs_workspace_bytes(B, P, P). */
int deft_afe_forward_scores(const float* au, const float* bv, int B, int P, const float* w2, const float* sc2, const float* sh2,
                            const float* w3, const float* sc3, const float* sh3, const float* w4, const float* b4,
                            const float* w5, float b5, const float* const* tc_host, const int* tc_bn_host, float false_constant,
                            float* out, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s);
/* AFE_module.loss forward (AFE.py:235-328): input/target [B,1,P+1,P+1] fp32, mask0/mask1 [B,1,P+1] uint8 ->
 * out7 = {loss_pre, loss_next, loss_similarity, (loss_pre+loss_next+loss+loss_similarity)/4, accuracy_pre,
 * accuracy_next, (accuracy_pre+accuracy_next)/2}, indexes_pre [B,1,P] int64 (first maximum wins, as torch.max). */
size_t deft_afe_loss_workspace_bytes(int B, int P);
int deft_afe_loss(const float* input, const float* target, const unsigned char* mask0, const unsigned char* mask1, int B, int P,
                  float* out7, long long* indexes_pre, void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- f1: the recorder / track memory on the device (tracker.py:46-137, 219-252, 886-894).
 * State, all caller-owned device memory: clock int[2] (clock[0] = index of the frame being associated),
 * slot_frame int[F] (-1 = empty; frame t lives in slot t % F), counts int[F], memory [F,P,C], au_mem [F,P,512],
 * node_frame/node_id int[T, L] (L = deft_track_nodes_per_track(), oldest first), node_cnt int[T].
 *  deft_memory_push      FeatureRecorder.update's bookkeeping: rows (or *m_dev) embeddings (+ their Au') -> slot
 *                        clock % F, slot_frame/counts updated, clock advanced.
 *  deft_track_similarity STrack.get_similarity for all T tracks: rows of aff [F,P,P+1] named by the track's nodes,
 *                        times the recorder's decay (decay^(d/3) for d < m_frame, else decay2^(d/3)), median over the
 *                        last mm rows (all mm+1 when there are exactly mm+1) -> sim [T,P+1] (live [:, :m+1], tracks
 *                        without a usable node: zeros) and cost [T,P] = 1 - sim[:, :m] (NULL = skip).
 *  deft_track_append     after association: match[t] = detection of track t or -1 -> node (clock, det) appended.
 *  deft_count_above      number of scores > thresh (merge_outputs' cut, detector.py:577-583) as a device count. */
int deft_track_nodes_per_track(void);
int deft_memory_push(const float* emb, const float* au, int rows, const int* m_dev, int C, int F, int P, int* clock,
                     int* slot_frame, int* counts, float* memory, float* au_mem, deft_stream_t s);
int deft_track_similarity(const float* aff, int F, int P, const int* clock, const int* slot_frame, const int* counts,
                          const int* node_frame, const int* node_id, const int* node_cnt, int T, int m, const int* m_dev,
                          int mm, int m_frame, int max_track_node, double decay, double decay2, float* sim, float* cost,
                          deft_stream_t s);
int deft_track_append(const int* match, const int* clock, int T, int* node_frame, int* node_id, int* node_cnt, deft_stream_t s);
int deft_count_above(const float* scores, int K, float thresh, int* out, deft_stream_t s);

/* ---- f3: motion fusion + linear assignment on the device (matching.py:40-55, 311-364; tracker.py:886-915).
 * cost [T, ldc] fp32 (= 1 - similarity, row per track), m live detections (or *m_dev).  Motion inputs (all or none):
 * pred_xy [T,2] the tracks' predicted centres for this frame, n_obs [T] observation counts, det_xy [m,2] detection
 * centres: the reference's LSTM branch of fuse_motion (>= 300 observations: gate at 5*chi2inv95[2] on the squared
 * centre distance, cost = lambda*cost + 0.05*(1-lambda)*d2; younger tracks: lambda*cost, the reference's empty-slice
 * "gaussian" metric).  Assignment = lap.lapjv(cost, extend_cost=True, cost_limit) solved exactly in fp64 (shortest
 * augmenting paths): match_track[T] = detection or -1, match_det[m] = track or -1.  ws >= deft_assign_workspace_bytes. */
size_t deft_assign_workspace_bytes(int T, int m);
int deft_fuse_and_assign(const float* cost, int ldc, int T, int m, const int* m_dev, const float* pred_xy, const int* n_obs,
                         const float* det_xy, double lambda_, double cost_limit, int* match_track, int* match_det, void* ws,
                         size_t ws_bytes, deft_stream_t s);

/* ---- f2: pre-processing (detector.py:377-422).  deft_stem_conv7x7_u8 = deft_stem_conv7x7 reading a uint8 HWC frame
 * [N,H,W,3] already at network resolution, normalisation (u8/255 - mean[c])/std[c] fused into the load (x_nchw: optional
 * fp32 NCHW copy of the normalised frame, NULL = none).  deft_preprocess_warp: general inverse-affine bilinear warp +
 * normalisation of a [src_h,src_w,3] uint8 frame into fp32 NCHW [3,H,W] (inv_affine: host float[6], dst -> src). */
int deft_stem_conv7x7_u8(const unsigned char* img, int N, int H, int W, const float* mean_host, const float* std_host,
                         const float* w_packed, const float* scale, const float* shift, float* out, int out_pitch,
                         float* x_nchw, deft_stream_t s);
int deft_preprocess_warp(const unsigned char* img, int src_h, int src_w, const float* inv_affine_host, const float* mean_host,
                         const float* std_host, float* out_nchw, int H, int W, deft_stream_t s);

/* ---- f4 (training, BASELINE config C5): backward of the DCN module (upstream dcn_v2_backward + the offset/mask conv's own
 * backward), fp32.  Inputs as deft_dcnv2_forward (no BN: the bare module, bias included) plus grad_y [N*H*W, gy_pitch].
 * Outputs: grad_x [N*H*W, gx_pitch] (written), grad_w [9*Cin][ldw] and grad_w_om [9*Cin][32] in the packed layout of
 * the forward weights, grad_b [Cout], grad_b_om [32] (entries >= 27 are 0).  Needs Cout % 16 == 0 (ldw == Cout).
 * grad_x's deformable part is accumulated with float atomics (as upstream's col2im): run-to-run bit differences. */
size_t deft_dcnv2_backward_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int deft_dcnv2_backward(const float* x, int x_pitch, int N, int H, int W, int Cin, const float* w_om_packed, const float* b_om,
                        const float* w_packed, int ldw, int Cout, const float* grad_y, int gy_pitch, float* grad_x, int gx_pitch,
                        float* grad_w_om, float* grad_b_om, float* grad_w, float* grad_b, void* ws, size_t ws_bytes, deft_stream_t s);

/* ---- a10: KalmanFilterLSTM.predict (kalman_filter_lstm.py:65-78) for T independent tracks in one
 * launch: one LSTM step (gate order i,f,g,o) + out1 (128->64, no activation) + out2 (64->nout).
 * x [T,F], h/c [T,128] -> h_out/c_out [T,128], pred [T,nout].  Weights in PyTorch layout
 * (w_ih [512,F], w_hh [512,128], w1 [64,128], w2 [nout,64]). */
int deft_lstm_step(const float* x, const float* h, const float* c, const float* w_ih, const float* w_hh,
                   const float* b_ih, const float* b_hh, const float* w1, const float* b1, const float* w2,
                   const float* b2, int T, int F, int nout, float* h_out, float* c_out, float* pred, deft_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* DEFT_B200_H */
