// Entry points either side of the per-frame dense path:
//   f1  device-resident track memory: frame clock, embedding ring push, per-track node lists and the decayed
//       per-track median over the last <= mm (+1) affinity rows -> ONE [tracks, dets] similarity / cost matrix
//       (reference: FeatureRecorder.update tracker.py:59-90, STrack.get_similarity tracker.py:219-252, 886-894).
//   f2  uint8 HWC frame -> (x/255 - mean)/std fused into the 7x7 stem's load path (detector.py:377-422).
//   a9  AFE_module.loss forward (AFE.py:235-328): masked dual-softmax cross-entropy + consistency + accuracies.
// All SIMT fp32: these are latency / HBM-bound element and reduction kernels, not GEMMs.
#include "common.cuh"
#include <math_constants.h>

namespace deft {

// ================================================================================================ f1
// State (all device-resident, owned by the caller):
//   clock[0]            current frame index (the frame whose detections are being associated); advanced by push
//   slot_frame[F]       frame index held by ring slot s (-1 = empty); a frame lives in slot (frame % F)
//   counts[F]           live rows of the slot
//   memory[F,P,C], au_mem[F,P,512]  embeddings and their cached first-layer products (Au')
//   node_frame[T,L], node_id[T,L], node_cnt[T]   the newest <= L nodes of every track, oldest first
constexpr int TRACK_L = 6;                 // mm + 2 newest nodes decide exactly which rows the reference's median sees

struct TrackSimArgs {
    const float* aff;        // [F, P, P+1]: affinity of the current detections (columns) against ring slot f (rows)
    const int* clock;
    const int* slot_frame;
    const int* counts;
    const int* node_frame;
    const int* node_id;
    const int* node_cnt;
    const int* m_dev;        // live current detections or nullptr (-> m)
    int F, P, T, m, mm, m_frame, max_track_node;
    double decay, decay2;
    float* sim;              // [T, P+1]   (live block [:, :m+1])
    float* cost;             // [T, P] = 1 - sim[:, :m]  or nullptr
};

// One block per track.  Thread-per-column median of the <= 5 selected rows (np.median semantics: mean of the two
// middle values for an even count), after the recorder's decay factor (tracker.py:76-90).
__global__ void __launch_bounds__(128) track_similarity_kernel(TrackSimArgs a) {
    DEFT_PDL_PROLOGUE();
    __shared__ int s_slot[TRACK_L], s_id[TRACK_L];
    __shared__ float s_delta[TRACK_L];
    __shared__ int s_R;
    const int t = blockIdx.x;
    const int m = a.m_dev ? min(max(__ldg(a.m_dev), 0), a.P) : a.m;
    if (threadIdx.x == 0) {
        const int cur = a.clock[0];
        const int cnt = min(a.node_cnt[t], TRACK_L);
        int nv = 0;
        // nodes are stored oldest first; a node counts when its frame is inside the window (tracker.py:225) and the
        // recorder still holds that frame (its ring slot has not been overwritten)
        for (int q = 0; q < cnt; ++q) {
            const int fr = a.node_frame[t * TRACK_L + q], id = a.node_id[t * TRACK_L + q];
            if (fr < 0 || fr >= cur || cur - fr >= a.max_track_node) continue;
            const int slot = fr % a.F;
            if (a.slot_frame[slot] != fr || id < 0 || id >= a.counts[slot]) continue;
            const int d = cur - fr;
            const double delta = (d < a.m_frame) ? pow(a.decay, (double)d / 3.0) : pow(a.decay2, (double)d / 3.0);
            s_slot[nv] = slot; s_id[nv] = id; s_delta[nv] = (float)delta;
            ++nv;
        }
        // tracker.py:236-248: more than mm rows -> the last mm, except that exactly mm+1 rows are all kept
        int R = nv, first = 0;
        if (nv > a.mm + 1) { R = a.mm; first = nv - a.mm; }
        for (int q = 0; q < R; ++q) { s_slot[q] = s_slot[first + q]; s_id[q] = s_id[first + q]; s_delta[q] = s_delta[first + q]; }
        s_R = R;
    }
    __syncthreads();
    const int R = s_R, Q = a.P + 1;
    for (int j = threadIdx.x; j <= a.P; j += blockDim.x) {
        float med = 0.f;
        if (R > 0 && j <= m) {
            float v[TRACK_L];
#pragma unroll
            for (int q = 0; q < TRACK_L; ++q)
                v[q] = q < R ? a.aff[((size_t)s_slot[q] * a.P + s_id[q]) * Q + j] * s_delta[q] : 3.0e38f;
            // insertion sort of <= 6 values (padding sorts last)
#pragma unroll
            for (int x = 1; x < TRACK_L; ++x) {
#pragma unroll
                for (int y = x; y > 0; --y) {
                    const float lo = fminf(v[y - 1], v[y]), hi = fmaxf(v[y - 1], v[y]);
                    v[y - 1] = lo; v[y] = hi;
                }
            }
            float lo = v[0], hi = v[0];
#pragma unroll
            for (int q = 0; q < TRACK_L; ++q) {
                if (q == (R - 1) / 2) lo = v[q];
                if (q == R / 2) hi = v[q];
            }
            med = (R & 1) ? lo : (lo + hi) * 0.5f;
        }
        a.sim[(size_t)t * Q + j] = med;
        if (a.cost != nullptr && j < a.P) a.cost[(size_t)t * a.P + j] = (j < m) ? 1.0f - med : 0.f;
    }
}

// After association: track t matched current detection match[t] (-1 = none) -> append the node (clock, det).
__global__ void track_append_kernel(const int* __restrict__ match, const int* __restrict__ clock, int T, int* __restrict__ node_frame,
                                    int* __restrict__ node_id, int* __restrict__ node_cnt) {
    DEFT_PDL_PROLOGUE();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int j = match[t];
    if (j < 0) return;
    int cnt = node_cnt[t];
    int* nf = node_frame + t * TRACK_L;
    int* ni = node_id + t * TRACK_L;
    if (cnt >= TRACK_L) {
        for (int q = 1; q < TRACK_L; ++q) { nf[q - 1] = nf[q]; ni[q - 1] = ni[q]; }
        cnt = TRACK_L - 1;
    }
    nf[cnt] = clock[0];
    ni[cnt] = j;
    node_cnt[t] = cnt + 1;
}

// FeatureRecorder.update's bookkeeping (tracker.py:59-74) on the device: the current frame's embeddings (and their
// cached first-layer products) go into ring slot clock % F, the slot's frame index / row count are recorded, and the
// clock advances.  One launch, graph-capturable: nothing about the slot is known to the host.
__global__ void memory_push_kernel(const float* __restrict__ emb, const float* __restrict__ au, int rows_max, const int* __restrict__ m_dev, int C,
                                   int F, int P, int* __restrict__ clock, int* __restrict__ slot_frame, int* __restrict__ counts,
                                   float* __restrict__ memory, float* __restrict__ au_mem, int* __restrict__ clock_next) {
    DEFT_PDL_PROLOGUE();
    const int cur = clock[0];
    const int slot = ((cur % F) + F) % F;
    int rows = m_dev ? min(max(__ldg(m_dev), 0), rows_max) : rows_max;
    rows = min(rows, P);
    const size_t n4 = (size_t)rows * (C / 4), a4 = au ? (size_t)rows * 128 : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4 + a4; i += stride) {
        if (i < n4) {
            reinterpret_cast<float4*>(memory + (size_t)slot * P * C)[i] = reinterpret_cast<const float4*>(emb)[i];
        } else {
            const size_t k = i - n4;
            reinterpret_cast<float4*>(au_mem + (size_t)slot * P * 512)[k] = reinterpret_cast<const float4*>(au)[k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        slot_frame[slot] = cur;
        counts[slot] = rows;
        clock_next[0] = cur + 1;       // a separate word: every block of this grid reads clock[0] above
    }
}
__global__ void clock_commit_kernel(int* __restrict__ clock, const int* __restrict__ clock_next) {
    DEFT_PDL_PROLOGUE();
    if (blockIdx.x == 0 && threadIdx.x == 0) clock[0] = clock_next[0];
}

// number of leading detections whose score exceeds a threshold (scores are sorted descending by decode): the device-side
// count that merge_outputs' `score > out_thresh` produces on the host in the reference (detector.py:577-583)
__global__ void count_above_kernel(const float* __restrict__ scores, int K, float thresh, int* __restrict__ out) {
    DEFT_PDL_PROLOGUE();
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < K; i += blockDim.x) c += scores[i] > thresh ? 1 : 0;
    atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) out[0] = cnt;
}

// ================================================================================================ f2
// uint8 HWC (RGB/BGR as the caller has it; channel c is normalised with mean[c], std[c]) -> the 7x7 stem, fused:
// x = (u8 / 255 - mean[c]) / std[c]  exactly as detector.py:408-409 ((resized / 255. - mean) / std in fp32), zero
// padding applies to the NORMALISED image (the conv pads with 0 after normalisation).
constexpr int STEM_TW = 64, STEM_TH = 16, STEM_LD = 72;
__global__ void __launch_bounds__(256) stem7x7_u8_kernel(const unsigned char* __restrict__ img, int H, int W, float m0, float m1, float m2, float s0,
                                                         float s1, float s2, const float* __restrict__ wp, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, float* __restrict__ out, int out_pitch,
                                                         float* __restrict__ x_nchw) {
    DEFT_PDL_PROLOGUE();
    __shared__ __align__(16) float tile[3][STEM_TH + 6][STEM_LD];
    __shared__ __align__(16) float ws[147 * 16];
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * STEM_TW, y0 = blockIdx.y * STEM_TH;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < 147 * 16; i += 256) ws[i] = wp[i];
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    for (int i = tid; i < (STEM_TH + 6) * (STEM_TW + 6); i += 256) {
        const int yy = i / (STEM_TW + 6), xx = i - yy * (STEM_TW + 6);
        const int gy = y0 + yy - 3, gx = x0 + xx - 3;
        float v[3] = {0.f, 0.f, 0.f};
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const unsigned char* p = img + (((size_t)n * H + gy) * W + gx) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = ((float)p[c] / 255.0f - mean[c]) / stdv[c];
            // optional fp32 NCHW copy of the normalised frame (what the reference hands to the network): written by the
            // block that owns the pixel (interior of its tile)
            if (x_nchw != nullptr && yy >= 3 && yy < STEM_TH + 3 && xx >= 3 && xx < STEM_TW + 3) {
#pragma unroll
                for (int c = 0; c < 3; ++c) x_nchw[((size_t)(n * 3 + c) * H + gy) * W + gx] = v[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) tile[c][yy][xx] = v[c];
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[p][o] = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* row = &tile[c][ty + ky][tx * 4];
            float v[12];
            float4 a = *reinterpret_cast<const float4*>(row), b = *reinterpret_cast<const float4*>(row + 4);
            float2 d = *reinterpret_cast<const float2*>(row + 8);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = d.x; v[9] = d.y;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4* w4 = reinterpret_cast<const float4*>(&ws[((ky * 7 + kx) * 3 + c) * 16]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 w = w4[q];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        acc[p][q * 4 + 0] += v[kx + p] * w.x; acc[p][q * 4 + 1] += v[kx + p] * w.y;
                        acc[p][q * 4 + 2] += v[kx + p] * w.z; acc[p][q * 4 + 3] += v[kx + p] * w.w;
                    }
                }
            }
        }
    }
    const int oy = y0 + ty;
    if (oy >= H) return;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int ox = x0 + tx * 4 + p;
        if (ox >= W) break;
        float* o = out + (((size_t)n * H + oy) * W + ox) * out_pitch;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v;
            v.x = fmaxf(acc[p][q * 4 + 0] * scale[q * 4 + 0] + shift[q * 4 + 0], 0.f);
            v.y = fmaxf(acc[p][q * 4 + 1] * scale[q * 4 + 1] + shift[q * 4 + 1], 0.f);
            v.z = fmaxf(acc[p][q * 4 + 2] * scale[q * 4 + 2] + shift[q * 4 + 2], 0.f);
            v.w = fmaxf(acc[p][q * 4 + 3] * scale[q * 4 + 3] + shift[q * 4 + 3], 0.f);
            reinterpret_cast<float4*>(o)[q] = v;
        }
    }
}

// General pre-processing (detector.py:377-409): inverse-affine bilinear warp of a uint8 HWC frame into the network
// resolution + normalisation, fp32 NCHW out.  cv2.warpAffine evaluates the bilinear blend in 1/32-pixel fixed point;
// this kernel interpolates in fp32 at the exact source coordinate (difference <= 1/64 px of gradient, documented in
// INTEGRATION.md); with an identity transform it reproduces (u8/255 - mean)/std exactly.  Border: constant 0 before
// normalisation (cv2 BORDER_CONSTANT, borderValue 0).
__global__ void preprocess_warp_kernel(const unsigned char* __restrict__ img, int sh, int sw, float a00, float a01, float a02, float a10, float a11,
                                       float a12, float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out, int H, int W) {
    DEFT_PDL_PROLOGUE();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int y = i / W, x = i - y * W;
    // dst (x, y) -> src: the caller passes the INVERSE map (what cv2.warpAffine computes internally from trans_input)
    const float sx = a00 * x + a01 * y + a02, sy = a10 * x + a11 * y + a12;
    const float fx = floorf(sx), fy = floorf(sy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float lx = sx - fx, ly = sy - fy;
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
            v[q] = (yy >= 0 && yy < sh && xx >= 0 && xx < sw) ? (float)img[((size_t)yy * sw + xx) * 3 + c] : 0.f;
        }
        const float top = v[0] + lx * (v[1] - v[0]), bot = v[2] + lx * (v[3] - v[2]);
        const float p = top + ly * (bot - top);
        out[((size_t)c * H + y) * W + x] = (p / 255.0f - mean[c]) / stdv[c];
    }
}

// ================================================================================================ a9: loss
// AFE.loss (AFE.py:235-328) forward on input [B,1,Q,Q] (Q = P+1), target [B,1,Q,Q] (0/1), mask0/mask1 [B,1,Q] (0/1).
// Stage 1 (one block per (b, row i) and per (b, column j)): log-sum-exp statistics of the two masked softmaxes:
//   input_pre  = softmax_j(mask_region_pre  * input)   mask_region_pre[i,j]  = mask0[i] mask1[j] [i != P]
//   input_next = softmax_i(mask_region_next * input)   mask_region_next[i,j] = mask0[i] mask1[j] [j != P]
// Stage 2: every term of the losses / accuracies, block partials, then a fixed-order final reduction (deterministic).
struct LossArgs {
    const float* input;
    const float* target;
    const unsigned char* mask0;
    const unsigned char* mask1;
    int B, P;
    float* rmax; float* rsum;     // [B, Q] row statistics of input_pre
    float* cmax; float* csum;     // [B, Q] column statistics of input_next
    double* partial;              // [B*Q, LOSS_TERMS]
    long long* idx_pre;           // [B, P] argmax_j input_all[i, :]   (AFE.py:292-293)
    int* idx_next;                // [B, P] argmax_i input_next[:, j]  (AFE.py:307-308)
    float* out;                   // the 7 scalars of the reference's return tuple
};
constexpr int LOSS_TERMS = 12;

__device__ __forceinline__ float mreg_pre(const LossArgs& a, int b, int i, int j) {
    const int Q = a.P + 1;
    return (i != a.P && a.mask0[b * Q + i] && a.mask1[b * Q + j]) ? 1.f : 0.f;
}
__device__ __forceinline__ float mreg_next(const LossArgs& a, int b, int i, int j) {
    const int Q = a.P + 1;
    return (j != a.P && a.mask0[b * Q + i] && a.mask1[b * Q + j]) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(128) loss_stats_kernel(LossArgs a) {
    DEFT_PDL_PROLOGUE();
    const int Q = a.P + 1;
    const int b = blockIdx.y, r = blockIdx.x % Q, is_col = blockIdx.x / Q;
    const float* in = a.input + (size_t)b * Q * Q;
    __shared__ float red[128];
    float mx = -3.0e38f;
    for (int k = threadIdx.x; k < Q; k += 128) {
        const float v = is_col ? mreg_next(a, b, k, r) * in[k * Q + r] : mreg_pre(a, b, r, k) * in[r * Q + k];
        mx = fmaxf(mx, v);
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
    mx = red[0];
    __syncthreads();
    float s = 0.f;
    for (int k = threadIdx.x; k < Q; k += 128) {
        const float v = is_col ? mreg_next(a, b, k, r) * in[k * Q + r] : mreg_pre(a, b, r, k) * in[r * Q + k];
        s += expf(v - mx);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        if (is_col) { a.cmax[b * Q + r] = mx; a.csum[b * Q + r] = red[0]; }
        else { a.rmax[b * Q + r] = mx; a.rsum[b * Q + r] = red[0]; }
    }
}

// terms: 0 sum(target) 1 sum(target_pre) 2 sum(target_next) 3 sum(target_union) 4 -sum(target_pre log input_pre)
//        5 -sum(target_next log input_next) 6 -sum(target_pre log input_all) 7 sum(target_union |input_next - input_pre|)
//        8 mask_pre_num 9 correct_pre 10 mask_next_num 11 correct_next
__global__ void __launch_bounds__(128) loss_terms_kernel(LossArgs a) {
    DEFT_PDL_PROLOGUE();
    const int Q = a.P + 1, P = a.P;
    const int b = blockIdx.y, i = blockIdx.x;          // one block per (b, row i)
    const float* in = a.input + (size_t)b * Q * Q;
    const float* tg = a.target + (size_t)b * Q * Q;
    double t[LOSS_TERMS];
#pragma unroll
    for (int q = 0; q < LOSS_TERMS; ++q) t[q] = 0.0;
    // argmax over j of input_all[i, :] and of target_pre[i, :] (first maximum wins, like torch.max on the CPU)
    float best_all = -3.0e38f, best_tp = -3.0e38f;
    int arg_all = 0x7fffffff, arg_tp = 0x7fffffff;
    const float rmax = a.rmax[b * Q + i], rsum = a.rsum[b * Q + i];
    for (int j = threadIdx.x; j < Q; j += 128) {
        const float mp = mreg_pre(a, b, i, j), mn = mreg_next(a, b, i, j);
        const float x = in[i * Q + j], tv = tg[i * Q + j];
        const float ip = expf(mp * x - rmax) / rsum;
        const float inx = expf(mn * x - a.cmax[b * Q + j]) / a.csum[b * Q + j];
        const float iall = (i < P && j < P) ? (ip + inx) / 2.0f : ip;
        const float tp = mp * tv, tn = mn * tv, tu = mp * mn * tv;
        t[0] += tv; t[1] += tp; t[2] += tn; t[3] += tu;
        if (tp != 0.f) { t[4] -= (double)(tp * logf(ip)); t[6] -= (double)(tp * logf(iall)); }
        if (tn != 0.f) t[5] -= (double)(tn * logf(inx));
        if (tu != 0.f) t[7] += (double)(tu * fabsf((1.f - ip) - (1.f - inx)));
        if (iall > best_all) { best_all = iall; arg_all = j; }
        if (tp > best_tp) { best_tp = tp; arg_tp = j; }
    }
    __shared__ float s_v[128];
    __shared__ int s_i[128];
    __shared__ double s_d[128];
    // block argmax (value desc, index asc) for input_all, then for target_pre
    int res_all = 0, res_tp = 0;
    for (int pass = 0; pass < 2; ++pass) {
        s_v[threadIdx.x] = pass ? best_tp : best_all;
        s_i[threadIdx.x] = pass ? arg_tp : arg_all;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float v2 = s_v[threadIdx.x + o]; const int i2 = s_i[threadIdx.x + o];
                if (v2 > s_v[threadIdx.x] || (v2 == s_v[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_v[threadIdx.x] = v2; s_i[threadIdx.x] = i2; }
            }
            __syncthreads();
        }
        if (pass) res_tp = s_i[0]; else res_all = s_i[0];
        __syncthreads();
    }
    if (threadIdx.x == 0 && i < P) {
        a.idx_pre[(size_t)b * P + i] = res_all;
        if (a.mask0[b * Q + i]) { t[8] += 1.0; if (res_all == res_tp) t[9] += 1.0; }
    }
    for (int q = 0; q < LOSS_TERMS; ++q) {
        s_d[threadIdx.x] = t[q];
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) { if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) a.partial[((size_t)b * Q + i) * LOSS_TERMS + q] = s_d[0];
        __syncthreads();
    }
}

// accuracy_next (AFE.py:305-317): per column j < P, argmax_i input_next[:, j] against argmax_i target_next[:, j]
__global__ void __launch_bounds__(128) loss_cols_kernel(LossArgs a) {
    DEFT_PDL_PROLOGUE();
    const int Q = a.P + 1, P = a.P;
    const int b = blockIdx.y, j = blockIdx.x;
    const float* in = a.input + (size_t)b * Q * Q;
    const float* tg = a.target + (size_t)b * Q * Q;
    float bi = -3.0e38f, bt = -3.0e38f;
    int ai = 0x7fffffff, at = 0x7fffffff;
    const float cmax = a.cmax[b * Q + j], csum = a.csum[b * Q + j];
    for (int i = threadIdx.x; i < Q; i += 128) {
        const float mn = mreg_next(a, b, i, j);
        const float inx = expf(mn * in[i * Q + j] - cmax) / csum;
        const float tn = mn * tg[i * Q + j];
        if (inx > bi) { bi = inx; ai = i; }
        if (tn > bt) { bt = tn; at = i; }
    }
    __shared__ float s_v[128];
    __shared__ int s_i[128];
    int r0 = 0, r1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
        s_v[threadIdx.x] = pass ? bt : bi;
        s_i[threadIdx.x] = pass ? at : ai;
        __syncthreads();
        for (int o = 64; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float v2 = s_v[threadIdx.x + o]; const int i2 = s_i[threadIdx.x + o];
                if (v2 > s_v[threadIdx.x] || (v2 == s_v[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_v[threadIdx.x] = v2; s_i[threadIdx.x] = i2; }
            }
            __syncthreads();
        }
        if (pass) r1 = s_i[0]; else r0 = s_i[0];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.idx_next[(size_t)b * P + j] = (r0 == r1) ? (r0 | 0x40000000) : r0;     // bit 30: agrees with the target
}

__global__ void __launch_bounds__(256) loss_final_kernel(LossArgs a) {
    DEFT_PDL_PROLOGUE();
    const int Q = a.P + 1, P = a.P;
    __shared__ double s_d[256];
    __shared__ double tot[LOSS_TERMS];
    const int rows = a.B * Q;
    for (int q = 0; q < LOSS_TERMS; ++q) {
        double s = 0.0;
        for (int r = threadIdx.x; r < rows; r += 256) s += a.partial[(size_t)r * LOSS_TERMS + q];
        s_d[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) tot[q] = s_d[0];
        __syncthreads();
    }
    // accuracy_next terms from the column pass
    double mnum = 0.0, corr = 0.0;
    for (int r = threadIdx.x; r < a.B * P; r += 256) {
        const int b = r / P, j = r - b * P;
        if (a.mask1[b * Q + j]) { mnum += 1.0; if (a.idx_next[r] & 0x40000000) corr += 1.0; }
    }
    s_d[threadIdx.x] = mnum;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) tot[10] = s_d[0];
    __syncthreads();
    s_d[threadIdx.x] = corr;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s_d[threadIdx.x] += s_d[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        tot[11] = s_d[0];
        const bool npre = (long long)tot[1] != 0, nnext = (long long)tot[2] != 0, nuni = (long long)tot[3] != 0;
        const float loss_pre = (float)(npre ? tot[4] / tot[1] : tot[4]);
        const float loss_next = (float)(nnext ? tot[5] / tot[2] : tot[5]);
        const float loss = (float)((npre && nnext) ? tot[6] / tot[1] : tot[6]);
        const float loss_sim = (float)(nuni ? tot[7] / tot[0] : tot[7]);
        const float acc_pre = (float)(tot[8] != 0.0 ? tot[9] / tot[8] : tot[9] + 1.0);
        const float acc_next = (float)(tot[10] != 0.0 ? tot[11] / tot[10] : tot[11] + 1.0);
        a.out[0] = loss_pre; a.out[1] = loss_next; a.out[2] = loss_sim;
        a.out[3] = (loss_pre + loss_next + loss + loss_sim) / 4.0f;
        a.out[4] = acc_pre; a.out[5] = acc_next; a.out[6] = (acc_pre + acc_next) / 2.0f;
    }
}


// ================================================================================================ f3
// Motion fusion (matching.py:311-364, the LSTM branch the tracker takes: tracker.py:903-913) and the linear assignment
// (matching.py:40-55: lap.lapjv(cost, extend_cost=True, cost_limit=thresh)) on the device.
//
// fuse_motion, use_lstm=True, only_position=True: for a track with >= 300 observations the gate is the squared
// distance between the detection centre and the track's predicted centre (covariance = identity, tracker.py:200):
// cost = inf where d2 > 5 * chi2inv95[2], then cost = lambda * cost + 0.05 * (1 - lambda) * d2.  For younger tracks the
// reference asks for metric="gaussian", whose slice [3:-1] of a 2-vector is empty (kalman_filter_lstm.py:91-94), so the
// gating distance is identically 0 and the row is just lambda * cost.  Both are restated, quirk included.
//
// lapjv with extend_cost/cost_limit solves  min sum_matched c_ij + (limit/2) (#unmatched rows + #unmatched cols)
// == every detection is either matched to a distinct track at c - limit or stays unmatched at 0.  Solved exactly in
// fp64 by shortest augmenting paths with potentials (one detection row at a time; the column scan and the arg-min of
// every Dijkstra step run across the block), which visits the same optimum as lapjv (unique for generic costs).
__global__ void fuse_motion_kernel(const float* __restrict__ cost, int ldc, int T, int m, const int* __restrict__ m_dev, const float* __restrict__ pred_xy,
                                   const int* __restrict__ n_obs, const float* __restrict__ det_xy, double lambda, double gate, int min_obs,
                                   double limit, double* __restrict__ out, int ldT) {
    DEFT_PDL_PROLOGUE();
    const int mm = m_dev ? min(max(__ldg(m_dev), 0), m) : m;
    const size_t total = (size_t)mm * T;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / T), t = (int)(idx - (size_t)i * T);
        double c = (double)cost[(size_t)t * ldc + i];
        if (pred_xy != nullptr) {
            double d2 = 0.0;
            if (n_obs[t] >= min_obs) {
                const double dx = (double)det_xy[2 * i] - (double)pred_xy[2 * t], dy = (double)det_xy[2 * i + 1] - (double)pred_xy[2 * t + 1];
                d2 = dx * dx + dy * dy;
                if (d2 > gate) c = CUDART_INF;
                c = lambda * c + 0.05 * (1.0 - lambda) * d2;
            } else {
                c = lambda * c + 0.0005 * (1.0 - lambda) * 0.0;
            }
        }
        out[(size_t)i * ldT + t] = c - limit;             // shifted: matching is worth it iff < 0 (a dummy costs 0)
    }
}

__global__ void __launch_bounds__(1024) assign_kernel(const double* __restrict__ cs, int ldT, int T, int m, const int* __restrict__ m_dev,
                                                      int* __restrict__ match_track, int* __restrict__ match_det) {
    DEFT_PDL_PROLOGUE();
    extern __shared__ unsigned char asm_[];
    const int mm = m_dev ? min(max(__ldg(m_dev), 0), m) : m;
    double* v = reinterpret_cast<double*>(asm_);
    double* minv = v + T;
    double* u = minv + T;
    int* p = reinterpret_cast<int*>(u + m);
    int* way = p + T;
    unsigned char* used = reinterpret_cast<unsigned char*>(way + T);
    __shared__ double r_val[32];
    __shared__ int r_idx[32];
    __shared__ double s_delta, s_minvD, s_ui0;
    __shared__ int s_j1, s_wayD, s_j0, s_done, s_isD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int t = tid; t < T; t += blockDim.x) { v[t] = 0.0; p[t] = -1; }
    for (int i = tid; i < m; i += blockDim.x) u[i] = 0.0;
    __syncthreads();
    for (int i0 = 0; i0 < mm; ++i0) {
        for (int t = tid; t < T; t += blockDim.x) { minv[t] = CUDART_INF; used[t] = 0; way[t] = -1; }
        if (tid == 0) { s_minvD = CUDART_INF; s_wayD = -2; s_j0 = -1; s_done = 0; s_ui0 = 0.0; }
        __syncthreads();
        while (true) {
            const int j0 = s_j0;
            const int i = (j0 < 0) ? i0 : p[j0];
            const double ui = u[i] + ((i == i0) ? 0.0 : 0.0);
            // candidates of row i: every unused track column, and the row's private dummy (cost 0, potential 0)
            double best = CUDART_INF;
            int bidx = 0x7fffffff;
            const double* row = cs + (size_t)i * ldT;
            for (int t = tid; t < T; t += blockDim.x) {
                if (used[t]) continue;
                const double cur = row[t] - ui - v[t];
                if (cur < minv[t]) { minv[t] = cur; way[t] = j0; }
                if (minv[t] < best) { best = minv[t]; bidx = t; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            if (lane == 0) { r_val[warp] = best; r_idx[warp] = bidx; }
            __syncthreads();
            if (tid == 0) {
                double b = CUDART_INF; int bi = 0x7fffffff;
                for (int w = 0; w < nw; ++w) if (r_val[w] < b || (r_val[w] == b && r_idx[w] < bi)) { b = r_val[w]; bi = r_idx[w]; }
                const double candD = 0.0 - ui;
                if (candD < s_minvD) { s_minvD = candD; s_wayD = j0; }
                if (s_minvD <= b) { s_delta = s_minvD; s_isD = 1; s_j1 = -1; }     // the dummy wins ties: no match at cost == limit
                else { s_delta = b; s_isD = 0; s_j1 = bi; }
            }
            __syncthreads();
            const double delta = s_delta;
            // potentials: visited columns (and the virtual start column holding i0) move by delta, the rest get closer
            for (int t = tid; t < T; t += blockDim.x) {
                if (used[t]) { u[p[t]] += delta; v[t] -= delta; }
                else minv[t] -= delta;
            }
            __syncthreads();
            if (tid == 0) {
                u[i0] += delta;
                s_minvD -= delta;
                if (s_isD) s_done = 1;
                else {
                    used[s_j1] = 1;
                    s_j0 = s_j1;
                    if (p[s_j1] < 0) s_done = 2;
                }
            }
            __syncthreads();
            if (s_done) break;
        }
        if (tid == 0) {
            // augment: shift the assignments back along the path; a path that ends in a dummy leaves its last row unmatched
            int j = (s_done == 1) ? s_wayD : s_j0;
            while (j >= 0) {
                const int jp = way[j];
                p[j] = (jp < 0) ? i0 : p[jp];
                j = jp;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < m; i += blockDim.x) match_det[i] = -1;
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) {
        match_track[t] = p[t];
        if (p[t] >= 0) match_det[p[t]] = t;
    }
}

}  // namespace deft

using namespace deft;

extern "C" {

int deft_track_nodes_per_track(void) { return TRACK_L; }

int deft_track_similarity(const float* aff, int F, int P, const int* clock, const int* slot_frame, const int* counts, const int* node_frame,
                          const int* node_id, const int* node_cnt, int T, int m, const int* m_dev, int mm, int m_frame, int max_track_node,
                          double decay, double decay2, float* sim, float* cost, deft_stream_t s) {
    DEFT_REQUIRE(aff && clock && slot_frame && counts && node_frame && node_id && node_cnt && sim, "track_similarity: null argument");
    DEFT_REQUIRE(F >= 1 && P >= 1 && T >= 1 && m >= 0 && m <= P && mm >= 1 && mm + 2 <= TRACK_L, "track_similarity: need mm+2 <= %d, 0<=m<=P", TRACK_L);
    TrackSimArgs a{aff, clock, slot_frame, counts, node_frame, node_id, node_cnt, m_dev, F, P, T, m, mm, m_frame, max_track_node, decay, decay2, sim, cost};
    launch_pdl(track_similarity_kernel, dim3(T), dim3(128), 0, to_stream(s), a);
    DEFT_CHECK_LAUNCH("track_similarity");
    return DEFT_OK;
}

int deft_track_append(const int* match, const int* clock, int T, int* node_frame, int* node_id, int* node_cnt, deft_stream_t s) {
    DEFT_REQUIRE(match && clock && node_frame && node_id && node_cnt && T >= 1, "track_append: bad args");
    launch_pdl(track_append_kernel, dim3(ceil_div(T, 128)), dim3(128), 0, to_stream(s), match, clock, T, node_frame, node_id, node_cnt);
    DEFT_CHECK_LAUNCH("track_append");
    return DEFT_OK;
}

int deft_memory_push(const float* emb, const float* au, int rows, const int* m_dev, int C, int F, int P, int* clock /* [2]: now, scratch */,
                     int* slot_frame, int* counts, float* memory, float* au_mem, deft_stream_t s) {
    DEFT_REQUIRE(emb && clock && slot_frame && counts && memory && rows >= 0 && rows <= P && C % 4 == 0 && F >= 1 && (!au || au_mem),
                 "memory_push: bad args");
    launch_pdl(memory_push_kernel, dim3(64), dim3(256), 0, to_stream(s), emb, au, rows, m_dev, C, F, P, clock, slot_frame, counts, memory, au_mem, clock + 1);
    DEFT_CHECK_LAUNCH("memory_push");
    launch_pdl(clock_commit_kernel, dim3(1), dim3(32), 0, to_stream(s), clock, (const int*)(clock + 1));
    DEFT_CHECK_LAUNCH("clock_commit");
    return DEFT_OK;
}

int deft_count_above(const float* scores, int K, float thresh, int* out, deft_stream_t s) {
    DEFT_REQUIRE(scores && out && K >= 1, "count_above: bad args");
    launch_pdl(count_above_kernel, dim3(1), dim3(256), 0, to_stream(s), scores, K, thresh, out);
    DEFT_CHECK_LAUNCH("count_above");
    return DEFT_OK;
}

size_t deft_assign_workspace_bytes(int T, int m) { return align_up((size_t)m * ((T + 3) / 4 * 4) * sizeof(double), 256); }

int deft_fuse_and_assign(const float* cost, int ldc, int T, int m, const int* m_dev, const float* pred_xy, const int* n_obs, const float* det_xy,
                         double lambda, double cost_limit, int* match_track, int* match_det, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(cost && match_track && match_det && ws && T >= 1 && T <= 4096 && m >= 1 && m <= 1024 && ldc >= m, "fuse_and_assign: need 1<=T<=4096, 1<=m<=1024");
    DEFT_REQUIRE((pred_xy == nullptr) == (n_obs == nullptr) && (pred_xy == nullptr) == (det_xy == nullptr), "fuse_and_assign: motion inputs come together");
    if (ws_bytes < deft_assign_workspace_bytes(T, m)) { set_error("fuse_and_assign: workspace too small"); return DEFT_EWORKSPACE; }
    const int ldT = (T + 3) / 4 * 4;
    double* cs = reinterpret_cast<double*>(ws);
    cudaStream_t st = to_stream(s);
    const double gate = 5.0 * 5.9915;                     // 5 * chi2inv95[2] (kalman_filter.py:13-15, matching.py:326,343)
    int blocks = ceil_div(T * m, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_pdl(fuse_motion_kernel, dim3(blocks), dim3(256), 0, st, cost, ldc, T, m, m_dev, pred_xy, n_obs, det_xy, lambda, gate, 300, cost_limit, cs, ldT);
    DEFT_CHECK_LAUNCH("fuse_motion");
    const size_t smem = (size_t)T * (8 + 8 + 4 + 4 + 1) + (size_t)m * 8 + 64;
    static size_t granted[DEFT_MAX_DEVICES] = {};
    int rc = ensure_dyn_smem(assign_kernel, smem, granted, "assign_kernel");
    if (rc != DEFT_OK) return rc;
    launch_pdl(assign_kernel, dim3(1), dim3(1024), smem, st, (const double*)cs, ldT, T, m, m_dev, match_track, match_det);
    DEFT_CHECK_LAUNCH("assign");
    return DEFT_OK;
}

int deft_stem_conv7x7_u8(const unsigned char* img, int N, int H, int W, const float* mean, const float* std, const float* w_packed,
                         const float* scale, const float* shift, float* out, int out_pitch, float* x_nchw, deft_stream_t s) {
    DEFT_REQUIRE(img && mean && std && w_packed && scale && shift && out && out_pitch >= 16 && out_pitch % 4 == 0, "stem_u8: bad args");
    dim3 grid(ceil_div(W, STEM_TW), ceil_div(H, STEM_TH), N), block(256);
    launch_pdl(stem7x7_u8_kernel, grid, block, 0, to_stream(s), img, H, W, mean[0], mean[1], mean[2], std[0], std[1], std[2], w_packed, scale, shift,
               out, out_pitch, x_nchw);
    DEFT_CHECK_LAUNCH("stem7x7_u8");
    return DEFT_OK;
}

int deft_preprocess_warp(const unsigned char* img, int src_h, int src_w, const float* inv_affine /* host [6] */, const float* mean, const float* std,
                         float* out_nchw, int H, int W, deft_stream_t s) {
    DEFT_REQUIRE(img && inv_affine && mean && std && out_nchw && H > 0 && W > 0 && src_h > 0 && src_w > 0, "preprocess_warp: bad args");
    launch_pdl(preprocess_warp_kernel, dim3(ceil_div(H * W, 256)), dim3(256), 0, to_stream(s), img, src_h, src_w, inv_affine[0], inv_affine[1],
               inv_affine[2], inv_affine[3], inv_affine[4], inv_affine[5], mean[0], mean[1], mean[2], std[0], std[1], std[2], out_nchw, H, W);
    DEFT_CHECK_LAUNCH("preprocess_warp");
    return DEFT_OK;
}

size_t deft_afe_loss_workspace_bytes(int B, int P) {
    const size_t Q = (size_t)P + 1;
    return align_up(4 * (size_t)B * Q * sizeof(float), 256) + align_up((size_t)B * Q * LOSS_TERMS * sizeof(double), 256) +
           align_up((size_t)B * P * sizeof(int), 256);
}

int deft_afe_loss(const float* input, const float* target, const unsigned char* mask0, const unsigned char* mask1, int B, int P, float* out7,
                  long long* indexes_pre, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(input && target && mask0 && mask1 && out7 && indexes_pre && ws && B >= 1 && P >= 1, "afe_loss: bad args");
    if (ws_bytes < deft_afe_loss_workspace_bytes(B, P)) { set_error("afe_loss: workspace too small"); return DEFT_EWORKSPACE; }
    const size_t Q = (size_t)P + 1;
    char* base = reinterpret_cast<char*>(ws);
    LossArgs a{};
    a.input = input; a.target = target; a.mask0 = mask0; a.mask1 = mask1; a.B = B; a.P = P;
    float* st = reinterpret_cast<float*>(base);
    a.rmax = st; a.rsum = st + B * Q; a.cmax = st + 2 * B * Q; a.csum = st + 3 * B * Q;
    size_t off = align_up(4 * (size_t)B * Q * sizeof(float), 256);
    a.partial = reinterpret_cast<double*>(base + off);
    off += align_up((size_t)B * Q * LOSS_TERMS * sizeof(double), 256);
    a.idx_next = reinterpret_cast<int*>(base + off);
    a.idx_pre = indexes_pre;
    a.out = out7;
    cudaStream_t cs = to_stream(s);
    launch_pdl(loss_stats_kernel, dim3(2 * (unsigned)Q, B), dim3(128), 0, cs, a);
    DEFT_CHECK_LAUNCH("loss_stats");
    launch_pdl(loss_terms_kernel, dim3((unsigned)Q, B), dim3(128), 0, cs, a);
    DEFT_CHECK_LAUNCH("loss_terms");
    launch_pdl(loss_cols_kernel, dim3(P, B), dim3(128), 0, cs, a);
    DEFT_CHECK_LAUNCH("loss_cols");
    launch_pdl(loss_final_kernel, dim3(1), dim3(256), 0, cs, a);
    DEFT_CHECK_LAUNCH("loss_final");
    return DEFT_OK;
}

}  // extern "C"
