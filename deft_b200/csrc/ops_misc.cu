// Non-conv entry points: sigmoid_output, CenterNet decode (NMS + exact two-stage top-K + gathers),
// AFE pairwise-MLP affinity + dual softmax, batched LSTM step.  (The AFE extractor lives in ops_afe.cu.)
#include "igemm_simt.cuh"
#include "tc_engine.cuh"

namespace deft {

// ================================================================================================
// a6: sigmoid_output
__global__ void sigmoid_output_kernel(float* hm, size_t n_hm, float* dep, size_t n_dep, float depth_scale) {
    DEFT_PDL_PROLOGUE();
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = i; j < n_hm; j += stride) hm[j] = sigmoidf_(hm[j]);
    for (size_t j = i; j < n_dep; j += stride) dep[j] = (1.0f / (sigmoidf_(dep[j]) + 1e-6f) - 1.0f) * depth_scale;
}

// ================================================================================================
// a6: decode
__device__ __forceinline__ uint32_t f2u(float f) {
    f = f + 0.0f;  // -0 -> +0
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float u2f(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
    return __uint_as_float(b);
}

// heat * (maxpool3x3(heat) == heat) for one element (utils.py:69-74)
__device__ __forceinline__ float nms_value(const float* __restrict__ pl, int p, int H, int W) {
    int y = p / W, x = p - y * W;
    float h = pl[p], m = h;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) m = fmaxf(m, pl[yy * W + xx]);
        }
    return (m == h) ? h : 0.0f;
}

__device__ void bitonic_sort_desc(unsigned long long* a, int n /* pow2 */) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n; t += blockDim.x) {
                int ixj = t ^ j;
                if (ixj > t) {
                    unsigned long long x = a[t], y = a[ixj];
                    bool desc = (t & k) == 0;
                    if (desc ? (x < y) : (x > y)) { a[t] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// Stage A: every block takes a 4096-element chunk of one (class, batch) plane, applies the 3x3 NMS on the fly and
// keeps the chunk's top-K by (value desc, index asc).  The global top-K is a subset of the union of chunk top-Ks.
constexpr int TOPK_CHUNK = 4096;
__global__ void __launch_bounds__(1024) topk_chunk_kernel(const float* __restrict__ hm, int H, int W, int K, int nchunks,
                                                          unsigned long long* __restrict__ cand) {
    DEFT_PDL_PROLOGUE();
    __shared__ unsigned long long keys[TOPK_CHUNK];
    const int HW = H * W;
    const size_t plane = (size_t)blockIdx.z * gridDim.y + blockIdx.y;
    const float* pl = hm + plane * HW;
    for (int e = threadIdx.x; e < TOPK_CHUNK; e += 1024) {
        int i = blockIdx.x * TOPK_CHUNK + e;
        unsigned long long key = 0ull;
        if (i < HW) key = ((unsigned long long)f2u(nms_value(pl, i, H, W)) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
        keys[e] = key;
    }
    __syncthreads();
    bitonic_sort_desc(keys, TOPK_CHUNK);
    unsigned long long* out = cand + (plane * nchunks + blockIdx.x) * K;
    for (int i = threadIdx.x; i < K; i += 1024) out[i] = keys[i];
}

// Stage B: per (class, batch) merge of the chunk winners -> the per-class top-K of utils.py:92.
__global__ void __launch_bounds__(1024) topk_class_merge_kernel(const unsigned long long* __restrict__ cand_chunk, int n, int Np, int K,
                                                                unsigned long long* __restrict__ cand) {
    DEFT_PDL_PROLOGUE();
    extern __shared__ unsigned long long mkeys[];
    const size_t plane = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    const unsigned long long* src = cand_chunk + plane * n;
    for (int i = threadIdx.x; i < Np; i += 1024) mkeys[i] = (i < n) ? src[i] : 0ull;
    __syncthreads();
    bitonic_sort_desc(mkeys, Np);
    for (int i = threadIdx.x; i < K; i += 1024) cand[plane * K + i] = mkeys[i];
}

struct DecodeArgs {
    const unsigned long long* cand;
    const float* reg;
    const float* wh;
    const float* ltrb;
    float* scores; long long* inds; int* clses; float* ys; float* xs; float* bboxes; float* bboxes_amodal;
    int C, H, W, K;
};

// Second stage (utils.py:98-104) + gathers + box arithmetic (decode.py:113-183).  One block per image.
__global__ void __launch_bounds__(1024) decode_merge_kernel(DecodeArgs a, int Np) {
    DEFT_PDL_PROLOGUE();
    extern __shared__ unsigned long long skeys[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int CK = a.C * a.K, HW = a.H * a.W;
    const unsigned long long* cand = a.cand + (size_t)b * CK;
    for (int i = tid; i < Np; i += blockDim.x) {
        unsigned long long k = 0ull;
        if (i < CK) k = (cand[i] & 0xFFFFFFFF00000000ull) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
        skeys[i] = k;
    }
    __syncthreads();
    bitonic_sort_desc(skeys, Np);
    for (int k = tid; k < a.K; k += blockDim.x) {
        unsigned int flatpos = 0xFFFFFFFFu - (unsigned int)(skeys[k] & 0xFFFFFFFFull);
        int cls = (int)(flatpos / (unsigned int)a.K);
        unsigned int idx = 0xFFFFFFFFu - (unsigned int)(cand[flatpos] & 0xFFFFFFFFull);
        float score = u2f((unsigned int)(cand[flatpos] >> 32));
        int yi = (int)(idx / (unsigned int)a.W), xi = (int)(idx - (unsigned int)yi * a.W);
        float ys0 = (float)yi, xs0 = (float)xi;
        size_t o = (size_t)b * a.K + k;
        a.scores[o] = score; a.inds[o] = (long long)idx; a.clses[o] = cls; a.ys[o] = ys0; a.xs[o] = xs0;
        float xs = __fadd_rn(xs0, 0.5f), ys = __fadd_rn(ys0, 0.5f);
        if (a.reg) {
            xs = __fadd_rn(xs0, a.reg[((size_t)b * 2 + 0) * HW + idx]);
            ys = __fadd_rn(ys0, a.reg[((size_t)b * 2 + 1) * HW + idx]);
        }
        if (a.wh && a.bboxes) {
            float w0 = a.wh[((size_t)b * 2 + 0) * HW + idx], w1 = a.wh[((size_t)b * 2 + 1) * HW + idx];
            if (w0 < 0.f) w0 = 0.f;
            if (w1 < 0.f) w1 = 0.f;
            float hw0 = __fmul_rn(w0, 0.5f), hw1 = __fmul_rn(w1, 0.5f);
            a.bboxes[o * 4 + 0] = __fsub_rn(xs, hw0); a.bboxes[o * 4 + 1] = __fsub_rn(ys, hw1);
            a.bboxes[o * 4 + 2] = __fadd_rn(xs, hw0); a.bboxes[o * 4 + 3] = __fadd_rn(ys, hw1);
        }
        if (a.ltrb) {
            float l0 = a.ltrb[((size_t)b * 4 + 0) * HW + idx], l1 = a.ltrb[((size_t)b * 4 + 1) * HW + idx];
            float l2 = a.ltrb[((size_t)b * 4 + 2) * HW + idx], l3 = a.ltrb[((size_t)b * 4 + 3) * HW + idx];
            float bb0 = __fadd_rn(xs0, l0), bb1 = __fadd_rn(ys0, l1), bb2 = __fadd_rn(xs0, l2), bb3 = __fadd_rn(ys0, l3);
            if (a.bboxes_amodal) { a.bboxes_amodal[o * 4 + 0] = bb0; a.bboxes_amodal[o * 4 + 1] = bb1; a.bboxes_amodal[o * 4 + 2] = bb2; a.bboxes_amodal[o * 4 + 3] = bb3; }
            if (a.bboxes) { a.bboxes[o * 4 + 0] = bb0; a.bboxes[o * 4 + 1] = bb1; a.bboxes[o * 4 + 2] = bb2; a.bboxes[o * 4 + 3] = bb3; }
        }
    }
}

__global__ void gather_feat_kernel(const float* __restrict__ feat, int C, int HW, const long long* __restrict__ inds, int K,
                                   float* __restrict__ out, int total) {
    DEFT_PDL_PROLOGUE();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = i % C;
    int k = (i / C) % K;
    int b = i / (C * K);
    out[i] = feat[((size_t)b * C + c) * HW + inds[(size_t)b * K + k]];
}

__global__ void boxes_to_centres_kernel(const float* __restrict__ b, int N, float W, float H, float* __restrict__ c) {
    DEFT_PDL_PROLOGUE();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    // convert_detection: w = x2-x1; x1/=W; w/=W; c = 2*x1 + w - 1
    float x1 = b[i * 4 + 0], y1 = b[i * 4 + 1], x2 = b[i * 4 + 2], y2 = b[i * 4 + 3];
    c[i * 2 + 0] = (2.f * (x1 / W) + (x2 - x1) / W) - 1.0f;
    c[i * 2 + 1] = (2.f * (y1 / H) + (y2 - y1) / H) - 1.0f;
}

// ================================================================================================
// a8: affinity tail.  S[p] = relu(h4[p,:64] . w5 + b5); then the dual softmax.
__global__ void affinity_final_kernel(const float* __restrict__ h4, const float* __restrict__ w5, float b5, float* __restrict__ S, int M) {
    DEFT_PDL_PROLOGUE();
    int p = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);   // 8 lanes per pair
    int l = threadIdx.x & 7;
    float s = 0.f;
    if (p < M) {
        const float* r = h4 + (size_t)p * 64 + l * 8;
        float4 a = ldg4(r), b = ldg4(r + 4), wa = ldg4(w5 + l * 8), wb = ldg4(w5 + l * 8 + 4);
        s = a.x * wa.x + a.y * wa.y + a.z * wa.z + a.w * wa.w + b.x * wb.x + b.y * wb.y + b.z * wb.z + b.w * wb.w;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (p < M && l == 0) S[p] = fmaxf(s + b5, 0.f);
}

struct SoftmaxArgs {
    int n[AFF_MAX_FRAMES];
    int off[AFF_MAX_FRAMES + 1];
    const int* n_dev;    // [F] live rows per memory frame on the device (overrides n[]) or nullptr
    const int* m_dev;    // live current detections on the device (overrides m) or nullptr
    int m, ms, P;        // m: live columns; ms: row stride of S (the pair grid's column count)
};
// One block per memory frame.  X = [[S (n x m), 0...; 0...], ones row/col] of size (P+1)^2 as in
// AFE.py:123-134; rows/cols beyond n/m are exact zeros and enter the denominators analytically.
// The whole [P, P+1] block of `out` is written (zeros outside the live [n, m+1] corner), so a reused buffer never
// shows values of an earlier, larger frame.
__global__ void __launch_bounds__(256) affinity_softmax_kernel(const float* __restrict__ S, SoftmaxArgs a, float* __restrict__ out) {
    DEFT_PDL_PROLOGUE();
    extern __shared__ float sm[];
    const int P = a.P, ms = a.ms, f = blockIdx.x;
    int m = a.m_dev ? __ldg(a.m_dev) : a.m;
    int n = a.n_dev ? __ldg(a.n_dev + f) : a.n[f];
    m = min(max(m, 0), min(ms, P));
    n = min(max(n, 0), P);
    float* rmax = sm; float* rsum = sm + P; float* cmax = sm + 2 * P; float* csum = sm + 3 * P;
    const float* Sf = S + a.off[f];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < n; i += 8) {
        float mx = 1.0f;                       // the appended false-object column
        if (m < P) mx = fmaxf(mx, 0.f);
        for (int j = lane; j < m; j += 32) mx = fmaxf(mx, Sf[i * ms + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int j = lane; j < m; j += 32) s += expf(Sf[i * ms + j] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { rmax[i] = mx; rsum[i] = s + (float)(P - m) * expf(0.f - mx) + expf(1.0f - mx); }
    }
    for (int j = tid; j < m; j += 256) {
        float mx = 1.0f;
        if (n < P) mx = fmaxf(mx, 0.f);
        for (int i = 0; i < n; ++i) mx = fmaxf(mx, Sf[i * ms + j]);
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += expf(Sf[i * ms + j] - mx);
        cmax[j] = mx;
        csum[j] = s + (float)(P - n) * expf(0.f - mx) + expf(1.0f - mx);
    }
    __syncthreads();
    float* of = out + (size_t)f * P * (P + 1);
    for (int idx = tid; idx < P * (P + 1); idx += 256) {
        int i = idx / (P + 1), j = idx - i * (P + 1);
        float v = 0.f;
        if (i < n && j < m) {
            float s = Sf[i * ms + j];
            v = fmaxf(expf(s - rmax[i]) / rsum[i], expf(s - cmax[j]) / csum[j]);
        } else if (i < n && j == m) {
            v = expf(1.0f - rmax[i]) / rsum[i];
        }
        of[idx] = v;
    }
}

// a9 (AFE.py:74-86 + add_unmatched_dim 215-233): the raw score matrix of every batch element with the false-object
// column and row of ones appended: out[b] = [[S_b (P x P), 1], [1 ... 1]]  ((P+1)^2, no softmax, no zeroing).
__global__ void affinity_scores_kernel(const float* __restrict__ S, int P, float false_constant, float* __restrict__ out) {
    DEFT_PDL_PROLOGUE();
    const int b = blockIdx.y, Q = P + 1;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < Q * Q; idx += gridDim.x * blockDim.x) {
        int i = idx / Q, j = idx - i * Q;
        out[(size_t)b * Q * Q + idx] = (i < P && j < P) ? S[(size_t)b * P * P + i * P + j] : false_constant;
    }
}

// ================================================================================================
// a10: batched single LSTM step + out1 + out2.  One block (128 threads) per track.
__global__ void __launch_bounds__(128) lstm_step_kernel(const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ c,
                                                        const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                        const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2, int F, int nout,
                                                        float* __restrict__ h_out, float* __restrict__ c_out, float* __restrict__ pred) {
    __shared__ float gates[512];
    __shared__ __align__(16) float hn[128];
    __shared__ __align__(16) float o1[64];
    DEFT_PDL_PROLOGUE();
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4 hv = ldg4(h + (size_t)t * 128 + lane * 4);
    const float xv = (lane < F) ? x[(size_t)t * F + lane] : 0.f;
    for (int r = warp * 128; r < warp * 128 + 128; ++r) {
        float4 w = ldg4(w_hh + (size_t)r * 128 + lane * 4);
        float s = w.x * hv.x + w.y * hv.y + w.z * hv.z + w.w * hv.w;
        if (lane < F) s += __ldg(w_ih + (size_t)r * F + lane) * xv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) gates[r] = s + b_ih[r] + b_hh[r];
    }
    __syncthreads();
    {
        float ig = sigmoidf_(gates[tid]), fg = sigmoidf_(gates[128 + tid]);
        float gg = tanhf(gates[256 + tid]), og = sigmoidf_(gates[384 + tid]);
        float cn = fg * c[(size_t)t * 128 + tid] + ig * gg;
        float hh = og * tanhf(cn);
        c_out[(size_t)t * 128 + tid] = cn;
        h_out[(size_t)t * 128 + tid] = hh;
        hn[tid] = hh;
    }
    __syncthreads();
    {
        const float4 hq = *reinterpret_cast<const float4*>(&hn[lane * 4]);
        for (int r = warp * 16; r < warp * 16 + 16; ++r) {
            float4 w = ldg4(w1 + (size_t)r * 128 + lane * 4);
            float s = w.x * hq.x + w.y * hq.y + w.z * hq.z + w.w * hq.w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) o1[r] = s + b1[r];
        }
    }
    __syncthreads();
    for (int r = warp; r < nout; r += 4) {
        float s = __ldg(w2 + (size_t)r * 64 + lane) * o1[lane] + __ldg(w2 + (size_t)r * 64 + 32 + lane) * o1[32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) pred[(size_t)t * nout + r] = s + b2[r];
    }
}

}  // namespace deft

using namespace deft;

extern "C" {

int deft_sigmoid_output(float* hm, size_t n_hm, float* dep, size_t n_dep, float depth_scale, deft_stream_t s) {
    size_t n = n_hm > n_dep ? n_hm : n_dep;
    if (n == 0) return DEFT_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_pdl(sigmoid_output_kernel, dim3(blocks), dim3(256), 0, to_stream(s), hm, hm ? n_hm : 0, dep, dep ? n_dep : 0, depth_scale);
    DEFT_CHECK_LAUNCH("sigmoid_output");
    return DEFT_OK;
}

static int decode_chunks(int H, int W) { return ceil_div(H * W, TOPK_CHUNK); }

size_t deft_decode_workspace_bytes(int B, int C, int H, int W, int K) {
    return align_up((size_t)B * C * decode_chunks(H, W) * K * sizeof(unsigned long long), 256) +
           align_up((size_t)B * C * K * sizeof(unsigned long long), 256);
}

int deft_decode(const float* hm, int B, int C, int H, int W, int K, const float* reg, const float* wh, const float* ltrb_amodal,
                float* scores, int64_t* inds, int32_t* clses, float* ys, float* xs, float* bboxes, float* bboxes_amodal,
                void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(hm && scores && inds && clses && ys && xs && ws, "decode: null argument");
    DEFT_REQUIRE(K >= 1 && K <= 1024 && K <= H * W, "decode: K must be in [1, min(1024, H*W)] (K=%d)", K);
    const int nchunks = decode_chunks(H, W);
    int Np = 1, Mp = 1;
    while (Np < C * K) Np <<= 1;
    while (Mp < nchunks * K) Mp <<= 1;
    DEFT_REQUIRE(Np <= 16384 && Mp <= 16384, "decode: C*K or chunks*K too large (%d, %d)", C * K, nchunks * K);
    if (ws_bytes < deft_decode_workspace_bytes(B, C, H, W, K)) {
        set_error("decode: workspace too small");
        return DEFT_EWORKSPACE;
    }
    unsigned long long* cand_chunk = reinterpret_cast<unsigned long long*>(ws);
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<char*>(ws) + align_up((size_t)B * C * nchunks * K * sizeof(unsigned long long), 256));
    launch_pdl(topk_chunk_kernel, dim3(nchunks, C, B), dim3(1024), 0, to_stream(s), hm, H, W, K, nchunks, cand_chunk);
    DEFT_CHECK_LAUNCH("topk_chunk");
    static size_t granted_b[DEFT_MAX_DEVICES] = {}, granted_c[DEFT_MAX_DEVICES] = {};
    size_t smem_b = (size_t)Mp * sizeof(unsigned long long);
    int rc_attr = ensure_dyn_smem(topk_class_merge_kernel, smem_b, granted_b, "topk_class_merge");
    if (rc_attr != DEFT_OK) return rc_attr;
    launch_pdl(topk_class_merge_kernel, dim3(C, B), dim3(1024), smem_b, to_stream(s), cand_chunk, nchunks * K, Mp, K, cand);
    DEFT_CHECK_LAUNCH("topk_class_merge");
    DecodeArgs a{cand, reg, wh, ltrb_amodal, scores, (long long*)inds, clses, ys, xs, bboxes, bboxes_amodal, C, H, W, K};
    size_t smem = (size_t)Np * sizeof(unsigned long long);
    if ((rc_attr = ensure_dyn_smem(decode_merge_kernel, smem, granted_c, "decode_merge")) != DEFT_OK) return rc_attr;
    launch_pdl(decode_merge_kernel, dim3(B), dim3(1024), smem, to_stream(s), a, Np);
    DEFT_CHECK_LAUNCH("decode_merge");
    return DEFT_OK;
}

int deft_gather_feat(const float* feat, int B, int C, int HW, const int64_t* inds, int K, float* out, deft_stream_t s) {
    DEFT_REQUIRE(feat && inds && out && B > 0 && C > 0 && K > 0, "gather_feat: bad args");
    int total = B * K * C;
    launch_pdl(gather_feat_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, to_stream(s), feat, C, HW, (const long long*)inds, K, out, total);
    DEFT_CHECK_LAUNCH("gather_feat");
    return DEFT_OK;
}

int deft_boxes_to_centres(const float* boxes, int N, float W, float H, float* centres, deft_stream_t s) {
    DEFT_REQUIRE(boxes && centres && N > 0 && W > 0 && H > 0, "boxes_to_centres: bad args");
    launch_pdl(boxes_to_centres_kernel, dim3(ceil_div(N, 128)), dim3(128), 0, to_stream(s), boxes, N, W, H, centres);
    DEFT_CHECK_LAUNCH("boxes_to_centres");
    return DEFT_OK;
}

struct AffinityWs { float *au, *bv, *h2, *h3, *h4, *S; size_t total; };
// pairs = rows of the pair grid; au_rows / bv_rows = rows of the projected first-layer halves kept in the workspace
// (0 when the caller supplies them: cached Au' of the memory ring, deft_afe_project output).
static AffinityWs affinity_ws_layout(void* base, size_t pairs, size_t au_rows, size_t bv_rows) {
    AffinityWs w{};
    size_t off = 0;
    char* b = reinterpret_cast<char*>(base);
    auto take = [&](size_t bytes) { float* p = reinterpret_cast<float*>(b + off); off += align_up(bytes, 256); return p; };
    w.au = take(au_rows * 512 * 4);
    w.bv = take(bv_rows * 512 * 4);
    w.h2 = take(pairs * 256 * 4);
    w.h3 = take(pairs * 128 * 4);
    w.h4 = take(pairs * 64 * 4);
    w.S = take(pairs * 4);
    w.total = off;
    return w;
}

size_t deft_afe_affinity_workspace_bytes(int F, int P, int Cemb) {
    (void)Cemb;
    return affinity_ws_layout(nullptr, (size_t)F * P * P, (size_t)F * P, (size_t)P).total;
}
size_t deft_afe_pairs_workspace_bytes(int F, int rows, int cols) {
    return affinity_ws_layout(nullptr, (size_t)F * rows * cols, 0, 0).total;
}

static int plain_gemm(const float* A, int lda, int Mrows, int K, const float* Wp, const float* Wtc, int tcbn, int ldw, const float* sc, const float* sh,
                      float* out, int ldo, int Cout, int relu, unsigned flags, cudaStream_t st) {
    ConvAP ap{};
    ap.nsrc = 1; ap.Ctot = K; ap.src[0] = A; ap.c[0] = K; ap.pitch[0] = lda;
    ap.N = 1; ap.H = 1; ap.W = Mrows; ap.OH = 1; ap.OW = Mrows; ap.KH = 1; ap.KW = 1; ap.stride = 1; ap.pad = 0;
    Epilogue ep{sc, sh, nullptr, 0, out, ldo, Cout, relu};
    if (flags & DEFT_FLAG_TENSORCORE) {
        int rc = tc_conv_dispatch(ap, Wtc, tcbn, ldw, K, Mrows, ep, flags, nullptr, 0, st);
        if (rc != DEFT_EINVAL) return rc;
        note_simt_fallback("deft_afe_affinity(gemm)", "K=%d ldw=%d M=%d bn=%d", K, ldw, Mrows, tcbn);
    }
    return launch_igemm(ap, Wp, ldw, K, Mrows, ep, st);
}

// Layers 2..5 of the pairwise MLP (AFE.py:331-347 after the decomposed first layer) over the M pair rows `pap` describes:
// relu(Au'_i + Bv'_j) -> 256 -> 128 -> 64 -> relu(dot) = S[M].
struct AffinityMlp {
    const float *w2, *sc2, *sh2, *w3, *sc3, *sh3, *w4, *b4, *w5;
    float b5;
    const float *t2, *t3, *t4;
    int bn2, bn3, bn4;
};
static int affinity_mlp(const PairAP& pap, int M, const AffinityMlp& w, const AffinityWs& ws, unsigned flags, cudaStream_t st) {
    Epilogue e2{w.sc2, w.sh2, nullptr, 0, ws.h2, 256, 256, 1};
    int rc = DEFT_EINVAL;
    if (flags & DEFT_FLAG_TENSORCORE) rc = tc_pair_dispatch(pap, w.t2, w.bn2, 256, 512, M, e2, flags, st);
    if (rc == DEFT_EINVAL) {
        if (flags & DEFT_FLAG_TENSORCORE) note_simt_fallback("deft_afe_affinity(pair GEMM)", "M=%d bn=%d", M, w.bn2);
        rc = launch_igemm(pap, w.w2, 256, 512, M, e2, st);
    }
    if (rc != DEFT_OK) return rc;
    if ((rc = plain_gemm(ws.h2, 256, M, 256, w.w3, w.t3, w.bn3, 128, w.sc3, w.sh3, ws.h3, 128, 128, 1, flags, st)) != DEFT_OK) return rc;
    if ((rc = plain_gemm(ws.h3, 128, M, 128, w.w4, w.t4, w.bn4, 64, nullptr, w.b4, ws.h4, 64, 64, 1, flags, st)) != DEFT_OK) return rc;
    launch_pdl(affinity_final_kernel, dim3(ceil_div(M, 32)), dim3(256), 0, st, (const float*)ws.h4, w.w5, w.b5, ws.S, M);
    DEFT_CHECK_LAUNCH("affinity_final");
    return DEFT_OK;
}
static AffinityMlp make_mlp(const float* w2, const float* sc2, const float* sh2, const float* w3, const float* sc3, const float* sh3, const float* w4,
                            const float* b4, const float* w5, float b5, const float* const* tc, const int* tcbn) {
    AffinityMlp w{w2, sc2, sh2, w3, sc3, sh3, w4, b4, w5, b5, nullptr, nullptr, nullptr, 0, 0, 0};
    if (tc && tcbn) { w.t2 = tc[2]; w.t3 = tc[3]; w.t4 = tc[4]; w.bn2 = tcbn[2]; w.bn3 = tcbn[3]; w.bn4 = tcbn[4]; }
    return w;
}

int deft_afe_project(const float* emb, int rows, int Cemb, const float* w1a, const float* w1b, const float* sh1, const float* const* tc,
                     const int* tcbn, float* au, float* bv, unsigned flags, deft_stream_t s) {
    DEFT_REQUIRE(emb && w1a && w1b && sh1 && (au || bv) && rows >= 1 && Cemb % 16 == 0, "afe_project: bad args");
    cudaStream_t st = to_stream(s);
    const unsigned fl = flags & ~DEFT_FLAG_TF32X1;
    int rc;
    if (au && (rc = plain_gemm(emb, Cemb, rows, Cemb, w1a, tc ? tc[0] : nullptr, tcbn ? tcbn[0] : 0, 512, nullptr, sh1, au, 512, 512, 0, fl, st)) != DEFT_OK) return rc;
    if (bv && (rc = plain_gemm(emb, Cemb, rows, Cemb, w1b, tc ? tc[1] : nullptr, tcbn ? tcbn[1] : 0, 512, nullptr, nullptr, bv, 512, 512, 0, fl, st)) != DEFT_OK) return rc;
    return DEFT_OK;
}

int deft_afe_affinity(const float* xp, const int* n_host, int F, const float* xn, int m, int P, int Cemb, const float* w1a,
                      const float* w1b, const float* sh1, const float* w2, const float* sc2, const float* sh2, const float* w3,
                      const float* sc3, const float* sh3, const float* w4, const float* b4, const float* w5, float b5,
                      const float* const* tc, const int* tcbn, float* out,
                      unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(xp && n_host && xn && out && ws && w1a && w1b && sh1 && w2 && w3 && w4 && w5, "afe_affinity: null argument");
    DEFT_REQUIRE(F >= 1 && F <= AFF_MAX_FRAMES && m >= 1 && m <= P && P <= 1024 && Cemb % 16 == 0,
                 "afe_affinity: need 1<=F<=%d, 1<=m<=P<=1024, Cemb%%16==0", AFF_MAX_FRAMES);
    if (ws_bytes < deft_afe_affinity_workspace_bytes(F, P, Cemb)) { set_error("afe_affinity: workspace too small"); return DEFT_EWORKSPACE; }
    const AffinityWs w = affinity_ws_layout(ws, (size_t)F * P * P, (size_t)F * P, (size_t)P);
    cudaStream_t st = to_stream(s);
    PairAP pap{};
    SoftmaxArgs sa{};
    int M = 0;
    for (int f = 0; f < F; ++f) {
        DEFT_REQUIRE(n_host[f] >= 0 && n_host[f] <= P, "afe_affinity: n[%d]=%d out of range", f, n_host[f]);
        pap.off[f] = M; sa.off[f] = M; sa.n[f] = n_host[f];
        M += n_host[f] * m;
    }
    pap.off[F] = M; sa.off[F] = M;
    sa.m = m; sa.ms = m; sa.P = P;
    int rc;
    if (M > 0) {
        // layer 1, decomposed: Au' = bnfold(W1a u_i + b1), Bv' = bnfold-scale(W1b v_j)
        if ((rc = deft_afe_project(xp, F * P, Cemb, w1a, w1b, sh1, tc, tcbn, w.au, nullptr, flags, s)) != DEFT_OK) return rc;
        if ((rc = deft_afe_project(xn, m, Cemb, w1a, w1b, sh1, tc, tcbn, nullptr, w.bv, flags, s)) != DEFT_OK) return rc;
        pap.au = w.au; pap.bv = w.bv; pap.P = P; pap.m = m; pap.F = F; pap.ld = 512;
        if ((rc = affinity_mlp(pap, M, make_mlp(w2, sc2, sh2, w3, sc3, sh3, w4, b4, w5, b5, tc, tcbn), w, flags, st)) != DEFT_OK) return rc;
    }
    // M == 0 (no live pair): the softmax kernel still clears the output blocks
    launch_pdl(affinity_softmax_kernel, dim3(F), dim3(256), 4 * P * sizeof(float), st, (const float*)w.S, sa, out);
    DEFT_CHECK_LAUNCH("affinity_softmax");
    return DEFT_OK;
}

int deft_afe_affinity_cached(const float* au_mem, int F, int P, int rows, const int* n_dev, const float* bv, int m_max, const int* m_dev,
                             const float* w2, const float* sc2, const float* sh2, const float* w3, const float* sc3, const float* sh3,
                             const float* w4, const float* b4, const float* w5, float b5, const float* const* tc, const int* tcbn,
                             float* out, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(au_mem && n_dev && bv && out && ws && w2 && w3 && w4 && w5, "afe_affinity_cached: null argument");
    DEFT_REQUIRE(F >= 1 && F <= AFF_MAX_FRAMES && rows >= 1 && rows <= P && m_max >= 1 && m_max <= P && P <= 1024,
                 "afe_affinity_cached: need 1<=F<=%d, 1<=rows<=P, 1<=m_max<=P<=1024", AFF_MAX_FRAMES);
    if (ws_bytes < deft_afe_pairs_workspace_bytes(F, rows, m_max)) { set_error("afe_affinity_cached: workspace too small"); return DEFT_EWORKSPACE; }
    const AffinityWs w = affinity_ws_layout(ws, (size_t)F * rows * m_max, 0, 0);
    // uniform pair grid: frame f owns rows [f*rows*m_max, (f+1)*rows*m_max); rows >= n_dev[f] / columns >= *m_dev are
    // computed (the live counts stay on the device, the launch is graph-capturable) and ignored by the softmax
    PairAP pap{};
    SoftmaxArgs sa{};
    for (int f = 0; f <= F; ++f) { pap.off[f] = f * rows * m_max; sa.off[f] = pap.off[f]; }
    pap.au = au_mem; pap.bv = bv; pap.P = P; pap.m = m_max; pap.F = F; pap.ld = 512;
    sa.n_dev = n_dev; sa.m_dev = m_dev; sa.m = m_max; sa.ms = m_max; sa.P = P;
    cudaStream_t st = to_stream(s);
    int rc = affinity_mlp(pap, F * rows * m_max, make_mlp(w2, sc2, sh2, w3, sc3, sh3, w4, b4, w5, b5, tc, tcbn), w, flags, st);
    if (rc != DEFT_OK) return rc;
    launch_pdl(affinity_softmax_kernel, dim3(F), dim3(256), 4 * P * sizeof(float), st, (const float*)w.S, sa, out);
    DEFT_CHECK_LAUNCH("affinity_softmax");
    return DEFT_OK;
}

int deft_afe_forward_scores(const float* au, const float* bv, int B, int P, const float* w2, const float* sc2, const float* sh2,
                            const float* w3, const float* sc3, const float* sh3, const float* w4, const float* b4, const float* w5, float b5,
                            const float* const* tc, const int* tcbn, float false_constant, float* out, unsigned flags, void* ws, size_t ws_bytes,
                            deft_stream_t s) {
    DEFT_REQUIRE(au && bv && out && ws && w2 && w3 && w4 && w5, "afe_forward_scores: null argument");
    DEFT_REQUIRE(B >= 1 && B <= AFF_MAX_FRAMES && P >= 1 && P <= 1024, "afe_forward_scores: need 1<=B<=%d, 1<=P<=1024", AFF_MAX_FRAMES);
    if (ws_bytes < deft_afe_pairs_workspace_bytes(B, P, P)) { set_error("afe_forward_scores: workspace too small"); return DEFT_EWORKSPACE; }
    const AffinityWs w = affinity_ws_layout(ws, (size_t)B * P * P, 0, 0);
    PairAP pap{};
    for (int f = 0; f <= B; ++f) pap.off[f] = f * P * P;
    pap.au = au; pap.bv = bv; pap.P = P; pap.m = P; pap.F = B; pap.ld = 512; pap.bvf = (long long)P * 512;
    cudaStream_t st = to_stream(s);
    int rc = affinity_mlp(pap, B * P * P, make_mlp(w2, sc2, sh2, w3, sc3, sh3, w4, b4, w5, b5, tc, tcbn), w, flags, st);
    if (rc != DEFT_OK) return rc;
    launch_pdl(affinity_scores_kernel, dim3(ceil_div((P + 1) * (P + 1), 256), B), dim3(256), 0, st, (const float*)w.S, P, false_constant, out);
    DEFT_CHECK_LAUNCH("affinity_scores");
    return DEFT_OK;
}

int deft_lstm_step(const float* x, const float* h, const float* c, const float* w_ih, const float* w_hh, const float* b_ih,
                   const float* b_hh, const float* w1, const float* b1, const float* w2, const float* b2, int T, int F, int nout,
                   float* h_out, float* c_out, float* pred, deft_stream_t s) {
    DEFT_REQUIRE(x && h && c && w_ih && w_hh && b_ih && b_hh && w1 && b1 && w2 && b2 && h_out && c_out && pred, "lstm_step: null argument");
    DEFT_REQUIRE(T >= 1 && F >= 1 && F <= 32 && nout >= 1 && nout <= 64, "lstm_step: need F<=32, nout<=64 (F=%d nout=%d)", F, nout);
    launch_pdl(lstm_step_kernel, dim3(T), dim3(128), 0, to_stream(s), x, h, c, w_ih, w_hh, b_ih, b_hh, w1, b1, w2, b2, F, nout, h_out, c_out, pred);
    DEFT_CHECK_LAUNCH("lstm_step");
    return DEFT_OK;
}

}  // extern "C"
