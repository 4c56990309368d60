// a7: AFE embedding extractor (AFE_module.forward_feature_extracter, AFE.py:88-92,162-188), sparse form.
//
// The reference runs 13 dense selector convs (9.7 GMAC at 1088x608) and then samples N points; here
// relu(conv3x3) is evaluated only at the (up to) 4 bilinear corners of every centre and the corners are
// blended with grid_sample's weights (bilinear, padding_mode=border).  Two kernels:
//   afe_corner_kernel : grid (ceil(N/8), 13 maps, K-splits).  A block owns 8 detections = 32 corner rows of one
//                       map and one channel slice; the 4x4 input patches of its detections sit in smem, the
//                       selector weights are streamed once per block (coalesced over the output channel) and
//                       reused by all 32 rows -> raw partial sums [map][split][4N rows][64].
//   afe_blend_kernel  : emb[n, off_k + o] = sum_corner w_corner * relu(bias + sum_split partial).
#include "common.cuh"

namespace deft {

constexpr int AFE_DETS = 8;          // detections per block
constexpr int AFE_CCHUNK = 64;       // channels staged in smem at a time
constexpr int AFE_MAXSPLIT = 4;

struct AfeArgs {
    const float* map[DEFT_AFE_MAPS];
    const float* w[DEFT_AFE_MAPS];
    const float* b[DEFT_AFE_MAPS];
    int pitch[DEFT_AFE_MAPS], C[DEFT_AFE_MAPS], H[DEFT_AFE_MAPS], W[DEFT_AFE_MAPS], Co[DEFT_AFE_MAPS], off[DEFT_AFE_MAPS], ns[DEFT_AFE_MAPS];
    int Cemb, align_corners, N;
};

// grid_sample source index, padding_mode=border (clip), for one axis
__device__ __forceinline__ float afe_src_index(float g, int size, int align_corners) {
    float i = align_corners ? ((g + 1.f) / 2.f) * (float)(size - 1) : ((g + 1.f) * (float)size - 1.f) / 2.f;
    return fminf((float)(size - 1), fmaxf(i, 0.f));
}

__global__ void __launch_bounds__(256) afe_corner_kernel(AfeArgs a, const float* __restrict__ centres, float* __restrict__ partial) {
    DEFT_PDL_PROLOGUE();
    __shared__ __align__(16) float patch[AFE_DETS * 16 * AFE_CCHUNK];   // [det][4x4 pixel][channel]
    __shared__ int sx0[AFE_DETS], sy0[AFE_DETS];
    const int k = blockIdx.y, z = blockIdx.z, tid = threadIdx.x;
    const int nsplit = a.ns[k];
    if (z >= nsplit) return;
    const int C = a.C[k], H = a.H[k], W = a.W[k], Co = a.Co[k], pitch = a.pitch[k];
    const int cps = C / nsplit;                    // channels handled by this split
    const int c_begin = z * cps;
    const int det0 = blockIdx.x * AFE_DETS;
    if (tid < AFE_DETS) {
        int n = det0 + tid;
        int x0 = 0, y0 = 0;
        if (n < a.N) {
            x0 = (int)floorf(afe_src_index(centres[n * 2 + 0], W, a.align_corners));
            y0 = (int)floorf(afe_src_index(centres[n * 2 + 1], H, a.align_corners));
        }
        sx0[tid] = x0;
        sy0[tid] = y0;
    }
    __syncthreads();
    const int o = tid & 63, rg = tid >> 6;         // output channel, row group (2 detections x 4 corners)
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    const float* m = a.map[k];
    const float* wk = a.w[k];
    for (int cc0 = c_begin; cc0 < c_begin + cps; cc0 += AFE_CCHUNK) {
        const int cw = min(AFE_CCHUNK, c_begin + cps - cc0);
        const int cw4 = cw >> 2;
        // stage the 8 detections' 4x4 patches (zero outside the map = the conv's zero padding)
        for (int i = tid; i < AFE_DETS * 16 * cw4; i += 256) {
            int c4 = i % cw4;
            int t = i / cw4;
            int pp = t & 15, d = t >> 4;
            int yy = sy0[d] - 1 + (pp >> 2), xx = sx0[d] - 1 + (pp & 3);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (det0 + d < a.N && yy >= 0 && yy < H && xx >= 0 && xx < W) v = ldg4(m + ((size_t)yy * W + xx) * pitch + cc0 + c4 * 4);
            *reinterpret_cast<float4*>(&patch[(d * 16 + pp) * AFE_CCHUNK + c4 * 4]) = v;
        }
        __syncthreads();
        if (o < Co) {
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
                const float* wr = wk + ((size_t)tap * C + cc0) * Co + o;
                // channel counts are multiples of 16: 4 channel quads per step, their 16 weight loads issued together
                // (the loop is latency-bound on these L2 loads otherwise)
                for (int c4 = 0; c4 < cw4; c4 += 4) {
                    float w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = __ldg(wr + (size_t)(c4 * 4 + j) * Co);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const int d = rg * 2 + (r >> 2), cy = (r >> 1) & 1, cx = r & 1;
                            const float4 pv = *reinterpret_cast<const float4*>(&patch[(d * 16 + (cy + ky) * 4 + (cx + kx)) * AFE_CCHUNK + (c4 + u) * 4]);
                            acc[r] += pv.x * w[u * 4] + pv.y * w[u * 4 + 1] + pv.z * w[u * 4 + 2] + pv.w * w[u * 4 + 3];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (o < Co) {
        const size_t rows = (size_t)a.N * 4;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int n = det0 + rg * 2 + (r >> 2);
            if (n < a.N) partial[(((size_t)k * AFE_MAXSPLIT + z) * rows + (size_t)n * 4 + (r & 3)) * 64 + o] = acc[r];
        }
    }
}

__global__ void afe_blend_kernel(AfeArgs a, const float* __restrict__ centres, const float* __restrict__ partial, float* __restrict__ emb) {
    DEFT_PDL_PROLOGUE();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.N * a.Cemb) return;
    const int n = idx / a.Cemb, e = idx - n * a.Cemb;
    int k = 0;
    while (k < DEFT_AFE_MAPS - 1 && e >= a.off[k + 1]) ++k;
    const int o = e - a.off[k];
    const int H = a.H[k], W = a.W[k];
    const float ix = afe_src_index(centres[n * 2 + 0], W, a.align_corners), iy = afe_src_index(centres[n * 2 + 1], H, a.align_corners);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float tx = ix - fx, ty = iy - fy;
    // grid_sample corner weights: nw, ne, sw, se (corner index = cy*2 + cx)
    float cwt[4] = {(fx + 1.f - ix) * (fy + 1.f - iy), tx * (fy + 1.f - iy), (fx + 1.f - ix) * ty, tx * ty};
    const size_t rows = (size_t)a.N * 4;
    const float bias = a.b[k][o];
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool inb = (y0 + (c >> 1) <= H - 1) && (x0 + (c & 1) <= W - 1);    // torch skips out-of-bounds corners
        if (!inb) continue;
        float s = bias;
        for (int z = 0; z < a.ns[k]; ++z) s += partial[(((size_t)k * AFE_MAXSPLIT + z) * rows + (size_t)n * 4 + c) * 64 + o];
        v += cwt[c] * fmaxf(s, 0.f);
    }
    emb[idx] = v;
}

}  // namespace deft

using namespace deft;

extern "C" {

size_t deft_afe_extract_workspace_bytes(int N) {
    return align_up((size_t)DEFT_AFE_MAPS * AFE_MAXSPLIT * (size_t)N * 4 * 64 * sizeof(float), 256);
}

int deft_afe_extract(const float* const* maps, const int* pitch, const int* C, const int* H, const int* W,
                     const float* const* selw, const float* const* selb, const int* Co, const float* centres, int N,
                     float* emb, int align_corners, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(maps && pitch && C && H && W && selw && selb && Co && centres && emb && ws && N > 0, "afe_extract: bad args");
    if (ws_bytes < deft_afe_extract_workspace_bytes(N)) {
        set_error("afe_extract: workspace too small");
        return DEFT_EWORKSPACE;
    }
    AfeArgs a{};
    int off = 0;
    for (int k = 0; k < DEFT_AFE_MAPS; ++k) {
        DEFT_REQUIRE(maps[k] && selw[k] && selb[k] && C[k] % 16 == 0 && Co[k] <= 64 && Co[k] >= 1 && pitch[k] % 4 == 0,
                     "afe_extract: map %d needs C%%16==0, Co<=64", k);
        a.map[k] = maps[k]; a.w[k] = selw[k]; a.b[k] = selb[k];
        a.pitch[k] = pitch[k]; a.C[k] = C[k]; a.H[k] = H[k]; a.W[k] = W[k]; a.Co[k] = Co[k]; a.off[k] = off;
        a.ns[k] = C[k] >= 512 ? 4 : (C[k] >= 256 ? 2 : 1);       // keep every block's channel slice <= 128
        off += Co[k];
    }
    a.Cemb = off;
    a.align_corners = align_corners;
    a.N = N;
    float* partial = reinterpret_cast<float*>(ws);
    launch_pdl(afe_corner_kernel, dim3(ceil_div(N, AFE_DETS), DEFT_AFE_MAPS, AFE_MAXSPLIT), dim3(256), 0, to_stream(s), a, centres, partial);
    DEFT_CHECK_LAUNCH("afe_corner");
    launch_pdl(afe_blend_kernel, dim3(ceil_div(N * a.Cemb, 256)), dim3(256), 0, to_stream(s), a, centres, partial, emb);
    DEFT_CHECK_LAUNCH("afe_blend");
    return DEFT_OK;
}

}  // extern "C"
