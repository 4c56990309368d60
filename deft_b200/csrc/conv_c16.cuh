// Direct fp32 3x3 convolution for the two full-resolution 16-input-channel layers of DLA-34
// (base.level0: 16->16 stride 1 @ HxW, base.level1: 16->32 stride 2, dla.py:306-309).  These layers are
// HBM-bound (42 MB in, 42 MB out at 1088x608) with only 144 MACs per output value: an implicit GEMM tile
// (K = 144, N = 16) wastes its instruction budget on im2col bookkeeping.  Here a block stages its input halo
// tile and the whole weight matrix in shared memory once and every thread computes PX output pixels x COUT
// channels from registers (PX x 4 x COUT FMAs per PX + COUT LDS.128).
//   smem input layout: [channel quad][row][col] of float4, so the 16 lanes of a row read consecutive 16-byte
//   words (conflict-free); a thread's pixels are strided by 16 columns for the same reason.
#pragma once
#include "common.cuh"

namespace deft {

template <int COUT, int STRIDE, int PX>
struct C16Cfg {
    static constexpr int TW = 16 * PX;                     // output tile width
    static constexpr int TH = 8;                           // output tile height (8 thread rows)
    static constexpr int IW = (TW - 1) * STRIDE + 3;       // input tile width incl. halo
    static constexpr int IH = (TH - 1) * STRIDE + 3;
    static constexpr int SMEM_BYTES = (4 * IH * IW * 4 + 144 * COUT) * 4;
};

template <int COUT, int STRIDE, int PX>
__global__ void __launch_bounds__(128) conv3x3_c16_kernel(const float* __restrict__ x, int x_pitch, int H, int W, int OH, int OW,
                                                          const float* __restrict__ wp, int ldw, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float* __restrict__ out, int out_pitch, int relu) {
    DEFT_PDL_PROLOGUE();
    using Cfg = C16Cfg<COUT, STRIDE, PX>;
    extern __shared__ __align__(16) float c16_smem[];
    float4* tile = reinterpret_cast<float4*>(c16_smem);                 // [4][IH][IW]
    float* ws = c16_smem + 4 * Cfg::IH * Cfg::IW * 4;                   // [144][COUT]
    const int n = blockIdx.z;
    const int ox0 = blockIdx.x * Cfg::TW, oy0 = blockIdx.y * Cfg::TH;
    const int ix0 = ox0 * STRIDE - 1, iy0 = oy0 * STRIDE - 1;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < 144 * COUT / 4; i += 128) {
        int k = i / (COUT / 4), c4 = i - k * (COUT / 4);
        reinterpret_cast<float4*>(ws)[i] = ldg4(wp + (size_t)k * ldw + c4 * 4);
    }
    for (int i = tid; i < Cfg::IH * Cfg::IW * 4; i += 128) {
        int q = i & 3, pix = i >> 2;
        int r = pix / Cfg::IW, c = pix - r * Cfg::IW;
        int gy = iy0 + r, gx = ix0 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = ldg4(x + (((size_t)n * H + gy) * W + gx) * x_pitch + q * 4);
        tile[(q * Cfg::IH + r) * Cfg::IW + c] = v;
    }
    __syncthreads();
    float acc[PX][COUT];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[p][o] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 in[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) in[p] = tile[(q * Cfg::IH + ty * STRIDE + ky) * Cfg::IW + (tx + 16 * p) * STRIDE + kx];
            const float* wr = ws + (tap * 16 + q * 4) * COUT;
#pragma unroll
            for (int o4 = 0; o4 < COUT / 4; ++o4) {
                const float4 w0 = *reinterpret_cast<const float4*>(wr + o4 * 4);
                const float4 w1 = *reinterpret_cast<const float4*>(wr + COUT + o4 * 4);
                const float4 w2 = *reinterpret_cast<const float4*>(wr + 2 * COUT + o4 * 4);
                const float4 w3 = *reinterpret_cast<const float4*>(wr + 3 * COUT + o4 * 4);
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    acc[p][o4 * 4 + 0] += in[p].x * w0.x + in[p].y * w1.x + in[p].z * w2.x + in[p].w * w3.x;
                    acc[p][o4 * 4 + 1] += in[p].x * w0.y + in[p].y * w1.y + in[p].z * w2.y + in[p].w * w3.y;
                    acc[p][o4 * 4 + 2] += in[p].x * w0.z + in[p].y * w1.z + in[p].z * w2.z + in[p].w * w3.z;
                    acc[p][o4 * 4 + 3] += in[p].x * w0.w + in[p].y * w1.w + in[p].z * w2.w + in[p].w * w3.w;
                }
            }
        }
    }
    const int oy = oy0 + ty;
    if (oy >= OH) return;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const int ox = ox0 + tx + 16 * p;
        if (ox >= OW) continue;
        float* o = out + (((size_t)n * OH + oy) * OW + ox) * out_pitch;
#pragma unroll
        for (int o4 = 0; o4 < COUT / 4; ++o4) {
            const float4 sc = scale ? ldg4(scale + o4 * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 sh = shift ? ldg4(shift + o4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 v;
            v.x = acc[p][o4 * 4 + 0] * sc.x + sh.x; v.y = acc[p][o4 * 4 + 1] * sc.y + sh.y;
            v.z = acc[p][o4 * 4 + 2] * sc.z + sh.z; v.w = acc[p][o4 * 4 + 3] * sc.w + sh.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            reinterpret_cast<float4*>(o)[o4] = v;
        }
    }
}

template <int COUT, int STRIDE, int PX>
inline int launch_conv3x3_c16(const float* x, int x_pitch, int N, int H, int W, int OH, int OW, const float* wp, int ldw, const float* scale,
                              const float* shift, float* out, int out_pitch, int relu, cudaStream_t st) {
    using Cfg = C16Cfg<COUT, STRIDE, PX>;
    static size_t granted[DEFT_MAX_DEVICES] = {};
    {
        const int rc_attr = ensure_dyn_smem(conv3x3_c16_kernel<COUT, STRIDE, PX>, Cfg::SMEM_BYTES, granted, "conv3x3_c16");
        if (rc_attr != DEFT_OK) return rc_attr;
    }
    dim3 grid(ceil_div(OW, Cfg::TW), ceil_div(OH, Cfg::TH), N);
    launch_pdl(conv3x3_c16_kernel<COUT, STRIDE, PX>, dim3(grid), dim3(128), Cfg::SMEM_BYTES, st, x, x_pitch, H, W, OH, OW, wp, ldw, scale, shift, out, out_pitch, relu);
    DEFT_CHECK_LAUNCH("conv3x3_c16");
    return DEFT_OK;
}

}  // namespace deft
