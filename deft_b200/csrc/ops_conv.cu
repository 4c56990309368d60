// Convolution-family entry points of the C ABI (SIMT fp32 engine): conv+BN+act, 7x7 stem, maxpool,
// DCNv2, depthwise-deconv upsample+add, CenterNet heads, layout converters.
#include "igemm_simt.cuh"
#include "tc_engine.cuh"
#include "conv_c16.cuh"

namespace deft {

// ------------------------------------------------------------------------------------------------
// layout converters (tiled smem transpose)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int dst_pitch, int C, int HW) {
    __shared__ float tile[32][33];
    int n = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) tile[i][threadIdx.x] = src[((size_t)n * C + c) * HW + p];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (c < C && p < HW) dst[((size_t)n * HW + p) * dst_pitch + c] = tile[threadIdx.x][i];
    }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int src_pitch, float* __restrict__ dst, int C, int HW) {
    __shared__ float tile[32][33];
    int n = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (c < C && p < HW) tile[i][threadIdx.x] = src[((size_t)n * HW + p) * src_pitch + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------
// 7x7 stem straight from the NCHW image: a block computes 64x16 output pixels, each thread 4 consecutive
// pixels x 16 channels (the weight LDS.128s are amortised over 4 pixels); input halo + weights in smem.
constexpr int STEM_TW = 64, STEM_TH = 16, STEM_LD = 72;   // tile row stride (floats), 16 B aligned
__global__ void __launch_bounds__(256) stem7x7_kernel(const float* __restrict__ x, int H, int W,
                                                      const float* __restrict__ wp, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, float* __restrict__ out, int out_pitch) {
    DEFT_PDL_PROLOGUE();
    __shared__ __align__(16) float tile[3][STEM_TH + 6][STEM_LD];
    __shared__ __align__(16) float ws[147 * 16];
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * STEM_TW, y0 = blockIdx.y * STEM_TH;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < 147 * 16; i += 256) ws[i] = wp[i];
    for (int i = tid; i < 3 * (STEM_TH + 6) * (STEM_TW + 6); i += 256) {
        int c = i / ((STEM_TH + 6) * (STEM_TW + 6)), r = i - c * ((STEM_TH + 6) * (STEM_TW + 6));
        int yy = r / (STEM_TW + 6), xx = r - yy * (STEM_TW + 6);
        int gy = y0 + yy - 3, gx = x0 + xx - 3;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = x[((size_t)(n * 3 + c) * H + gy) * W + gx];
        tile[c][yy][xx] = v;
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

// ------------------------------------------------------------------------------------------------
__global__ void maxpool2x2_kernel(const float* __restrict__ in, int in_pitch, float* __restrict__ out, int out_pitch,
                                  int N, int H, int W, int C4) {
    DEFT_PDL_PROLOGUE();
    const int OH = H / 2, OW = W / 2;
    size_t total = (size_t)N * OH * OW * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4);
        size_t p = i / C4;
        int ox = (int)(p % OW);
        size_t t = p / OW;
        int oy = (int)(t % OH);
        int n = (int)(t / OH);
        const float* b = in + (((size_t)n * H + 2 * oy) * W + 2 * ox) * in_pitch + c4 * 4;
        float4 a = ldg4(b), bq = ldg4(b + in_pitch), c = ldg4(b + (size_t)W * in_pitch), d = ldg4(b + (size_t)(W + 1) * in_pitch);
        float4 m;
        m.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(c.x, d.x)); m.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(c.y, d.y));
        m.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(c.z, d.z)); m.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(c.w, d.w));
        *reinterpret_cast<float4*>(out + p * out_pitch + c4 * 4) = m;
    }
}

// ------------------------------------------------------------------------------------------------
// Depthwise ConvTranspose2d(k=2f, s=f, p=f/2) + skip add.  Each output pixel receives exactly the
// 2x2 input pixels {iy1-1, iy1} x {ix1-1, ix1}, iy1 = (Y+p)/f, with taps ky = (Y+p) - iy*f.
__global__ void upsample_add_kernel(const float* __restrict__ x, int x_pitch, const float* __restrict__ wp,
                                    const float* __restrict__ skip, int skip_pitch, float* __restrict__ out, int out_pitch,
                                    int N, int h, int w, int C, int f) {
    DEFT_PDL_PROLOGUE();
    const int OH = h * f, OW = w * f, C4 = C / 4, k = 2 * f, pd = f / 2;
    size_t total = (size_t)N * OH * OW * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        size_t p = i / C4;
        int X = (int)(p % OW);
        size_t t = p / OW;
        int Y = (int)(t % OH);
        int n = (int)(t / OH);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (skip) acc = ldg4(skip + p * skip_pitch + c);
        int iy1 = (Y + pd) / f, ix1 = (X + pd) / f;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            int iy = iy1 - a;
            int ky = Y + pd - iy * f;
            if (iy < 0 || iy >= h || ky >= k) continue;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                int ix = ix1 - b;
                int kx = X + pd - ix * f;
                if (ix < 0 || ix >= w || kx >= k) continue;
                float4 v = ldg4(x + (((size_t)n * h + iy) * w + ix) * x_pitch + c);
                float4 ww = ldg4(wp + (size_t)(ky * k + kx) * C + c);
                acc.x += v.x * ww.x; acc.y += v.y * ww.y; acc.z += v.z * ww.z; acc.w += v.w * ww.w;
            }
        }
        *reinterpret_cast<float4*>(out + p * out_pitch + c) = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Head 1x1 convs: out_h[n, c, pix] = b2[row] + sum_j hidden[pix, h*HC + j] * w2[row, j].
// One block = 32 consecutive pixels; a warp reduces one pixel at a time; results are staged in smem
// so the planar NCHW stores are coalesced.
struct HeadsOut {
    float* out[DEFT_MAX_HEADS];
    int c[DEFT_MAX_HEADS];
    int row0[DEFT_MAX_HEADS];
    int nheads, total_c;
};
__global__ void __launch_bounds__(256) heads1x1_kernel(const float* __restrict__ hidden, int hid_pitch, int HC,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       HeadsOut ho, int HW, int total_pix) {
    extern __shared__ __align__(16) float sm[];
    float* w2s = sm;                               // [total_c][HC]
    float* res = sm + (size_t)ho.total_c * HC;     // [total_c][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < ho.total_c * HC; i += 256) w2s[i] = w2[i];
    __syncthreads();
    const int pix0 = blockIdx.x * 32;
    for (int pp = warp; pp < 32; pp += 8) {
        int pix = pix0 + pp;
        if (pix >= total_pix) break;
        const float* hrow = hidden + (size_t)pix * hid_pitch;
        for (int h = 0; h < ho.nheads; ++h) {
            // HC = 256: every lane owns 8 consecutive hidden channels
            float hv[8];
            {
                const int j0 = lane * 8;
                float4 a = ldg4(hrow + h * HC + j0), b = ldg4(hrow + h * HC + j0 + 4);
                hv[0] = a.x; hv[1] = a.y; hv[2] = a.z; hv[3] = a.w; hv[4] = b.x; hv[5] = b.y; hv[6] = b.z; hv[7] = b.w;
            }
            for (int c = 0; c < ho.c[h]; ++c) {
                const float* wr = w2s + (size_t)(ho.row0[h] + c) * HC + lane * 8;
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += hv[j] * wr[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) res[(ho.row0[h] + c) * 32 + pp] = s + b2[ho.row0[h] + c];
            }
        }
    }
    __syncthreads();
    // coalesced planar stores: 32 pixels of one output channel at a time (a block never straddles
    // an image when HW % 32 == 0; otherwise handle per pixel)
    for (int i = tid; i < ho.total_c * 32; i += 256) {
        int row = i >> 5, pp = i & 31;
        int pix = pix0 + pp;
        if (pix >= total_pix) continue;
        int h = 0;
        while (h < ho.nheads - 1 && row >= ho.row0[h + 1]) ++h;
        int c = row - ho.row0[h];
        int n = pix / HW, q = pix - n * HW;
        ho.out[h][((size_t)n * ho.c[h] + c) * HW + q] = res[i];
    }
}

}  // namespace deft

using namespace deft;

extern "C" {

int deft_tc_tile_n(int ldw, int M, int kind) { return tc_pick_bn(ldw, M, kind); }

size_t deft_tc_blob_bytes(int ldw, int K, int bn) {
    if (ldw <= 0 || bn <= 0 || ldw % bn != 0 || K <= 0) return 0;
    return align_up(tc_blob_floats(ldw, K, bn) * sizeof(float), 256);
}

int deft_tc_pack_weights(const float* w_packed, int ldw, int K, int bn, int taps, float* blob, deft_stream_t s) {
    DEFT_REQUIRE(w_packed && blob && ldw > 0 && bn > 0 && ldw % bn == 0 && K > 0 && taps >= 1, "tc_pack_weights: bad args");
    return tc_pack_weights(w_packed, ldw, K, bn, taps, blob, to_stream(s));
}

size_t deft_tc_scratch_bytes(int ldw, int K, int M, int kind) { return tc_splitk_scratch_bytes(ldw, K, M, kind); }

int deft_nchw_to_nhwc(const float* src, float* dst, int dst_pitch, int N, int C, int H, int W, deft_stream_t s) {
    DEFT_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && dst_pitch >= C, "nchw_to_nhwc: bad args");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, to_stream(s)>>>(src, dst, dst_pitch, C, H * W);
    DEFT_CHECK_LAUNCH("nchw_to_nhwc");
    return DEFT_OK;
}

int deft_nhwc_to_nchw(const float* src, int src_pitch, float* dst, int N, int C, int H, int W, deft_stream_t s) {
    DEFT_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && src_pitch >= C, "nhwc_to_nchw: bad args");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, to_stream(s)>>>(src, src_pitch, dst, C, H * W);
    DEFT_CHECK_LAUNCH("nhwc_to_nchw");
    return DEFT_OK;
}

int deft_conv2d_bn_act(const float* const* srcs, const int* src_c, const int* src_pitch, int nsrc, int N, int H, int W,
                       const float* w_packed, int ldw, const float* w_tc, int tc_bn, const float* scale, const float* shift,
                       const float* residual, int res_pitch, float* out, int out_pitch, int Cout, int KH, int KW, int stride, int pad,
                       unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(srcs && src_c && src_pitch && nsrc >= 1 && nsrc <= DEFT_MAX_SRC, "conv2d: nsrc must be 1..%d", DEFT_MAX_SRC);
    DEFT_REQUIRE(w_packed && out && N > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && stride > 0, "conv2d: bad args");
    ConvAP ap{};
    ap.nsrc = nsrc;
    ap.Ctot = 0;
    for (int i = 0; i < nsrc; ++i) {
        DEFT_REQUIRE(srcs[i] && src_c[i] % 16 == 0 && src_pitch[i] >= src_c[i] && src_pitch[i] % 4 == 0,
                     "conv2d: source %d needs C%%16==0 and pitch%%4==0 (C=%d pitch=%d)", i, src_c[i], src_pitch[i]);
        ap.src[i] = srcs[i];
        ap.c[i] = src_c[i];
        ap.pitch[i] = src_pitch[i];
        ap.Ctot += src_c[i];
    }
    ap.N = N; ap.H = H; ap.W = W; ap.KH = KH; ap.KW = KW; ap.stride = stride; ap.pad = pad;
    ap.OH = (H + 2 * pad - KH) / stride + 1;
    ap.OW = (W + 2 * pad - KW) / stride + 1;
    DEFT_REQUIRE(out_pitch % 4 == 0 && (!residual || res_pitch % 4 == 0), "conv2d: pitches must be multiples of 4");
    Epilogue ep{scale, shift, residual, res_pitch, out, out_pitch, Cout, (flags & DEFT_FLAG_RELU) ? 1 : 0};
    int M = N * ap.OH * ap.OW;
    int K = KH * KW * ap.Ctot;
    // the two 16-channel full-resolution layers (DLA level0 / level1): tensor-core path when a weight image packed for it is
    // supplied (K padded to 192, taps = 1: ops.conv_tc_pack), else the direct fp32 kernel
    if (nsrc == 1 && src_c[0] == 16 && KH == 3 && KW == 3 && pad == 1 && (flags & DEFT_FLAG_TENSORCORE) && w_tc != nullptr && !(flags & DEFT_FLAG_TF32X1)) {
        int rc = tc_conv_dispatch(ap, w_tc, tc_bn, ldw, K, M, ep, flags, ws, ws_bytes, to_stream(s));
        if (rc != DEFT_EINVAL) return rc;
    }
    if (nsrc == 1 && src_c[0] == 16 && KH == 3 && KW == 3 && pad == 1 && residual == nullptr && !(flags & DEFT_FLAG_NO_DIRECT)) {
        const int relu = (flags & DEFT_FLAG_RELU) ? 1 : 0;
        if (Cout == 16 && ldw == 16 && stride == 1)
            return launch_conv3x3_c16<16, 1, 4>(srcs[0], src_pitch[0], N, H, W, ap.OH, ap.OW, w_packed, ldw, scale, shift, out, out_pitch, relu, to_stream(s));
        if (Cout == 32 && ldw == 32 && stride == 2)
            return launch_conv3x3_c16<32, 2, 2>(srcs[0], src_pitch[0], N, H, W, ap.OH, ap.OW, w_packed, ldw, scale, shift, out, out_pitch, relu, to_stream(s));
    }
    if (flags & DEFT_FLAG_TENSORCORE) {
        int rc = tc_conv_dispatch(ap, w_tc, tc_bn, ldw, K, M, ep, flags, ws, ws_bytes, to_stream(s));
        if (rc != DEFT_EINVAL) return rc;   // EINVAL from the TC engine = shape not covered -> SIMT
        note_simt_fallback("deft_conv2d_bn_act", "Ctot=%d Cout=%d k=%dx%d s=%d M=%d bn=%d blob=%p", ap.Ctot, Cout, KH, KW, stride, M, tc_bn, (const void*)w_tc);
    }
    return launch_igemm(ap, w_packed, ldw, K, M, ep, to_stream(s));
}

int deft_stem_conv7x7(const float* x, int N, int H, int W, const float* w_packed, const float* scale, const float* shift,
                      float* out, int out_pitch, deft_stream_t s) {
    DEFT_REQUIRE(x && w_packed && scale && shift && out && out_pitch >= 16 && out_pitch % 4 == 0, "stem: bad args");
    dim3 grid(ceil_div(W, STEM_TW), ceil_div(H, STEM_TH), N), block(256);
    launch_pdl(stem7x7_kernel, dim3(grid), dim3(block), 0, to_stream(s), x, H, W, w_packed, scale, shift, out, out_pitch);
    DEFT_CHECK_LAUNCH("stem7x7");
    return DEFT_OK;
}

int deft_maxpool2x2(const float* in, int in_pitch, float* out, int out_pitch, int N, int H, int W, int C, deft_stream_t s) {
    DEFT_REQUIRE(in && out && H % 2 == 0 && W % 2 == 0 && C % 4 == 0 && in_pitch % 4 == 0 && out_pitch % 4 == 0,
                 "maxpool: needs even H,W and C%%4==0");
    size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(maxpool2x2_kernel, dim3(blocks), dim3(256), 0, to_stream(s), in, in_pitch, out, out_pitch, N, H, W, C / 4);
    DEFT_CHECK_LAUNCH("maxpool2x2");
    return DEFT_OK;
}

static size_t dcn_om_bytes(int N, int H, int W) { return align_up((size_t)N * H * W * 32 * sizeof(float), 256); }

size_t deft_dcnv2_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
    const int M = N * H * W, ldw = (Cout + 15) / 16 * 16;
    size_t s1 = tc_splitk_scratch_bytes(32, 9 * Cin, M, TC_KIND_CONV), s2 = tc_splitk_scratch_bytes(ldw, 9 * Cin, M, TC_KIND_DCN);
    return dcn_om_bytes(N, H, W) + (s1 > s2 ? s1 : s2);
}

int deft_dcnv2_forward(const float* x, int x_pitch, int N, int H, int W, int Cin, const float* w_om_packed,
                       const float* b_om, const float* w_om_tc, int om_bn, const float* w_packed, int ldw, const float* w_tc,
                       int tc_bn, const float* scale, const float* shift,
                       float* y, int y_pitch, int Cout, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(x && w_om_packed && b_om && w_packed && y && ws, "dcnv2: null argument");
    DEFT_REQUIRE(Cin % 16 == 0 && x_pitch % 4 == 0 && y_pitch % 4 == 0, "dcnv2: Cin%%16 / pitch%%4 (Cin=%d)", Cin);
    if (ws_bytes < deft_dcnv2_workspace_bytes(N, H, W, Cin, Cout)) {
        set_error("dcnv2: workspace too small");
        return DEFT_EWORKSPACE;
    }
    float* om = reinterpret_cast<float*>(ws);
    const int M = N * H * W;
    void* scratch = reinterpret_cast<char*>(ws) + dcn_om_bytes(N, H, W);
    const size_t scratch_bytes = ws_bytes - dcn_om_bytes(N, H, W);
    // 1) offset/mask conv: plain 3x3 conv Cin -> 27 (padded to 32), + bias, no activation
    ConvAP cap{};
    cap.nsrc = 1; cap.Ctot = Cin; cap.src[0] = x; cap.c[0] = Cin; cap.pitch[0] = x_pitch;
    cap.N = N; cap.H = H; cap.W = W; cap.OH = H; cap.OW = W; cap.KH = 3; cap.KW = 3; cap.stride = 1; cap.pad = 1;
    Epilogue e1{nullptr, b_om, nullptr, 0, om, 32, 32, 0};
    int rc = DEFT_EINVAL;
    if (flags & DEFT_FLAG_TENSORCORE) rc = tc_conv_dispatch(cap, w_om_tc, om_bn, 32, 9 * Cin, M, e1, flags & ~DEFT_FLAG_TF32X1, scratch, scratch_bytes, to_stream(s));
    if (rc == DEFT_EINVAL) {
        if (flags & DEFT_FLAG_TENSORCORE) note_simt_fallback("deft_dcnv2_forward(offset conv)", "Cin=%d M=%d bn=%d", Cin, M, om_bn);
        rc = launch_igemm(cap, w_om_packed, 32, 9 * Cin, M, e1, to_stream(s));
    }
    if (rc != DEFT_OK) return rc;
    // 2) deformable gather fused into the GEMM's A producer
    DcnAP dap{};
    dap.x = x; dap.pitch = x_pitch; dap.N = N; dap.H = H; dap.W = W; dap.Cin = Cin; dap.om = om;
    Epilogue e2{scale, shift, nullptr, 0, y, y_pitch, Cout, (flags & DEFT_FLAG_RELU) ? 1 : 0};
    if (flags & DEFT_FLAG_TENSORCORE) {
        rc = tc_dcn_dispatch(dap, w_tc, tc_bn, ldw, 9 * Cin, M, e2, flags, scratch, scratch_bytes, to_stream(s));
        if (rc != DEFT_EINVAL) return rc;
        note_simt_fallback("deft_dcnv2_forward(gather GEMM)", "Cin=%d Cout=%d M=%d bn=%d", Cin, Cout, M, tc_bn);
    }
    return launch_igemm(dap, w_packed, ldw, 9 * Cin, M, e2, to_stream(s));
}

int deft_upsample_add(const float* x, int x_pitch, const float* w_packed, const float* skip, int skip_pitch, float* out,
                      int out_pitch, int N, int h, int w, int C, int f, deft_stream_t s) {
    DEFT_REQUIRE(x && w_packed && out && C % 4 == 0 && (f == 2 || f == 4 || f == 8) && x_pitch % 4 == 0 && out_pitch % 4 == 0 &&
                     (!skip || skip_pitch % 4 == 0), "upsample_add: bad args (C=%d f=%d)", C, f);
    size_t total = (size_t)N * h * f * w * f * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(upsample_add_kernel, dim3(blocks), dim3(256), 0, to_stream(s), x, x_pitch, w_packed, skip, skip_pitch, out, out_pitch, N, h, w, C, f);
    DEFT_CHECK_LAUNCH("upsample_add");
    return DEFT_OK;
}

size_t deft_heads_workspace_bytes(int N, int H, int W, int nheads, int head_conv) {
    return align_up((size_t)N * H * W * nheads * head_conv * sizeof(float), 256);
}

int deft_heads_forward(const float* feat, int feat_pitch, int N, int H, int W, int Cin, const float* w1_packed, const float* w1_tc,
                       const float* b1, int head_conv, const float* w2, const float* b2, const int* head_c, int nheads,
                       float* const* outs, unsigned flags, void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(feat && w1_packed && b1 && w2 && b2 && head_c && outs && ws, "heads: null argument");
    DEFT_REQUIRE(nheads >= 1 && nheads <= DEFT_MAX_HEADS && head_conv == 256 && Cin % 16 == 0, "heads: nheads<=%d, head_conv==256", DEFT_MAX_HEADS);
    float* hidden = reinterpret_cast<float*>(ws);
    const int HCt = nheads * head_conv;
    const int M = N * H * W;
    HeadsOut ho{};
    ho.nheads = nheads;
    int row = 0;
    for (int h = 0; h < nheads; ++h) {
        DEFT_REQUIRE(outs[h] && head_c[h] > 0, "heads: bad head %d", h);
        ho.out[h] = outs[h]; ho.c[h] = head_c[h]; ho.row0[h] = row;
        row += head_c[h];
    }
    ho.total_c = row;
    ConvAP cap{};
    cap.nsrc = 1; cap.Ctot = Cin; cap.src[0] = feat; cap.c[0] = Cin; cap.pitch[0] = feat_pitch;
    cap.N = N; cap.H = H; cap.W = W; cap.OH = H; cap.OW = W; cap.KH = 3; cap.KW = 3; cap.stride = 1; cap.pad = 1;
    if ((flags & DEFT_FLAG_TENSORCORE) && w1_tc != nullptr) {
        // fused path: 3x3 conv + bias + ReLU + the head's 1x1 conv in one tcgen05 kernel, no hidden map in HBM
        HeadsArgs ha{};
        ha.b1 = b1; ha.w2 = w2; ha.b2 = b2; ha.nheads = nheads; ha.HW = H * W;
        for (int h = 0; h < nheads; ++h) { ha.out[h] = ho.out[h]; ha.c[h] = ho.c[h]; ha.row0[h] = ho.row0[h]; }
        int rc = tc_heads_dispatch(cap, w1_tc, 9 * Cin, M, ha, flags, to_stream(s));
        if (rc != DEFT_EINVAL) return rc;
        note_simt_fallback("deft_heads_forward", "Cin=%d nheads=%d M=%d", Cin, nheads, M);
    }
    if (ws_bytes < deft_heads_workspace_bytes(N, H, W, nheads, head_conv)) {   // only the unfused path needs the hidden map
        set_error("heads: workspace too small");
        return DEFT_EWORKSPACE;
    }
    Epilogue e1{nullptr, b1, nullptr, 0, hidden, HCt, HCt, 1};
    int rc = launch_igemm(cap, w1_packed, HCt, 9 * Cin, M, e1, to_stream(s));
    if (rc != DEFT_OK) return rc;
    size_t smem = ((size_t)row * head_conv + (size_t)row * 32) * sizeof(float);
    DEFT_REQUIRE(smem <= 200 * 1024, "heads: too many output channels (%d)", row);
    static size_t granted[DEFT_MAX_DEVICES] = {};
    if ((rc = ensure_dyn_smem(heads1x1_kernel, smem, granted, "heads1x1")) != DEFT_OK) return rc;
    heads1x1_kernel<<<ceil_div(M, 32), 256, smem, to_stream(s)>>>(hidden, HCt, head_conv, w2, b2, ho, H * W, M);
    DEFT_CHECK_LAUNCH("heads1x1");
    return DEFT_OK;
}

}  // extern "C"
