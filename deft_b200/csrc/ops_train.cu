// Training side (BASELINE config C5): the backward of the DCNv2 module -- the one native
// operator the reference's training path depends on (upstream dcn_v2_backward: grad_input, grad_offset, grad_mask,
// grad_weight, grad_bias; call site dla.py:652-663 under autograd, trainer.py:222-227).
//
//   y = DCN(x) = b + sum_{tap, ci} W[co, ci, tap] * sigmoid(m[tap]) * bilinear(x[ci]; p + tap + off[tap]),  (off, m) = conv3x3(x; w_om) + b_om
//
// fp32 SIMT throughout (exact arithmetic for gradient parity; the tensor-core forward engine stays the inference
// path).  Structure = the published DCNv2 backward, NHWC:
//   1. om      = offset/mask conv forward (recomputed, not saved)                       [P, 32]
//   2. gcol    = grad_y . W^T                                                           [P, 9 Cin]   (SIMT implicit GEMM)
//   3. scatter : grad_x += w_corner * mask * gcol (atomicAdd), grad_off / grad_mask per (pixel, tap) -> grad_om
//   4. col     = mask * bilinear gather (recomputed into the same buffer)               [P, 9 Cin]
//   5. grad_W  = col^T . grad_y, grad_b = column sums of grad_y                          (split-K A^T B GEMM, fixed-order reduce)
//   6. offset conv backward: grad_x += conv3x3^T(grad_om), grad_w_om = im2col(x)^T . grad_om, grad_b_om = column sums
// grad_x's scatter uses float atomics like upstream's col2im: the summation order (not the value set) is run-dependent.
#include "igemm_simt.cuh"

namespace deft {

// ------------------------------------------------------------------------------------------------ small helpers
// dst[c][r] = src[r][c]   (src [R x ld_src] first C columns; dst [C x ld_dst])
__global__ void transpose_kernel(const float* __restrict__ src, int R, int C, int ld_src, float* __restrict__ dst, int ld_dst) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) dst[(size_t)c * ld_dst + r] = tile[threadIdx.x][i];
    }
}

// W'[(tap', j)][ci] = w_om[(8 - tap', ci)][j]: the offset conv's weights flipped for its transposed (data-gradient) conv
__global__ void flip_om_weights_kernel(const float* __restrict__ w_om, int Cin, float* __restrict__ wf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * 32 * Cin) return;
    const int ci = i % Cin, r = i / Cin, j = r % 32, tap = r / 32;
    wf[i] = w_om[((size_t)(8 - tap) * Cin + ci) * 32 + j];
}

// out[c] = sum_p a[p][c]  (deterministic: one block per 32 columns, fixed strided order)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, int P, int C, int ld, float* __restrict__ out) {
    __shared__ double red[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    double s = 0.0;
    if (c < C)
        for (int p = w; p < P; p += 8) s += (double)a[(size_t)p * ld + c];
    red[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < C) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x & 31];
        out[c] = (float)t;
    }
}

// plain im2col of a 3x3 / pad 1 conv over NHWC x: col[p][(tap, ci)]
__global__ void im2col3x3_kernel(const float* __restrict__ x, int pitch, int N, int H, int W, int Cin, float* __restrict__ col) {
    const int c4 = Cin / 4;
    const size_t total = (size_t)N * H * W * 9 * c4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        size_t r = i / c4;
        const int tap = (int)(r % 9);
        const size_t p = r / 9;
        const int xx = (int)(p % W), y = (int)((p / W) % H);
        const int n = (int)(p / ((size_t)W * H));
        const int iy = y - 1 + tap / 3, ix = xx - 1 + tap % 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = ldg4(x + (((size_t)n * H + iy) * W + ix) * pitch + c);
        *reinterpret_cast<float4*>(col + (p * 9 + tap) * Cin + c) = v;
    }
}

// ------------------------------------------------------------------------------------------------ sampling geometry
struct Corner {
    int o[4];          // pixel indices of (y0,x0) (y0,x1) (y1,x0) (y1,x1); -1 = outside the map
    float ly, lx;      // fractional position
    float mk;          // sigmoid(mask logit)
    bool inside;       // sample point inside (-1, H) x (-1, W)
};
__device__ __forceinline__ Corner corner_of(const float* __restrict__ om, size_t p, int tap, int H, int W) {
    Corner c;
    const int n = (int)(p / ((size_t)H * W));
    const int r = (int)(p - (size_t)n * H * W);
    const int y = r / W, xx = r - y * W;
    const float* q = om + p * 32;
    const float dy = q[2 * tap], dx = q[2 * tap + 1];
    c.mk = sigmoidf_(q[18 + tap]);
    const float py = (float)(y - 1 + tap / 3) + dy, px = (float)(xx - 1 + tap % 3) + dx;
    c.inside = py > -1.f && py < (float)H && px > -1.f && px < (float)W;
    const float fy = floorf(py), fx = floorf(px);
    const int y0 = (int)fy, x0 = (int)fx;
    c.ly = py - fy; c.lx = px - fx;
    const int base = n * H * W;
    const bool y0ok = y0 >= 0 && y0 <= H - 1, y1ok = y0 + 1 >= 0 && y0 + 1 <= H - 1, x0ok = x0 >= 0 && x0 <= W - 1, x1ok = x0 + 1 >= 0 && x0 + 1 <= W - 1;
    c.o[0] = (c.inside && y0ok && x0ok) ? base + y0 * W + x0 : -1;
    c.o[1] = (c.inside && y0ok && x1ok) ? base + y0 * W + x0 + 1 : -1;
    c.o[2] = (c.inside && y1ok && x0ok) ? base + (y0 + 1) * W + x0 : -1;
    c.o[3] = (c.inside && y1ok && x1ok) ? base + (y0 + 1) * W + x0 + 1 : -1;
    return c;
}

// One warp per (pixel, tap); lanes stride the channels.
//   mode 0 (col):     col[p][(tap, ci)] = mk * bilinear(x[ci])
//   mode 1 (scatter): g = gcol[p][(tap, ci)];  grad_x[corner, ci] += w_corner * mk * g;
//                     grad_mask_logit = mk (1 - mk) * sum_ci g * bilinear;  grad_off_{y,x} = mk * sum_ci g * d bilinear / d{py,px}
__global__ void __launch_bounds__(256) dcn_gather_bwd_kernel(const float* __restrict__ x, int pitch, int N, int H, int W, int Cin,
                                                             const float* __restrict__ om, float* __restrict__ colbuf, int mode,
                                                             float* __restrict__ grad_x, int gx_pitch, float* __restrict__ grad_om) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t total = (size_t)N * H * W * 9;
    if (warp >= total) return;
    const size_t p = warp / 9;
    const int tap = (int)(warp - p * 9);
    const Corner c = corner_of(om, p, tap, H, W);
    const float hy = 1.f - c.ly, hx = 1.f - c.lx;
    const float w[4] = {hy * hx, hy * c.lx, c.ly * hx, c.ly * c.lx};
    float* colrow = colbuf + (p * 9 + tap) * Cin;
    float s_val = 0.f, s_dy = 0.f, s_dx = 0.f;
    for (int ci = lane; ci < Cin; ci += 32) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = c.o[k] >= 0 ? __ldg(x + (size_t)c.o[k] * pitch + ci) : 0.f;
        const float val = w[0] * v[0] + w[1] * v[1] + w[2] * v[2] + w[3] * v[3];
        if (mode == 0) {
            colrow[ci] = c.mk * val;
        } else {
            const float g = colrow[ci];
            const float gm = g * c.mk;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c.o[k] >= 0 && w[k] != 0.f) atomicAdd(grad_x + (size_t)c.o[k] * gx_pitch + ci, w[k] * gm);
            s_val += g * val;
            // d val / d py, d val / d px (corners outside the map count as zeros, like upstream's get_coordinate_weight)
            s_dy += gm * (hx * (v[2] - v[0]) + c.lx * (v[3] - v[1]));
            s_dx += gm * (hy * (v[1] - v[0]) + c.ly * (v[3] - v[2]));
        }
    }
    if (mode == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
            s_dy += __shfl_xor_sync(0xffffffffu, s_dy, o);
            s_dx += __shfl_xor_sync(0xffffffffu, s_dx, o);
        }
        if (lane == 0) {
            float* g = grad_om + p * 32;
            g[2 * tap] = c.inside ? s_dy : 0.f;
            g[2 * tap + 1] = c.inside ? s_dx : 0.f;
            g[18 + tap] = c.inside ? s_val * c.mk * (1.f - c.mk) : 0.f;
            if (tap == 0) { g[27] = 0.f; g[28] = 0.f; g[29] = 0.f; g[30] = 0.f; g[31] = 0.f; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ C = A^T B (reduction over rows)
// A [K x lda] (first M columns), B [K x ldb] (first Nn columns): part[z][m][n] = sum_{k in slice z} A[k][m] B[k][n].
// 64 x 64 tile per block, 256 threads x (4 x 4); K slices of 16 through smem.
__global__ void __launch_bounds__(256) atb_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int K, int M, int Nn,
                                                       int k_per_split, float* __restrict__ part) {
    __shared__ __align__(16) float As[16][64];
    __shared__ __align__(16) float Bs[16][64];
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64, z = blockIdx.z;
    const int kbeg = z * k_per_split, kend = min(K, kbeg + k_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lk = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        const int k = k0 + lk;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
            if (m0 + lc + 3 < M) a = ldg4(A + (size_t)k * lda + m0 + lc);
            else { float t[4] = {0.f, 0.f, 0.f, 0.f}; for (int q = 0; q < 4; ++q) if (m0 + lc + q < M) t[q] = A[(size_t)k * lda + m0 + lc + q]; a = make_float4(t[0], t[1], t[2], t[3]); }
            if (n0 + lc + 3 < Nn) b = ldg4(B + (size_t)k * ldb + n0 + lc);
            else { float t[4] = {0.f, 0.f, 0.f, 0.f}; for (int q = 0; q < 4; ++q) if (n0 + lc + q < Nn) t[q] = B[(size_t)k * ldb + n0 + lc + q]; b = make_float4(t[0], t[1], t[2], t[3]); }
        }
        *reinterpret_cast<float4*>(&As[lk][lc]) = a;
        *reinterpret_cast<float4*>(&Bs[lk][lc]) = b;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a4[i] * b4[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < Nn) part[((size_t)z * M + m) * Nn + n] = acc[i][j];
        }
    }
}
__global__ void atb_reduce_kernel(const float* __restrict__ part, int splits, int M, int Nn, float* __restrict__ out, int ldo) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)M * Nn) return;
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += (double)part[(size_t)z * M * Nn + i];     // fixed order
    out[(i / Nn) * ldo + (i % Nn)] = (float)s;
}

static int atb_splits(int K) { int s = K / 2048; return s < 1 ? 1 : (s > 64 ? 64 : s); }

static int atb_gemm(const float* A, int lda, const float* B, int ldb, int K, int M, int Nn, float* part, float* out, int ldo, cudaStream_t st) {
    const int splits = atb_splits(K);
    const int kps = ceil_div(ceil_div(K, splits), 16) * 16;
    const int sp = ceil_div(K, kps);
    atb_gemm_kernel<<<dim3(ceil_div(M, 64), ceil_div(Nn, 64), sp), 256, 0, st>>>(A, lda, B, ldb, K, M, Nn, kps, part);
    DEFT_CHECK_LAUNCH("atb_gemm");
    atb_reduce_kernel<<<ceil_div(M * Nn, 256), 256, 0, st>>>(part, sp, M, Nn, out, ldo);
    DEFT_CHECK_LAUNCH("atb_reduce");
    return DEFT_OK;
}

struct BwdWs { float *om, *gom, *col, *gx0, *wT, *wf, *part; size_t total; };
static BwdWs bwd_layout(void* base, int N, int H, int W, int Cin, int Cout) {
    BwdWs w{};
    const size_t P = (size_t)N * H * W;
    const int ldw = (Cout + 15) / 16 * 16;
    size_t off = 0;
    char* b = reinterpret_cast<char*>(base);
    auto take = [&](size_t bytes) { float* p = reinterpret_cast<float*>(b + off); off += align_up(bytes, 256); return p; };
    w.om = take(P * 32 * 4);
    w.gom = take(P * 32 * 4);
    w.col = take(P * 9 * Cin * 4);
    w.gx0 = take(P * Cin * 4);
    w.wT = take((size_t)ldw * 9 * Cin * 4);
    w.wf = take((size_t)9 * 32 * Cin * 4);
    const size_t nmax = ldw > 32 ? ldw : 32;
    w.part = take((size_t)atb_splits((int)P) * 9 * Cin * nmax * 4);
    w.total = off;
    return w;
}

}  // namespace deft

using namespace deft;

extern "C" {

size_t deft_dcnv2_backward_workspace_bytes(int N, int H, int W, int Cin, int Cout) { return bwd_layout(nullptr, N, H, W, Cin, Cout).total; }

int deft_dcnv2_backward(const float* x, int x_pitch, int N, int H, int W, int Cin, const float* w_om_packed, const float* b_om,
                        const float* w_packed, int ldw, int Cout, const float* grad_y, int gy_pitch, float* grad_x, int gx_pitch,
                        float* grad_w_om /* [9Cin][32] */, float* grad_b_om /* [32] */, float* grad_w /* [9Cin][ldw] */, float* grad_b /* [Cout] */,
                        void* ws, size_t ws_bytes, deft_stream_t s) {
    DEFT_REQUIRE(x && w_om_packed && b_om && w_packed && grad_y && grad_x && grad_w_om && grad_b_om && grad_w && grad_b && ws, "dcnv2_backward: null argument");
    DEFT_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0 && ldw == Cout && x_pitch % 4 == 0 && gy_pitch % 4 == 0 && gx_pitch % 4 == 0,
                 "dcnv2_backward: needs Cin%%16==0, Cout%%16==0 (ldw == Cout), pitches %%4 (Cin=%d Cout=%d ldw=%d)", Cin, Cout, ldw);
    if (ws_bytes < deft_dcnv2_backward_workspace_bytes(N, H, W, Cin, Cout)) { set_error("dcnv2_backward: workspace too small"); return DEFT_EWORKSPACE; }
    const BwdWs w = bwd_layout(ws, N, H, W, Cin, Cout);
    cudaStream_t st = to_stream(s);
    const int M = N * H * W, K9 = 9 * Cin;
    int rc;
    // 1) offset/mask conv forward (recomputed)
    ConvAP cap{};
    cap.nsrc = 1; cap.Ctot = Cin; cap.src[0] = x; cap.c[0] = Cin; cap.pitch[0] = x_pitch;
    cap.N = N; cap.H = H; cap.W = W; cap.OH = H; cap.OW = W; cap.KH = 3; cap.KW = 3; cap.stride = 1; cap.pad = 1;
    Epilogue e1{nullptr, b_om, nullptr, 0, w.om, 32, 32, 0};
    if ((rc = launch_igemm(cap, w_om_packed, 32, K9, M, e1, st)) != DEFT_OK) return rc;
    // 2) gcol = grad_y . W^T   (W^T [Cout][9Cin] from the packed [9Cin][ldw])
    transpose_kernel<<<dim3(ceil_div(Cout, 32), ceil_div(K9, 32)), dim3(32, 8), 0, st>>>(w_packed, K9, Cout, ldw, w.wT, K9);
    DEFT_CHECK_LAUNCH("transpose_w");
    ConvAP gap{};
    gap.nsrc = 1; gap.Ctot = Cout; gap.src[0] = grad_y; gap.c[0] = Cout; gap.pitch[0] = gy_pitch;
    gap.N = 1; gap.H = 1; gap.W = M; gap.OH = 1; gap.OW = M; gap.KH = 1; gap.KW = 1; gap.stride = 1; gap.pad = 0;
    Epilogue e2{nullptr, nullptr, nullptr, 0, w.col, K9, K9, 0};
    if ((rc = launch_igemm(gap, w.wT, K9, Cout, M, e2, st)) != DEFT_OK) return rc;
    // 3) scatter: grad_x (gather part), grad_offset / grad_mask -> grad_om
    if (cudaMemsetAsync(w.gx0, 0, (size_t)M * Cin * sizeof(float), st) != cudaSuccess) { set_error("dcnv2_backward: memset failed"); return DEFT_ECUDA; }
    const unsigned warps = (unsigned)((size_t)M * 9);
    dcn_gather_bwd_kernel<<<ceil_div((int)warps, 8), 256, 0, st>>>(x, x_pitch, N, H, W, Cin, w.om, w.col, 1, w.gx0, Cin, w.gom);
    DEFT_CHECK_LAUNCH("dcn_scatter");
    // 4) col = mask * bilinear gather   (overwrites gcol)
    dcn_gather_bwd_kernel<<<ceil_div((int)warps, 8), 256, 0, st>>>(x, x_pitch, N, H, W, Cin, w.om, w.col, 0, nullptr, 0, nullptr);
    DEFT_CHECK_LAUNCH("dcn_col");
    // 5) grad_W = col^T . grad_y ; grad_b = column sums of grad_y
    if ((rc = atb_gemm(w.col, K9, grad_y, gy_pitch, M, K9, Cout, w.part, grad_w, ldw, st)) != DEFT_OK) return rc;
    colsum_kernel<<<ceil_div(Cout, 32), 256, 0, st>>>(grad_y, M, Cout, gy_pitch, grad_b);
    DEFT_CHECK_LAUNCH("colsum_grad_y");
    // 6) offset conv backward
    flip_om_weights_kernel<<<ceil_div(9 * 32 * Cin, 256), 256, 0, st>>>(w_om_packed, Cin, w.wf);
    DEFT_CHECK_LAUNCH("flip_om_weights");
    ConvAP tap{};
    tap.nsrc = 1; tap.Ctot = 32; tap.src[0] = w.gom; tap.c[0] = 32; tap.pitch[0] = 32;
    tap.N = N; tap.H = H; tap.W = W; tap.OH = H; tap.OW = W; tap.KH = 3; tap.KW = 3; tap.stride = 1; tap.pad = 1;
    Epilogue e3{nullptr, nullptr, w.gx0, Cin, grad_x, gx_pitch, Cin, 0};           // grad_x = conv^T(grad_om) + scatter part
    if ((rc = launch_igemm(tap, w.wf, Cin, 9 * 32, M, e3, st)) != DEFT_OK) return rc;
    {
        const size_t total = (size_t)M * 9 * (Cin / 4);
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 32) blocks = 148 * 32;
        im2col3x3_kernel<<<blocks, 256, 0, st>>>(x, x_pitch, N, H, W, Cin, w.col);
        DEFT_CHECK_LAUNCH("im2col3x3");
    }
    if ((rc = atb_gemm(w.col, K9, w.gom, 32, M, K9, 32, w.part, grad_w_om, 32, st)) != DEFT_OK) return rc;
    colsum_kernel<<<1, 256, 0, st>>>(w.gom, M, 32, 32, grad_b_om);
    DEFT_CHECK_LAUNCH("colsum_grad_om");
    return DEFT_OK;
}

}  // extern "C"
