// fp32 SIMT implicit-GEMM engine: out[M, Cout] = epilogue( A[M, K] * Wp[K, ldw] ).
// The A operand is never materialised: an "A producer" functor builds 128x16 tiles on the fly
// (im2col over a channel-concat of NHWC sources, DCNv2 bilinear gather, or the affinity MLP's
// relu(Au_i + Bv_j) pair expansion).  This is the exact-fp32 anchor every tensor-core kernel is
// validated against.
#pragma once
#include "common.cuh"

namespace deft {

constexpr int IG_BM = 128;
constexpr int IG_BK = 16;
constexpr int IG_ALD = IG_BK + 4;  // padded A row (20 floats): conflict-free LDS.128 for adjacent rows
constexpr int IG_THREADS = 256;

struct Epilogue {
    const float* scale;     // [Cout] or nullptr (=1)
    const float* shift;     // [Cout] or nullptr (=0)
    const float* residual;  // [M, res_pitch] or nullptr
    int res_pitch;
    float* out;             // [M, out_pitch]
    int out_pitch;
    int Cout;               // multiple of 4
    int relu;
};

// ------------------------------------------------------------------------------------------------
// A producers.  Interface:
//   __device__ void init(int m0, int M, float* extra_smem);     // called by all threads, may sync
//   __device__ float4 load(int i, int kb) const;                // i in {0,1}: the thread's i-th float4
// Thread t owns tile rows r_i = (t + i*256) >> 2 and k-quad kq = t & 3 (k = kb*16 + kq*4 .. +3).
// ------------------------------------------------------------------------------------------------
struct ConvAP {
    const float* src[DEFT_MAX_SRC];
    int c[DEFT_MAX_SRC];
    int pitch[DEFT_MAX_SRC];
    int nsrc, Ctot;
    int N, H, W, OH, OW, KH, KW, stride, pad;
    static constexpr int kExtraSmemBytes = 0;
    // per-thread state
    int n_[2], oy_[2], ox_[2];
    bool valid_[2];

    __device__ __forceinline__ void init(int m0, int M, float*) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int p = m0 + ((threadIdx.x + i * IG_THREADS) >> 2);
            valid_[i] = p < M;
            int q = valid_[i] ? p : 0;
            n_[i] = q / (OH * OW);
            int r = q - n_[i] * (OH * OW);
            oy_[i] = r / OW;
            ox_[i] = r - oy_[i] * OW;
        }
    }
    __device__ __forceinline__ float4 load(int i, int kb) const {
        int k0 = kb * IG_BK;
        int tap = k0 / Ctot;
        int cc = k0 - tap * Ctot + (threadIdx.x & 3) * 4;
        int ky = tap / KW, kx = tap - ky * KW;
        int iy = oy_[i] * stride - pad + ky;
        int ix = ox_[i] * stride - pad + kx;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!valid_[i] || iy < 0 || iy >= H || ix < 0 || ix >= W) return z;
        int s = 0;
        while (s < nsrc - 1 && cc >= c[s]) { cc -= c[s]; ++s; }
        size_t pix = ((size_t)n_[i] * H + iy) * W + ix;
        return ldg4(src[s] + pix * pitch[s] + cc);
    }
};

// DCNv2: A[p, tap*Cin + ci] = mask_tap(p) * bilinear(x[:, ci]; p + tap + offset_tap(p)).
// Corner offsets/weights of the 128 pixels x 9 taps of the tile are computed once into smem.
struct DcnAP {
    const float* x;
    int pitch;
    int N, H, W, Cin;
    const float* om;  // [P, 32]: 0..17 offsets (2k = dy, 2k+1 = dx), 18..26 mask logits
    static constexpr int kExtraSmemBytes = IG_BM * 9 * 32;
    int4* tabO_;
    float4* tabW_;

    __device__ __forceinline__ void init(int m0, int M, float* extra) {
        tabO_ = reinterpret_cast<int4*>(extra);
        tabW_ = reinterpret_cast<float4*>(extra + IG_BM * 9 * 4);
        for (int idx = threadIdx.x; idx < IG_BM * 9; idx += IG_THREADS) {
            int row = idx / 9, tap = idx - row * 9;
            int p = m0 + row;
            int4 o = make_int4(0, 0, 0, 0);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < M) {
                int n = p / (H * W);
                int r = p - n * (H * W);
                int y = r / W, xx = r - y * W;
                const float* q = om + (size_t)p * 32;
                float dy = __ldg(q + 2 * tap), dx = __ldg(q + 2 * tap + 1);
                float mk = sigmoidf_(__ldg(q + 18 + tap));
                int ky = tap / 3, kx = tap - ky * 3;
                float py = (float)(y - 1 + ky) + dy;
                float px = (float)(xx - 1 + kx) + dx;
                if (py > -1.f && py < (float)H && px > -1.f && px < (float)W) {
                    float fy = floorf(py), fx = floorf(px);
                    int y0 = (int)fy, x0 = (int)fx;
                    float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
                    int base = n * H * W;
                    bool y0ok = y0 >= 0, y1ok = y0 + 1 <= H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= W - 1;
                    if (y0ok && x0ok) { o.x = base + y0 * W + x0; w.x = hy * hx * mk; }
                    if (y0ok && x1ok) { o.y = base + y0 * W + x0 + 1; w.y = hy * lx * mk; }
                    if (y1ok && x0ok) { o.z = base + (y0 + 1) * W + x0; w.z = ly * hx * mk; }
                    if (y1ok && x1ok) { o.w = base + (y0 + 1) * W + x0 + 1; w.w = ly * lx * mk; }
                }
            }
            tabO_[idx] = o;
            tabW_[idx] = w;
        }
        __syncthreads();
    }
    __device__ __forceinline__ float4 load(int i, int kb) const {
        int k0 = kb * IG_BK;
        int tap = k0 / Cin;
        int cc = k0 - tap * Cin + (threadIdx.x & 3) * 4;
        int row = (threadIdx.x + i * IG_THREADS) >> 2;
        int4 o = tabO_[row * 9 + tap];
        float4 w = tabW_[row * 9 + tap];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w.x != 0.f) { float4 v = ldg4(x + (size_t)o.x * pitch + cc); a.x += w.x * v.x; a.y += w.x * v.y; a.z += w.x * v.z; a.w += w.x * v.w; }
        if (w.y != 0.f) { float4 v = ldg4(x + (size_t)o.y * pitch + cc); a.x += w.y * v.x; a.y += w.y * v.y; a.z += w.y * v.z; a.w += w.y * v.w; }
        if (w.z != 0.f) { float4 v = ldg4(x + (size_t)o.z * pitch + cc); a.x += w.z * v.x; a.y += w.z * v.y; a.z += w.z * v.z; a.w += w.z * v.w; }
        if (w.w != 0.f) { float4 v = ldg4(x + (size_t)o.w * pitch + cc); a.x += w.w * v.x; a.y += w.w * v.y; a.z += w.w * v.z; a.w += w.w * v.w; }
        return a;
    }
};

// Affinity MLP layer-2 input: A[p, k] = relu(Au[f(p), i(p), k] + Bv[j(p), k]) over live pairs only.
constexpr int AFF_MAX_FRAMES = 64;
struct PairAP {
    const float* au;  // [F*P, 512]
    const float* bv;  // [P, 512] (+ f * bvf: one current frame per memory frame in the batched training forward, a9)
    int P, m, F, ld;
    long long bvf;    // element stride of bv between frames; 0 = every frame meets the same current detections
    int off[AFF_MAX_FRAMES + 1];  // prefix offsets of n_f*m
    static constexpr int kExtraSmemBytes = 0;
    const float* pa_[2];
    const float* pb_[2];

    __device__ __forceinline__ void init(int m0, int M, float*) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int p = m0 + ((threadIdx.x + i * IG_THREADS) >> 2);
            pa_[i] = nullptr;
            pb_[i] = nullptr;
            if (p < M) {
                int f = 0;
                while (f < F - 1 && p >= off[f + 1]) ++f;
                int q = p - off[f];
                int ii = q / m, jj = q - ii * m;
                pa_[i] = au + ((size_t)f * P + ii) * ld;
                pb_[i] = bv + (size_t)f * bvf + (size_t)jj * ld;
            }
        }
    }
    __device__ __forceinline__ float4 load(int i, int kb) const {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pa_[i] == nullptr) return z;
        int k = kb * IG_BK + (threadIdx.x & 3) * 4;
        float4 a = ldg4(pa_[i] + k), b = ldg4(pb_[i] + k);
        z.x = fmaxf(a.x + b.x, 0.f); z.y = fmaxf(a.y + b.y, 0.f);
        z.z = fmaxf(a.z + b.z, 0.f); z.w = fmaxf(a.w + b.w, 0.f);
        return z;
    }
};

// ------------------------------------------------------------------------------------------------
template <class AP, int BN>
__global__ void __launch_bounds__(IG_THREADS) igemm_simt_kernel(AP ap, const float* __restrict__ Wp, int ldw, int K,
                                                                int M, Epilogue ep) {
    constexpr int TX = BN / 4;            // column groups (4 couts each)
    constexpr int TY = IG_THREADS / TX;   // row groups
    constexpr int TM = IG_BM / TY;        // rows per thread (strided by TY)
    constexpr int B_F4 = IG_BK * BN / 4;  // float4s in a B tile
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                             // [2][BM][ALD]
    float* Bs = smem + 2 * IG_BM * IG_ALD;        // [2][BK][BN]
    float* extra = Bs + 2 * IG_BK * BN;

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.x * IG_BM;
    const int n0 = blockIdx.y * BN;
    ap.init(m0, M, extra);

    float acc[TM][4];
#pragma unroll
    for (int r = 0; r < TM; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }

    const int nkb = K / IG_BK;
    const int a_row0 = tid >> 2, a_kq = (tid & 3) * 4;
    const int b_k = tid / (BN / 4), b_c4 = (tid % (BN / 4)) * 4;
    const bool b_act = tid < B_F4;

    float4 ar[2], br = make_float4(0.f, 0.f, 0.f, 0.f);
    ar[0] = ap.load(0, 0);
    ar[1] = ap.load(1, 0);
    if (b_act) br = ldg4(Wp + (size_t)b_k * ldw + n0 + b_c4);
    *reinterpret_cast<float4*>(&As[(a_row0)*IG_ALD + a_kq]) = ar[0];
    *reinterpret_cast<float4*>(&As[(a_row0 + 64) * IG_ALD + a_kq]) = ar[1];
    if (b_act) *reinterpret_cast<float4*>(&Bs[b_k * BN + b_c4]) = br;
    __syncthreads();

    for (int kb = 0; kb < nkb; ++kb) {
        const int cur = kb & 1;
        const bool more = kb + 1 < nkb;
        if (more) {
            ar[0] = ap.load(0, kb + 1);
            ar[1] = ap.load(1, kb + 1);
            if (b_act) br = ldg4(Wp + (size_t)((kb + 1) * IG_BK + b_k) * ldw + n0 + b_c4);
        }
        const float* A = As + cur * IG_BM * IG_ALD;
        const float* B = Bs + cur * IG_BK * BN;
#pragma unroll
        for (int kk = 0; kk < IG_BK; kk += 4) {
            float4 b0 = *reinterpret_cast<const float4*>(&B[(kk + 0) * BN + tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&B[(kk + 1) * BN + tx * 4]);
            float4 b2 = *reinterpret_cast<const float4*>(&B[(kk + 2) * BN + tx * 4]);
            float4 b3 = *reinterpret_cast<const float4*>(&B[(kk + 3) * BN + tx * 4]);
#pragma unroll
            for (int r = 0; r < TM; ++r) {
                float4 a = *reinterpret_cast<const float4*>(&A[(ty + r * TY) * IG_ALD + kk]);
                acc[r][0] += a.x * b0.x; acc[r][1] += a.x * b0.y; acc[r][2] += a.x * b0.z; acc[r][3] += a.x * b0.w;
                acc[r][0] += a.y * b1.x; acc[r][1] += a.y * b1.y; acc[r][2] += a.y * b1.z; acc[r][3] += a.y * b1.w;
                acc[r][0] += a.z * b2.x; acc[r][1] += a.z * b2.y; acc[r][2] += a.z * b2.z; acc[r][3] += a.z * b2.w;
                acc[r][0] += a.w * b3.x; acc[r][1] += a.w * b3.y; acc[r][2] += a.w * b3.z; acc[r][3] += a.w * b3.w;
            }
        }
        if (more) {
            float* An = As + (cur ^ 1) * IG_BM * IG_ALD;
            float* Bn = Bs + (cur ^ 1) * IG_BK * BN;
            *reinterpret_cast<float4*>(&An[(a_row0)*IG_ALD + a_kq]) = ar[0];
            *reinterpret_cast<float4*>(&An[(a_row0 + 64) * IG_ALD + a_kq]) = ar[1];
            if (b_act) *reinterpret_cast<float4*>(&Bn[b_k * BN + b_c4]) = br;
        }
        __syncthreads();
    }

    // epilogue
    const int co = n0 + tx * 4;
    if (co < ep.Cout) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.scale) sc = ldg4(ep.scale + co);
        if (ep.shift) sh = ldg4(ep.shift + co);
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int p = m0 + ty + r * TY;
            if (p < M) {
                float4 v;
                v.x = acc[r][0] * sc.x + sh.x; v.y = acc[r][1] * sc.y + sh.y;
                v.z = acc[r][2] * sc.z + sh.z; v.w = acc[r][3] * sc.w + sh.w;
                if (ep.residual) {
                    float4 rr = ldg4(ep.residual + (size_t)p * ep.res_pitch + co);
                    v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
                }
                if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4*>(ep.out + (size_t)p * ep.out_pitch + co) = v;
            }
        }
    }
}

template <class AP, int BN>
inline int launch_igemm_bn(const AP& ap, const float* Wp, int ldw, int K, int M, const Epilogue& ep, cudaStream_t st) {
    size_t smem = (size_t)(2 * IG_BM * IG_ALD + 2 * IG_BK * BN) * sizeof(float) + AP::kExtraSmemBytes;
    static size_t granted[DEFT_MAX_DEVICES] = {};  // per template instantiation and device
    {
        const int rc_attr = ensure_dyn_smem(igemm_simt_kernel<AP, BN>, smem, granted, "igemm_simt_kernel");
        if (rc_attr != DEFT_OK) return rc_attr;
    }
    dim3 grid(ceil_div(M, IG_BM), ldw / BN);
    igemm_simt_kernel<AP, BN><<<grid, IG_THREADS, smem, st>>>(ap, Wp, ldw, K, M, ep);
    DEFT_CHECK_LAUNCH("igemm_simt_kernel");
    return DEFT_OK;
}

// Picks the widest N tile that divides ldw (ldw is a multiple of 16 by the packing contract).
template <class AP>
inline int launch_igemm(const AP& ap, const float* Wp, int ldw, int K, int M, const Epilogue& ep, cudaStream_t st) {
    if (K % IG_BK != 0 || ldw % 16 != 0 || ep.Cout % 4 != 0 || ep.Cout > ldw || M <= 0) {
        set_error("igemm: unsupported shape K=%d ldw=%d Cout=%d M=%d", K, ldw, ep.Cout, M);
        return DEFT_EINVAL;
    }
    if (ldw % 64 == 0) return launch_igemm_bn<AP, 64>(ap, Wp, ldw, K, M, ep, st);
    if (ldw % 32 == 0) return launch_igemm_bn<AP, 32>(ap, Wp, ldw, K, M, ep, st);
    return launch_igemm_bn<AP, 16>(ap, Wp, ldw, K, M, ep, st);
}

}  // namespace deft
