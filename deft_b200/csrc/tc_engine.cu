// tcgen05 (5th-gen tensor core) implicit-GEMM engine for sm_100a.
//
//   D[128 px, BN couts] (fp32, TMEM)  +=  A[128 x 32] (smem, K-major, SWIZZLE_128B)  x  B[BN x 32] (smem, K-major)
//
// * A is never materialised in HBM: 8 producer warps build each 128x32 tile on the fly (im2col over NHWC
//   sources, the DCNv2 bilinear gather, or the affinity pair expansion) and write it straight into the UMMA
//   canonical layout.  Global loads are software-pipelined in registers (issue for k-block kb+D while kb is
//   converted/stored), so the mainloop runs at the tensor pipe's pace, not at L2 latency.
// * B (weights) is pre-packed by tc_pack_weights_kernel as ready-made swizzled smem tile images, streamed in
//   by one elected thread with cp.async.bulk (TMA bulk copy, mbarrier complete_tx).
// * fp32 parity: 3xTF32 error compensation -- x = hi + lo (hi = rna_tf32(x)), D += hi*hi + lo*hi + hi*lo, three
//   kind::tf32 MMAs per k-step into the same TMEM accumulator.  DEFT_FLAG_TF32X1 issues only hi*hi.
// * Small-M layers (DLA levels 4/5 at 1/16, 1/32 resolution) are split along K over blockIdx.z; partials go
//   to caller scratch and a deterministic reduce kernel applies the epilogue (no atomics).
// * Warp roles: warps 0-7 A producers then epilogue (tcgen05.ld -> epilogue functor), warp 8 lane 0 MMA
//   issuer, warp 9 TMEM allocator + lane 0 weight-tile bulk copies.  mbarrier ring of S stages:
//   full_a (256 producer arrivals), full_b (tx bytes), empty (tcgen05.commit).
#include "tc_engine.cuh"

namespace deft {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // one 128-byte swizzle row of fp32
constexpr int TC_PRODUCERS = 256;
constexpr int TC_THREADS = 320;                 // 8 producer/epilogue warps + MMA warp + TMEM/B warp
constexpr int TC_A_TILE = TC_BM * 128;          // bytes
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_NUM_SMS = 148;

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("deft tc_engine: mbarrier timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, 128-byte rows, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (ignored for swizzled K-major; canonical value 1)
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO = 1024 B
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, A and B K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
// byte offset of element (row, 16B-chunk) inside a swizzled [rows x 128 B] tile
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// ------------------------------------------------------------------------------------------------ A producers
// Thread t (0..255) owns 16-byte chunk c = t & 7 of rows r_i = (t >> 3) + 32 i, i = 0..3, of every k-block.
// issue(kb, raw): start the global loads of k-block kb into registers; combine(kb, raw, v): turn them into
// the 4 A-tile chunks.  DEPTH = how many k-blocks of loads are kept in flight.
struct TcConvAP {
    ConvAP a;
    int K;
    int perm_taps;       // 0 = tap-major K order, else KH*KW with channel-block-major order (tc_k_perm_taps)
    static constexpr int kExtraSmemBytes = 0;
    static constexpr int RAW = 4;
    static constexpr int DEPTH = 2;
    int n_[4], oy_[4], ox_[4];
    unsigned valid_;
    __device__ __forceinline__ void init(int m0, int M, unsigned char*) {
        valid_ = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int p = m0 + (threadIdx.x >> 3) + 32 * i;
            bool v = p < M;
            if (v) valid_ |= 1u << i;
            int q = v ? p : 0;
            n_[i] = q / (a.OH * a.OW);
            int r = q - n_[i] * (a.OH * a.OW);
            oy_[i] = (r / a.OW) * a.stride - a.pad;
            ox_[i] = (r - (r / a.OW) * a.OW) * a.stride - a.pad;
        }
    }
    __device__ __forceinline__ void issue(int kb, float4 (&raw)[RAW]) const {
        int tap, cc;
        bool kok = true;
        if (perm_taps) {          // K order = (32-channel block, tap, channel): the 9 taps of a block hit L1
            const int cb = kb / perm_taps;
            tap = kb - cb * perm_taps;
            cc = cb * TC_BK + (threadIdx.x & 7) * 4;
        } else {                  // K order = (tap, channel)
            const int k = kb * TC_BK + (threadIdx.x & 7) * 4;
            tap = k / a.Ctot;
            cc = k - tap * a.Ctot;
            kok = k < K;
        }
        int ky = tap / a.KW, kx = tap - ky * a.KW;
        int s = 0;
        while (s < a.nsrc - 1 && cc >= a.c[s]) { cc -= a.c[s]; ++s; }
        const float* base = a.src[s];
        const int pitch = a.pitch[s];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int iy = oy_[i] + ky;
            int ix = ox_[i] + kx;
            raw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kok && ((valid_ >> i) & 1u) && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                raw[i] = ldg4(base + (((size_t)n_[i] * a.H + iy) * a.W + ix) * pitch + cc);
        }
    }
    __device__ __forceinline__ void combine(int, const float4 (&raw)[RAW], float4 (&v)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = raw[i];
    }
};

struct TcDcnAP {
    DcnAP a;
    int K;
    static constexpr int kExtraSmemBytes = TC_BM * 9 * 32;
    static constexpr int RAW = 16;
    static constexpr int DEPTH = 1;
    int4* tabO_;
    float4* tabW_;
    __device__ __forceinline__ void init(int m0, int M, unsigned char* extra) {
        tabO_ = reinterpret_cast<int4*>(extra);
        tabW_ = reinterpret_cast<float4*>(extra + TC_BM * 9 * 16);
        const int H = a.H, W = a.W;
        for (int idx = threadIdx.x; idx < TC_BM * 9; idx += TC_PRODUCERS) {
            int row = idx / 9, tap = idx - row * 9;
            int p = m0 + row;
            int4 o = make_int4(0, 0, 0, 0);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < M) {
                int n = p / (H * W);
                int r = p - n * (H * W);
                int y = r / W, xx = r - y * W;
                const float* q = a.om + (size_t)p * 32;
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
            tabO_[idx] = o;       // corners with zero weight point at pixel 0: always a valid address
            tabW_[idx] = w;
        }
        // producers-only barrier (named barrier 1, 256 threads): the table is private to them
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    __device__ __forceinline__ void issue(int kb, float4 (&raw)[RAW]) const {
        // K order = (32-channel block, tap, channel): 9 consecutive k-blocks re-read the same neighbourhood -> L1 hits
        const int cb = kb / 9;
        const int tap = kb - cb * 9;
        const int cc = cb * TC_BK + (threadIdx.x & 7) * 4;
        const float* x = a.x + cc;
        const int pitch = a.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int4 o = tabO_[((threadIdx.x >> 3) + 32 * i) * 9 + tap];
            raw[i * 4 + 0] = ldg4(x + (size_t)o.x * pitch);
            raw[i * 4 + 1] = ldg4(x + (size_t)o.y * pitch);
            raw[i * 4 + 2] = ldg4(x + (size_t)o.z * pitch);
            raw[i * 4 + 3] = ldg4(x + (size_t)o.w * pitch);
        }
    }
    __device__ __forceinline__ void combine(int kb, const float4 (&raw)[RAW], float4 (&v)[4]) const {
        const int tap = kb % 9;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 w = tabW_[((threadIdx.x >> 3) + 32 * i) * 9 + tap];
            float4 acc;
            acc.x = w.x * raw[i * 4].x; acc.y = w.x * raw[i * 4].y; acc.z = w.x * raw[i * 4].z; acc.w = w.x * raw[i * 4].w;
            acc.x += w.y * raw[i * 4 + 1].x; acc.y += w.y * raw[i * 4 + 1].y; acc.z += w.y * raw[i * 4 + 1].z; acc.w += w.y * raw[i * 4 + 1].w;
            acc.x += w.z * raw[i * 4 + 2].x; acc.y += w.z * raw[i * 4 + 2].y; acc.z += w.z * raw[i * 4 + 2].z; acc.w += w.z * raw[i * 4 + 2].w;
            acc.x += w.w * raw[i * 4 + 3].x; acc.y += w.w * raw[i * 4 + 3].y; acc.z += w.w * raw[i * 4 + 3].z; acc.w += w.w * raw[i * 4 + 3].w;
            v[i] = acc;
        }
    }
};

// Affinity MLP layer 2: A[p, k] = relu(Au'[f(p), i(p), k] + Bv'[j(p), k]) over the live pairs.
struct TcPairAP {
    PairAP a;
    int K;
    static constexpr int kExtraSmemBytes = 0;
    static constexpr int RAW = 8;
    static constexpr int DEPTH = 2;
    const float* pa_[4];
    const float* pb_[4];
    __device__ __forceinline__ void init(int m0, int M, unsigned char*) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int p = m0 + (threadIdx.x >> 3) + 32 * i;
            pa_[i] = nullptr;
            pb_[i] = nullptr;
            if (p < M) {
                int f = 0;
                while (f < a.F - 1 && p >= a.off[f + 1]) ++f;
                int q = p - a.off[f];
                int ii = q / a.m, jj = q - ii * a.m;
                pa_[i] = a.au + ((size_t)f * a.P + ii) * a.ld;
                pb_[i] = a.bv + (size_t)jj * a.ld;
            }
        }
    }
    __device__ __forceinline__ void issue(int kb, float4 (&raw)[RAW]) const {
        const int k = kb * TC_BK + (threadIdx.x & 7) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            raw[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
            raw[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pa_[i] != nullptr && k < K) {
                raw[2 * i] = ldg4(pa_[i] + k);
                raw[2 * i + 1] = ldg4(pb_[i] + k);
            }
        }
    }
    __device__ __forceinline__ void combine(int, const float4 (&raw)[RAW], float4 (&v)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i].x = fmaxf(raw[2 * i].x + raw[2 * i + 1].x, 0.f); v[i].y = fmaxf(raw[2 * i].y + raw[2 * i + 1].y, 0.f);
            v[i].z = fmaxf(raw[2 * i].z + raw[2 * i + 1].z, 0.f); v[i].w = fmaxf(raw[2 * i].w + raw[2 * i + 1].w, 0.f);
        }
    }
};

// ------------------------------------------------------------------------------------------------ epilogues
// Interface: kSmemBytes; setup(smem, tid) by the 256 producer threads before the mainloop;
// run<BN>(tmem_base, smem, warp, lane, m0, M, ntile) by warps 0-7 once the accumulator is complete.
struct EpiStd {
    Epilogue e;
    float* partial;      // split-K scratch [splits][M][ldp] or nullptr
    int ldp;
    static constexpr int kSmemBytes = 0;
    __device__ __forceinline__ void setup(unsigned char*, int) const {}
    template <int BN>
    __device__ __forceinline__ void run(uint32_t tmem_base, unsigned char*, int warp, int lane, int m0, int M, int ntile) const {
        const int lane_base = (warp & 3) * 32;
        const int p = m0 + lane_base + lane;
        constexpr int COLS_PER_HALF = BN / 2;     // warps 0-3: first half of the columns, warps 4-7: second half
        const int col0 = (warp >> 2) * COLS_PER_HALF;
#pragma unroll 1
        for (int c = 0; c < COLS_PER_HALF; c += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(col0 + c), v);
            const int co = ntile * BN + col0 + c;
            if (p >= M) continue;
            if (partial != nullptr) {
                float* dst = partial + ((size_t)blockIdx.z * M + p) * ldp + co;
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                continue;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int cq = co + q * 4;
                if (cq >= e.Cout) break;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e.scale) sc = ldg4(e.scale + cq);
                if (e.shift) sh = ldg4(e.shift + cq);
                float4 o;
                o.x = v[q * 4 + 0] * sc.x + sh.x; o.y = v[q * 4 + 1] * sc.y + sh.y;
                o.z = v[q * 4 + 2] * sc.z + sh.z; o.w = v[q * 4 + 3] * sc.w + sh.w;
                if (e.residual) {
                    float4 rr = ldg4(e.residual + (size_t)p * e.res_pitch + cq);
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                if (e.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(e.out + (size_t)p * e.out_pitch + cq) = o;
            }
        }
    }
};

// CenterNet heads, fused: N tile = one head's 256 hidden channels.  hidden = relu(acc + b1) never leaves the SM;
// the head's 1x1 conv (256 -> C_h <= 16) is a per-row dot product against w2 staged in smem; planar NCHW store.
constexpr int HEADS_MAXC = 16;
struct EpiHeads {
    const float* b1;     // [nheads*256]
    const float* w2;     // [sum C_h][256]
    const float* b2;     // [sum C_h]
    float* out[DEFT_MAX_HEADS];
    int c[DEFT_MAX_HEADS];
    int row0[DEFT_MAX_HEADS];
    int HW;
    static constexpr int kSmemBytes = HEADS_MAXC * 256 * 4 + 256 * 4 + HEADS_MAXC * TC_BM * 4;
    __device__ __forceinline__ void setup(unsigned char* smem, int tid) const {
        float* w2s = reinterpret_cast<float*>(smem);
        float* b1s = w2s + HEADS_MAXC * 256;
        const int h = blockIdx.y;
        const int ch = c[h];
        for (int i = tid; i < ch * 256; i += TC_PRODUCERS) w2s[i] = w2[(size_t)row0[h] * 256 + i];
        b1s[tid] = b1[h * 256 + tid];
        // visibility to all epilogue warps is guaranteed by the producers' later named barrier in run()
    }
    template <int BN>
    __device__ __forceinline__ void run(uint32_t tmem_base, unsigned char* smem, int warp, int lane, int m0, int M, int ntile) const {
        static_assert(BN == 256, "EpiHeads needs one 256-wide tile per head");
        const float* w2s = reinterpret_cast<const float*>(smem);
        const float* b1s = w2s + HEADS_MAXC * 256;
        float* part = const_cast<float*>(b1s) + 256;          // [HEADS_MAXC][128] partials of the upper column half
        const int h = ntile;
        const int ch = c[h];
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const int p = m0 + row;
        const int col0 = (warp >> 2) * 128;
        float acc[HEADS_MAXC];
#pragma unroll
        for (int j = 0; j < HEADS_MAXC; ++j) acc[j] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");        // w2s / b1s written by setup() are visible
#pragma unroll 1
        for (int cb = 0; cb < 128; cb += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(col0 + cb), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + b1s[col0 + cb + j], 0.f);
#pragma unroll
            for (int o = 0; o < HEADS_MAXC; ++o) {
                if (o < ch) {
                    const float4 wa = *reinterpret_cast<const float4*>(w2s + o * 256 + col0 + cb);
                    const float4 wb = *reinterpret_cast<const float4*>(w2s + o * 256 + col0 + cb + 4);
                    acc[o] += v[0] * wa.x + v[1] * wa.y + v[2] * wa.z + v[3] * wa.w + v[4] * wb.x + v[5] * wb.y + v[6] * wb.z + v[7] * wb.w;
                }
            }
        }
        if (warp >= 4) {
#pragma unroll
            for (int o = 0; o < HEADS_MAXC; ++o)
                if (o < ch) part[o * TC_BM + row] = acc[o];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp < 4 && p < M) {
            const int n = p / HW, q = p - n * HW;
#pragma unroll
            for (int o = 0; o < HEADS_MAXC; ++o)
                if (o < ch) out[h][((size_t)n * ch + o) * HW + q] = acc[o] + part[o * TC_BM + row] + b2[row0[h] + o];
        }
    }
};

// ------------------------------------------------------------------------------------------------ kernel
template <class AP, int BN, int NPASS, class EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_igemm_kernel(AP ap, const float* __restrict__ wblob, int nkb_total, int kb_per_split,
                                                                 int M, int stages, EPI epi) {
    constexpr int B_TILE = BN * 128;                      // bytes of one B part (hi or lo)
    constexpr int STAGE_BYTES = NPASS == 3 ? 2 * (TC_A_TILE + B_TILE) : (TC_A_TILE + B_TILE);
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned carve-up (SWIZZLE_128B atoms must sit on 1024 B boundaries)
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    unsigned char* smem = smem_raw + pad;
    const uint32_t sbase = raw_addr + pad;
    // layout: [stages x STAGE_BYTES] [barriers + tmem slot: 256 B] [AP extra] [EPI extra]
    const uint32_t bar_base = sbase + (uint32_t)stages * STAGE_BYTES;
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto full_b = [&](int s) { return bar_base + 8u * (8 + s); };
    auto empty = [&](int s) { return bar_base + 8u * (16 + s); };
    const uint32_t accum_bar = bar_base + 8u * 24;
    const uint32_t tmem_slot = bar_base + 8u * 25;
    unsigned char* ap_extra = smem + (size_t)stages * STAGE_BYTES + 256;
    unsigned char* epi_extra = ap_extra + AP::kExtraSmemBytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TC_BM;
    const int ntile = blockIdx.y;
    const int kb0 = blockIdx.z * kb_per_split;
    const int kb1 = min(nkb_total, kb0 + kb_per_split);
    const int nkb = kb1 - kb0;                            // >= 1 by construction of the split

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_a(s), TC_PRODUCERS);
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (size_t)stages * STAGE_BYTES + 8 * 25);

    if (warp < 8) {
        // ===================== A producers =====================
        ap.init(m0, M, ap_extra);
        epi.setup(epi_extra, tid);
        const int chunk = tid & 7;
        const int row0 = tid >> 3;
        float4 raw[AP::DEPTH][AP::RAW];
#pragma unroll
        for (int d = 0; d < AP::DEPTH; ++d)
            if (d < nkb) ap.issue(kb0 + d, raw[d]);
        for (int it = 0; it < nkb; it += AP::DEPTH) {
#pragma unroll
            for (int d = 0; d < AP::DEPTH; ++d) {
                const int i = it + d;                     // local k-block index
                if (i < nkb) {
                    float4 v[4];
                    ap.combine(kb0 + i, raw[d], v);
                    if (i + AP::DEPTH < nkb) ap.issue(kb0 + i + AP::DEPTH, raw[d]);
                    const int s = i % stages;
                    const uint32_t ph = (uint32_t)(i / stages) & 1u;
                    mbar_wait(empty(s), ph ^ 1u);
                    unsigned char* a_hi = smem + (size_t)s * STAGE_BYTES;
                    unsigned char* a_lo = a_hi + TC_A_TILE;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const uint32_t off = swz(row0 + 32 * r, chunk);
                        float4 h;
                        h.x = tf32_rna(v[r].x); h.y = tf32_rna(v[r].y); h.z = tf32_rna(v[r].z); h.w = tf32_rna(v[r].w);
                        *reinterpret_cast<float4*>(a_hi + off) = h;
                        if (NPASS == 3) {
                            float4 l;
                            l.x = v[r].x - h.x; l.y = v[r].y - h.y; l.z = v[r].z - h.z; l.w = v[r].w - h.w;
                            *reinterpret_cast<float4*>(a_lo + off) = l;
                        }
                    }
                    fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
                    mbar_arrive(full_a(s));
                }
            }
        }
        // ===================== epilogue =====================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        epi.template run<BN>(tmem_base, epi_extra, warp, lane, m0, M, ntile);
        tc_fence_before();
    } else if (warp == 8) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                const uint32_t a_hi = sbase + (uint32_t)s * STAGE_BYTES;
                const uint32_t a_lo = a_hi + TC_A_TILE;
                const uint32_t b_hi = a_hi + (NPASS == 3 ? 2 * TC_A_TILE : TC_A_TILE);
                const uint32_t b_lo = b_hi + B_TILE;
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t da_hi = make_smem_desc(a_hi + k * 32), db_hi = make_smem_desc(b_hi + k * 32);
                    if (NPASS == 3) {
                        // small terms first, so the big hi*hi product is added to an already-formed correction
                        const uint64_t da_lo = make_smem_desc(a_lo + k * 32), db_lo = make_smem_desc(b_lo + k * 32);
                        umma_tf32(tmem_base, da_lo, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                        umma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, da_hi, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(empty(s));            // slot free once these MMAs have read it
            }
            umma_commit(accum_bar);               // accumulator complete
        }
        __syncwarp();
    } else {
        // ===================== weight-tile bulk copies (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t bytes = NPASS == 3 ? 2 * B_TILE : B_TILE;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(wblob) + ((size_t)ntile * nkb_total + kb0) * (2 * B_TILE);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                mbar_wait(empty(s), ph ^ 1u);
                const uint32_t dst = sbase + (uint32_t)s * STAGE_BYTES + (NPASS == 3 ? 2 * TC_A_TILE : TC_A_TILE);
                mbar_arrive_expect_tx(full_b(s), bytes);
                bulk_g2s(dst, src + (size_t)i * (2 * B_TILE), bytes, full_b(s));
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// Deterministic split-K tail: out = epilogue(sum_z partial[z]).
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int ldp, Epilogue e) {
    const int c4 = e.Cout / 4;
    size_t total = (size_t)M * c4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int p = (int)(i / c4), co = (int)(i - (size_t)p * c4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < splits; ++z) {
            float4 t = ldg4(partial + ((size_t)z * M + p) * ldp + co);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.scale) sc = ldg4(e.scale + co);
        if (e.shift) sh = ldg4(e.shift + co);
        float4 o;
        o.x = acc.x * sc.x + sh.x; o.y = acc.y * sc.y + sh.y; o.z = acc.z * sc.z + sh.z; o.w = acc.w * sc.w + sh.w;
        if (e.residual) {
            float4 rr = ldg4(e.residual + (size_t)p * e.res_pitch + co);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (e.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(e.out + (size_t)p * e.out_pitch + co) = o;
    }
}

// Weight packing: SIMT layout Wp[K][ldw] -> [n_tiles][nkb][hi|lo][BN rows x 128 B swizzled], split hi/lo.
// perm_taps > 0: the blob's K order is (32-channel block, tap, channel) instead of (tap, channel) -- see tc_k_perm_taps.
__global__ void tc_pack_weights_kernel(const float* __restrict__ wp, int ldw, int K, int BN, int nkb, int n_tiles, int perm_taps,
                                       float* __restrict__ blob) {
    size_t total = (size_t)n_tiles * nkb * BN * TC_BK;
    const int ctot = perm_taps ? K / perm_taps : 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int kk = (int)(i % TC_BK);
        size_t t = i / TC_BK;
        int n = (int)(t % BN);
        t /= BN;
        int kb = (int)(t % nkb);
        int nt = (int)(t / nkb);
        int k = kb * TC_BK + kk, co = nt * BN + n;
        if (perm_taps) {
            int cb = kb / perm_taps, tap = kb - cb * perm_taps;
            k = tap * ctot + cb * TC_BK + kk;
        }
        float w = (k < K && co < ldw) ? wp[(size_t)k * ldw + co] : 0.f;
        float hi = tf32_rna(w);
        float lo = tf32_rna(w - hi);
        size_t tile = ((size_t)nt * nkb + kb) * (size_t)(2 * BN * TC_BK);
        size_t off = (size_t)n * 32 + (size_t)(((kk >> 2) ^ (n & 7)) << 2) + (kk & 3);
        blob[tile + off] = hi;
        blob[tile + (size_t)BN * TC_BK + off] = lo;
    }
}

// ------------------------------------------------------------------------------------------------ host
// Tile-N choice.  Depends only on (ldw, M, kind) so the packer and the dispatcher always agree; the chosen
// value also travels explicitly with every call (tc_bn).
// K permutation rule shared by the packer and every dispatcher: multi-tap convs whose channel count is a
// multiple of 32 run channel-block-major (returns the tap count), everything else tap-major (returns 0).
int tc_k_perm_taps(int K, int taps) { return (taps > 1 && K % taps == 0 && (K / taps) % TC_BK == 0) ? taps : 0; }

int tc_pick_bn(int ldw, int M, int kind) {
    if (kind == TC_KIND_HEADS) return 256;
    const int max_bn = (kind == TC_KIND_DCN) ? 64 : 128;
    const int mt = ceil_div(M, TC_BM);
    int best = 0;
    for (int bn = max_bn; bn >= 32; bn >>= 1) {
        if (ldw % bn != 0) continue;
        if (best == 0) best = bn;                        // largest dividing tile
        if (mt * (ldw / bn) >= 120) return bn;           // enough CTAs for one wave
        if (bn == 64) break;                             // below 64 the redundant A gathers cost more than split-K
    }
    if (best == 0) best = (ldw % 32 == 0) ? 32 : 16;
    return best;                                         // too few CTAs either way: keep the widest tile, split K instead
}
static int tc_pick_splits(int ctas, int nkb) {
    if (ctas >= 100) return 1;
    int s = TC_NUM_SMS / ctas;
    int max_s = nkb / 6;                                 // keep >= 6 k-blocks per split
    if (s > max_s) s = max_s;
    if (s > 16) s = 16;
    return s < 1 ? 1 : s;
}
size_t tc_blob_floats(int ldw, int K, int bn) {
    int nkb = ceil_div(K, TC_BK);
    return (size_t)(ldw / bn) * nkb * 2 * bn * TC_BK;
}
size_t tc_splitk_scratch_bytes(int ldw, int K, int M, int kind) {
    const int bn = tc_pick_bn(ldw, M, kind);
    const int splits = tc_pick_splits(ceil_div(M, TC_BM) * (ldw / bn), ceil_div(K, TC_BK));
    return splits > 1 ? align_up((size_t)splits * M * ldw * sizeof(float), 256) : 0;
}

template <class AP, int BN, int NPASS, class EPI>
static int tc_launch(AP ap, const float* blob, int nkb, int M, int n_tiles, int splits, EPI epi, cudaStream_t st) {
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = NPASS == 3 ? 2 * (TC_A_TILE + B_TILE) : (TC_A_TILE + B_TILE);
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + AP::kExtraSmemBytes + EPI::kSmemBytes;
    const int kb_per_split = ceil_div(nkb, splits);
    splits = ceil_div(nkb, kb_per_split);                // no empty split
    int stages = (TC_SMEM_LIMIT - fixed) / STAGE_BYTES;
    if (stages > 4) stages = 4;
    if (stages > kb_per_split) stages = kb_per_split;
    if (stages < 1 || (stages < 2 && kb_per_split > 1)) return DEFT_EINVAL;
    size_t smem = (size_t)stages * STAGE_BYTES + fixed;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc_igemm_kernel<AP, BN, NPASS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT) != cudaSuccess) {
            set_error("tc_engine: cannot raise dynamic smem limit");
            return DEFT_ECUDA;
        }
        attr_done = true;
    }
    dim3 grid(ceil_div(M, TC_BM), n_tiles, splits);
    tc_igemm_kernel<AP, BN, NPASS, EPI><<<grid, TC_THREADS, smem, st>>>(ap, blob, nkb, kb_per_split, M, stages, epi);
    DEFT_CHECK_LAUNCH("tc_igemm_kernel");
    return DEFT_OK;
}

template <class AP>
static int tc_dispatch(AP ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                       size_t scratch_bytes, cudaStream_t st) {
    if (blob == nullptr || bn <= 0 || ldw % bn != 0 || ep.Cout % 4 != 0) return DEFT_EINVAL;
    const int nkb = ceil_div(K, TC_BK), nt = ldw / bn;
    int splits = tc_pick_splits(ceil_div(M, TC_BM) * nt, nkb);
    if (splits > 1 && (scratch == nullptr || scratch_bytes < (size_t)splits * M * ldw * sizeof(float))) splits = 1;
    if (splits > 1) splits = ceil_div(nkb, ceil_div(nkb, splits));
    EpiStd epi{ep, splits > 1 ? reinterpret_cast<float*>(scratch) : nullptr, ldw};
    const bool x1 = (flags & DEFT_FLAG_TF32X1) != 0;
    int rc = DEFT_EINVAL;
#define TC_CASE(BNV)                                                                                             \
    case BNV:                                                                                                    \
        rc = x1 ? tc_launch<AP, BNV, 1, EpiStd>(ap, blob, nkb, M, nt, splits, epi, st)                           \
                : tc_launch<AP, BNV, 3, EpiStd>(ap, blob, nkb, M, nt, splits, epi, st);                          \
        break;
    switch (bn) {
        TC_CASE(128)
        TC_CASE(64)
        TC_CASE(32)
        TC_CASE(16)
    }
#undef TC_CASE
    if (rc != DEFT_OK || splits == 1) return rc;
    size_t total = (size_t)M * (ep.Cout / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > TC_NUM_SMS * 8) blocks = TC_NUM_SMS * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(scratch), splits, M, ldw, ep);
    DEFT_CHECK_LAUNCH("splitk_reduce");
    return DEFT_OK;
}

int tc_conv_dispatch(const ConvAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                     size_t scratch_bytes, cudaStream_t st) {
    for (int i = 0; i < ap.nsrc; ++i)
        if (ap.c[i] % 4 != 0) return DEFT_EINVAL;
    TcConvAP t{};
    t.a = ap;
    t.K = K;
    t.perm_taps = tc_k_perm_taps(K, ap.KH * ap.KW);
    return tc_dispatch(t, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
}

int tc_dcn_dispatch(const DcnAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                    size_t scratch_bytes, cudaStream_t st) {
    if (ap.Cin % 32 != 0 || bn > 64) return DEFT_EINVAL;
    TcDcnAP t{};
    t.a = ap;
    t.K = K;
    return tc_dispatch(t, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
}

int tc_pair_dispatch(const PairAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st) {
    TcPairAP t{};
    t.a = ap;
    t.K = K;
    return tc_dispatch(t, blob, bn, ldw, K, M, ep, flags, nullptr, 0, st);
}

int tc_heads_dispatch(const ConvAP& ap, const float* blob, int K, int M, const HeadsArgs& ha, unsigned flags, cudaStream_t st) {
    if (blob == nullptr || ha.nheads < 1) return DEFT_EINVAL;
    EpiHeads epi{};
    epi.b1 = ha.b1; epi.w2 = ha.w2; epi.b2 = ha.b2; epi.HW = ha.HW;
    for (int h = 0; h < ha.nheads; ++h) {
        if (ha.c[h] > HEADS_MAXC) return DEFT_EINVAL;
        epi.out[h] = ha.out[h]; epi.c[h] = ha.c[h]; epi.row0[h] = ha.row0[h];
    }
    TcConvAP t{};
    t.a = ap;
    t.K = K;
    t.perm_taps = tc_k_perm_taps(K, ap.KH * ap.KW);
    const int nkb = ceil_div(K, TC_BK);
    if (flags & DEFT_FLAG_TF32X1) return tc_launch<TcConvAP, 256, 1, EpiHeads>(t, blob, nkb, M, ha.nheads, 1, epi, st);
    return tc_launch<TcConvAP, 256, 3, EpiHeads>(t, blob, nkb, M, ha.nheads, 1, epi, st);
}

int tc_pack_weights(const float* wp, int ldw, int K, int bn, int taps, float* blob, cudaStream_t st) {
    int nkb = ceil_div(K, TC_BK), nt = ldw / bn;
    size_t total = (size_t)nt * nkb * bn * TC_BK;
    int blocks = (int)((total + 255) / 256);
    if (blocks > TC_NUM_SMS * 8) blocks = TC_NUM_SMS * 8;
    tc_pack_weights_kernel<<<blocks, 256, 0, st>>>(wp, ldw, K, bn, nkb, nt, tc_k_perm_taps(K, taps), blob);
    DEFT_CHECK_LAUNCH("tc_pack_weights");
    return DEFT_OK;
}

}  // namespace deft
