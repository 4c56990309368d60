// tcgen05 (5th-gen tensor core) implicit-GEMM engine for sm_100a.
//
//   D[128 px, BN couts] (fp32, TMEM)  +=  A[128 x 32] (smem, K-major, SWIZZLE_128B)  x  B[BN x 32] (smem, K-major)
//
// * A is never materialised in HBM: 8 producer warps build each 128x32 tile on the fly (im2col over
//   NHWC sources, or the DCNv2 bilinear gather) and write it straight into the UMMA canonical layout.
// * B (weights) is pre-packed by tc_pack_weights_kernel as ready-made swizzled smem tile images, so one
//   elected thread streams it in with cp.async.bulk (TMA bulk copy, mbarrier complete_tx).
// * fp32 parity: 3xTF32 error compensation -- every operand is split x = hi + lo (hi = rna_tf32(x)),
//   D += hi*hi + lo*hi + hi*lo, three kind::tf32 MMAs per k-step into the same TMEM accumulator
//   (the dropped lo*lo term is O(2^-22)).  DEFT_FLAG_TF32X1 issues only hi*hi.
// * Warp roles: warps 0-7 A producers then epilogue (tcgen05.ld -> scale/shift/residual/ReLU -> NHWC),
//   warp 8 lane 0 MMA issuer, warp 9 TMEM allocator + lane 0 weight-tile bulk copies.
// * mbarrier ring of S stages: full_a (256 producer arrivals), full_b (tx bytes), empty (tcgen05.commit).
#include "tc_engine.cuh"

namespace deft {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // one 128-byte swizzle row of fp32
constexpr int TC_PRODUCERS = 256;
constexpr int TC_THREADS = 320;                 // 8 producer/epilogue warps + MMA warp + TMEM/B warp
constexpr int TC_A_TILE = TC_BM * 128;          // bytes
constexpr int TC_SMEM_LIMIT = 227 * 1024;

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
            printf("deft tc_engine: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
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
struct TcConvAP {
    ConvAP a;
    int K;
    static constexpr int kExtraSmemBytes = 0;
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
            oy_[i] = r / a.OW;
            ox_[i] = r - oy_[i] * a.OW;
        }
    }
    __device__ __forceinline__ void load4(int kb, float4 (&v)[4]) const {
        int k = kb * TC_BK + (threadIdx.x & 7) * 4;
        int tap = k / a.Ctot;
        int cc = k - tap * a.Ctot;
        int ky = tap / a.KW, kx = tap - ky * a.KW;
        int s = 0;
        while (s < a.nsrc - 1 && cc >= a.c[s]) { cc -= a.c[s]; ++s; }
        const float* base = a.src[s];
        const int pitch = a.pitch[s];
        const bool kok = k < K;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int iy = oy_[i] * a.stride - a.pad + ky;
            int ix = ox_[i] * a.stride - a.pad + kx;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kok && ((valid_ >> i) & 1u) && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                v[i] = ldg4(base + (((size_t)n_[i] * a.H + iy) * a.W + ix) * pitch + cc);
        }
    }
};

struct TcDcnAP {
    DcnAP a;
    int K;
    static constexpr int kExtraSmemBytes = TC_BM * 9 * 32;
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
            tabO_[idx] = o;
            tabW_[idx] = w;
        }
        // producers-only barrier (named barrier 1, 256 threads): the table is private to them
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    __device__ __forceinline__ void load4(int kb, float4 (&v)[4]) const {
        int k = kb * TC_BK + (threadIdx.x & 7) * 4;
        int tap = k / a.Cin;
        int cc = k - tap * a.Cin;
        const float* x = a.x;
        const int pitch = a.pitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = (threadIdx.x >> 3) + 32 * i;
            int4 o = tabO_[row * 9 + tap];
            float4 w = tabW_[row * 9 + tap];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w.x != 0.f) { float4 t = ldg4(x + (size_t)o.x * pitch + cc); acc.x += w.x * t.x; acc.y += w.x * t.y; acc.z += w.x * t.z; acc.w += w.x * t.w; }
            if (w.y != 0.f) { float4 t = ldg4(x + (size_t)o.y * pitch + cc); acc.x += w.y * t.x; acc.y += w.y * t.y; acc.z += w.y * t.z; acc.w += w.y * t.w; }
            if (w.z != 0.f) { float4 t = ldg4(x + (size_t)o.z * pitch + cc); acc.x += w.z * t.x; acc.y += w.z * t.y; acc.z += w.z * t.z; acc.w += w.z * t.w; }
            if (w.w != 0.f) { float4 t = ldg4(x + (size_t)o.w * pitch + cc); acc.x += w.w * t.x; acc.y += w.w * t.y; acc.z += w.w * t.z; acc.w += w.w * t.w; }
            v[i] = acc;
        }
    }
};

// Affinity MLP layer 2: A[p, k] = relu(Au'[f(p), i(p), k] + Bv'[j(p), k]) over the live pairs.
struct TcPairAP {
    PairAP a;
    int K;
    static constexpr int kExtraSmemBytes = 0;
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
    __device__ __forceinline__ void load4(int kb, float4 (&v)[4]) const {
        const int k = kb * TC_BK + (threadIdx.x & 7) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pa_[i] != nullptr && k < K) {
                float4 x = ldg4(pa_[i] + k), y = ldg4(pb_[i] + k);
                v[i].x = fmaxf(x.x + y.x, 0.f); v[i].y = fmaxf(x.y + y.y, 0.f);
                v[i].z = fmaxf(x.z + y.z, 0.f); v[i].w = fmaxf(x.w + y.w, 0.f);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ kernel
template <class AP, int BN, int NPASS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_igemm_kernel(AP ap, const float* __restrict__ wblob, int nkb, int M, int stages,
                                                                 Epilogue ep) {
    constexpr int B_TILE = BN * 128;                      // bytes of one B part (hi or lo)
    constexpr int STAGE_BYTES = NPASS == 3 ? 2 * (TC_A_TILE + B_TILE) : (TC_A_TILE + B_TILE);
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned carve-up (SWIZZLE_128B atoms must sit on 1024 B boundaries)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    unsigned char* smem = smem_raw + pad;
    const uint32_t sbase = raw + pad;
    // layout: [stages x STAGE_BYTES] [barriers 64 B-aligned] [tmem slot] [AP extra]
    const uint32_t bar_base = sbase + (uint32_t)stages * STAGE_BYTES;
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto full_b = [&](int s) { return bar_base + 8u * (8 + s); };
    auto empty = [&](int s) { return bar_base + 8u * (16 + s); };
    const uint32_t accum_bar = bar_base + 8u * 24;
    const uint32_t tmem_slot = bar_base + 8u * 25;
    unsigned char* extra = smem + (size_t)stages * STAGE_BYTES + 256;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TC_BM;
    const int ntile = blockIdx.y;

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
        ap.init(m0, M, extra);
        const int chunk = tid & 7;
        const int row0 = tid >> 3;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % stages;
            const uint32_t ph = (uint32_t)(kb / stages) & 1u;
            float4 v[4];
            ap.load4(kb, v);                      // global loads first, then wait for the slot
            mbar_wait(empty(s), ph ^ 1u);
            unsigned char* a_hi = smem + (size_t)s * STAGE_BYTES;
            unsigned char* a_lo = a_hi + TC_A_TILE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t off = swz(row0 + 32 * i, chunk);
                float4 h;
                h.x = tf32_rna(v[i].x); h.y = tf32_rna(v[i].y); h.z = tf32_rna(v[i].z); h.w = tf32_rna(v[i].w);
                *reinterpret_cast<float4*>(a_hi + off) = h;
                if (NPASS == 3) {
                    float4 l;
                    l.x = v[i].x - h.x; l.y = v[i].y - h.y; l.z = v[i].z - h.z; l.w = v[i].w - h.w;
                    *reinterpret_cast<float4*>(a_lo + off) = l;
                }
            }
            fence_proxy_async();                  // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(full_a(s));
        }
        // ===================== epilogue =====================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int lane_base = (warp & 3) * 32;
        const int p = m0 + lane_base + lane;
        constexpr int COLS_PER_HALF = BN / 2;     // warps 0-3: first half of the columns, warps 4-7: second half
        const int col0 = (warp >> 2) * COLS_PER_HALF;
#pragma unroll 1
        for (int c = 0; c < COLS_PER_HALF; c += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(col0 + c), v);
            const int co = ntile * BN + col0 + c;
            if (p < M && co < ep.Cout) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int cq = co + q * 4;
                    if (cq >= ep.Cout) break;
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ep.scale) sc = ldg4(ep.scale + cq);
                    if (ep.shift) sh = ldg4(ep.shift + cq);
                    float4 o;
                    o.x = v[q * 4 + 0] * sc.x + sh.x; o.y = v[q * 4 + 1] * sc.y + sh.y;
                    o.z = v[q * 4 + 2] * sc.z + sh.z; o.w = v[q * 4 + 3] * sc.w + sh.w;
                    if (ep.residual) {
                        float4 rr = ldg4(ep.residual + (size_t)p * ep.res_pitch + cq);
                        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                    }
                    if (ep.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(ep.out + (size_t)p * ep.out_pitch + cq) = o;
                }
            }
        }
        tc_fence_before();
    } else if (warp == 8) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
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
                    umma_tf32(tmem_base, da_hi, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    if (NPASS == 3) {
                        const uint64_t da_lo = make_smem_desc(a_lo + k * 32), db_lo = make_smem_desc(b_lo + k * 32);
                        umma_tf32(tmem_base, da_lo, db_hi, idesc, 1u);
                        umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
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
            const unsigned char* src = reinterpret_cast<const unsigned char*>(wblob) + ((size_t)ntile * nkb) * (2 * B_TILE);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                mbar_wait(empty(s), ph ^ 1u);
                const uint32_t dst = sbase + (uint32_t)s * STAGE_BYTES + (NPASS == 3 ? 2 * TC_A_TILE : TC_A_TILE);
                mbar_arrive_expect_tx(full_b(s), bytes);
                bulk_g2s(dst, src + (size_t)kb * (2 * B_TILE), bytes, full_b(s));
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

// Weight packing: SIMT layout Wp[K][ldw] -> [n_tiles][nkb][hi|lo][BN rows x 128 B swizzled], split hi/lo.
__global__ void tc_pack_weights_kernel(const float* __restrict__ wp, int ldw, int K, int BN, int nkb, int n_tiles, float* __restrict__ blob) {
    size_t total = (size_t)n_tiles * nkb * BN * TC_BK;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int kk = (int)(i % TC_BK);
        size_t t = i / TC_BK;
        int n = (int)(t % BN);
        t /= BN;
        int kb = (int)(t % nkb);
        int nt = (int)(t / nkb);
        int k = kb * TC_BK + kk, co = nt * BN + n;
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
int tc_tile_n(int ldw) {
    if (ldw % 128 == 0) return 128;
    if (ldw % 64 == 0) return 64;
    if (ldw % 32 == 0) return 32;
    return 16;
}
size_t tc_blob_floats(int ldw, int K, int bn) {
    int nkb = ceil_div(K, TC_BK);
    return (size_t)(ldw / bn) * nkb * 2 * bn * TC_BK;
}

template <class AP, int BN, int NPASS>
static int tc_launch(AP ap, const float* blob, int nkb, int M, int n_tiles, const Epilogue& ep, cudaStream_t st) {
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = NPASS == 3 ? 2 * (TC_A_TILE + B_TILE) : (TC_A_TILE + B_TILE);
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + AP::kExtraSmemBytes;
    int stages = (TC_SMEM_LIMIT - fixed) / STAGE_BYTES;
    if (stages > 4) stages = 4;
    if (stages > nkb) stages = nkb < 1 ? 1 : nkb;
    if (stages < 2 && nkb > 1) return DEFT_EINVAL;
    size_t smem = (size_t)stages * STAGE_BYTES + fixed;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc_igemm_kernel<AP, BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT) != cudaSuccess) {
            set_error("tc_engine: cannot raise dynamic smem limit");
            return DEFT_ECUDA;
        }
        attr_done = true;
    }
    dim3 grid(ceil_div(M, TC_BM), n_tiles);
    tc_igemm_kernel<AP, BN, NPASS><<<grid, TC_THREADS, smem, st>>>(ap, blob, nkb, M, stages, ep);
    DEFT_CHECK_LAUNCH("tc_igemm_kernel");
    return DEFT_OK;
}

template <class AP>
static int tc_dispatch(AP ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, int bn, cudaStream_t st) {
    if (blob == nullptr || ldw % bn != 0 || ep.Cout % 4 != 0) return DEFT_EINVAL;
    const int nkb = ceil_div(K, TC_BK), nt = ldw / bn;
    const bool x1 = (flags & DEFT_FLAG_TF32X1) != 0;
#define TC_CASE(BNV)                                                                                   \
    case BNV:                                                                                          \
        return x1 ? tc_launch<AP, BNV, 1>(ap, blob, nkb, M, nt, ep, st) : tc_launch<AP, BNV, 3>(ap, blob, nkb, M, nt, ep, st);
    switch (bn) {
        TC_CASE(128)
        TC_CASE(64)
        TC_CASE(32)
        TC_CASE(16)
    }
#undef TC_CASE
    return DEFT_EINVAL;
}

int tc_conv_dispatch(const ConvAP& ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st) {
    for (int i = 0; i < ap.nsrc; ++i)
        if (ap.c[i] % 4 != 0) return DEFT_EINVAL;
    TcConvAP t{};
    t.a = ap;
    t.K = K;
    return tc_dispatch(t, blob, ldw, K, M, ep, flags, tc_tile_n(ldw), st);
}

int tc_dcn_dispatch(const DcnAP& ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st) {
    if (ap.Cin % 32 != 0) return DEFT_EINVAL;
    TcDcnAP t{};
    t.a = ap;
    t.K = K;
    // the 36 KB corner table leaves room for 64-wide weight tiles only
    int bn = tc_tile_n(ldw) > 64 ? 64 : tc_tile_n(ldw);
    return tc_dispatch(t, blob, ldw, K, M, ep, flags, bn, st);
}

int tc_pair_dispatch(const PairAP& ap, const float* blob, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st) {
    TcPairAP t{};
    t.a = ap;
    t.K = K;
    return tc_dispatch(t, blob, ldw, K, M, ep, flags, tc_tile_n(ldw), st);
}

int tc_pack_weights(const float* wp, int ldw, int K, int bn, float* blob, cudaStream_t st) {
    int nkb = ceil_div(K, TC_BK), nt = ldw / bn;
    size_t total = (size_t)nt * nkb * bn * TC_BK;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    tc_pack_weights_kernel<<<blocks, 256, 0, st>>>(wp, ldw, K, bn, nkb, nt, blob);
    DEFT_CHECK_LAUNCH("tc_pack_weights");
    return DEFT_OK;
}

}  // namespace deft
