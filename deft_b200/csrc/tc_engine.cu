// tcgen05 (5th-gen tensor core) implicit-GEMM engine for sm_100a.
//
//   D[128 px, BN couts] (fp32, TMEM)  +=  A[128 x 32] (tf32)  x  B[BN x 32] (smem, K-major, SWIZZLE_128B)
//
// Two persistent kernels (grid <= 148, one CTA per SM, tiles strided over CTAs) share the PTX wrappers, the weight
// packer and the dispatch in this file:
//   tc_ts_kernel (tc_ts.cuh)   N tile <= 128: A goes through TMEM (TS-mode UMMA).  16 producer warps split every
//                              128x32 A block into tf32 hi + lo and tcgen05.st it; the operand source is either a
//                              TMA-loaded 18x10 halo patch in smem (3x3/s1 convs: TsPatchAP) or coalesced gathers
//                              transposed through a per-warp smem stage (1x1 / strided convs, the DCNv2 bilinear
//                              gather, the affinity pair expansion).  Two MMA issuer warps with private accumulators,
//                              4 epilogue warps (coalesced stores through an smem transpose), 1 weight-copy warp,
//                              1 patch-TMA warp.
//   tc_igemm_kernel (below)    both operands in smem (SS-mode).  Kept for the fused CenterNet heads (N tile = 256:
//                              one head's hidden layer never leaves the SM) and as the DEFT_TC_TS=0 reference path.
// * A is never materialised in HBM.
// * B (weights) is pre-packed by tc_pack_weights_kernel as ready-made swizzled smem tile images (hi and lo), streamed
//   in by one elected thread with cp.async.bulk (TMA bulk copy, mbarrier complete_tx).
// * fp32 parity: 3xTF32 error compensation -- x = hi + lo (hi = rna_tf32(x)), D += lo*hi + hi*lo + hi*hi, three
//   kind::tf32 MMAs per k-step into the same TMEM accumulator.  DEFT_FLAG_TF32X1 issues only hi*hi.
// * Small-M layers (DLA levels 4/5 at 1/16, 1/32 resolution) are split along K; partials go to caller scratch and
//   a deterministic reduce kernel applies the epilogue (no atomics).  A stream-K variant was measured and dropped
//   (profiles/r1_streamk_experiment.md).
// * Every mbarrier wait is bounded (mbar_wait): a protocol bug reports its site and traps, it never hangs the GPU.
#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "tc_engine.cuh"
#include <stdlib.h>

namespace deft {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // one 128-byte swizzle row of fp32
constexpr int TC_PWARPS = 16;                   // A-producer warps
constexpr int TC_EWARPS = 4;                    // epilogue warps (one per TMEM lane quadrant)
constexpr int TC_PRODUCERS = TC_PWARPS * 32;
constexpr int TC_MMA_WARP = TC_PWARPS + TC_EWARPS;
constexpr int TC_MMA_WARPS = 2;                 // the TS kernel alternates k-blocks between two issuer warps
constexpr int TC_AUX_WARP = TC_MMA_WARP + TC_MMA_WARPS;    // TMEM allocator + weight-tile bulk copies
constexpr int TC_THREADS = (TC_AUX_WARP + 2) * 32;   // + the TS kernel's halo-patch TMA loader warp
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
// Bounded wait: a protocol bug must trap, never hang the GPU.  After ~0.5 s every stuck warp reports where it waits
// (site / aux are caller-defined: which barrier of the protocol, which k-block or tile), after ~1.5 s the kernel traps.
#ifndef DEFT_TS_PROGRESS
#define DEFT_TS_PROGRESS 1    // 0 = no progress-board store before a wait (A/B builds: what the board costs)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int site = 0, int aux = 0, volatile int* prog = nullptr, int slot = 0) {
    // prog (TS kernel): a per-CTA progress board in shared memory -- every role's leader posts (site << 24 | counter) before it
    // waits, so that a stall report shows where ALL roles of the CTA stand, not only the warp that timed out
#if DEFT_TS_PROGRESS
    if (prog != nullptr && (threadIdx.x & 31) == 0) prog[slot] = (site << 24) | (aux & 0xffffff);
#endif
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    bool reported = false;
    while (!mbar_try_wait(bar, parity)) {
        const long long dt = clock64() - t0;
        if (dt > 1000000000LL && !reported) {
            reported = true;
            if ((threadIdx.x & 31) == 0) {
                printf("deft tc_engine: mbarrier stall block %d warp %d site %d aux %d bar %u parity %u\n", blockIdx.x, threadIdx.x >> 5, site, aux, bar, parity);
                if (prog != nullptr)
                    printf("deft tc_engine: block %d roles (site<<24|count): prod %x %x %x %x | issuer %x %x | bcopy %x | epi %x | tma %x\n", blockIdx.x,
                           prog[0], prog[1], prog[2], prog[3], prog[4], prog[5], prog[6], prog[7], prog[8]);
            }
        }
        if (dt > 3000000000LL) __trap();
    }
}
// Two barriers at once (the TS issuers: A stage and weight stage of a k-block).  Both try_waits are issued back to back so
// that their latencies overlap; waited one after the other they cost the issuer ~170 + ~400 cycles per k-block even when both
// phases had completed long before (r2_trace_conv32: 1715 cycles per own k-block of which 542 issue MMAs).
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b, int site_a, int site_b, int aux,
                                           volatile int* prog, int slot) {
#if DEFT_TS_PROGRESS
    if (prog != nullptr && (threadIdx.x & 31) == 0) prog[slot] = (site_a << 24) | (aux & 0xffffff);
#endif
    uint32_t oa = mbar_try_wait(bar_a, par_a), ob = mbar_try_wait(bar_b, par_b);
    if (oa & ob) return;
    long long t0 = clock64();
    bool reported = false;
    while (true) {
        if (!oa) oa = mbar_try_wait(bar_a, par_a);
        if (!ob) ob = mbar_try_wait(bar_b, par_b);
        if (oa & ob) return;
        const long long dt = clock64() - t0;
        if (dt > 1000000000LL && !reported) {
            reported = true;
            if ((threadIdx.x & 31) == 0) {
                printf("deft tc_engine: mbarrier stall block %d warp %d site %d aux %d bar %u parity %u\n", blockIdx.x, threadIdx.x >> 5, oa ? site_b : site_a, aux,
                       oa ? bar_b : bar_a, oa ? par_b : par_a);
                if (prog != nullptr)
                    printf("deft tc_engine: block %d roles (site<<24|count): prod %x %x %x %x | issuer %x %x | bcopy %x | epi %x | tma %x\n", blockIdx.x,
                           prog[0], prog[1], prog[2], prog[3], prog[4], prog[5], prog[6], prog[7], prog[8]);
            }
        }
        if (dt > 3000000000LL) __trap();
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
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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

// Warp-convergent leader election (CUTLASS elect_one_sync): tcgen05.mma / commit / bulk copies issued under this
// predicate compile to plain uniform-datapath instructions; issued from a divergent `lane == 0` branch instead, every
// UTCHMMA gets wrapped in an ELECT + BRA.U.ANY serialisation loop (~90 cycles per MMA, measured).
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred)::"memory");
    return pred;
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
// Thread t (0..511) owns 16-byte chunk c = t & 7 of rows r_i = (t >> 3) + 64 i, i = 0..1, of every k-block.
// The mainloop calls issue() once per k-block in ascending order (global loads into registers, DEPTH k-blocks in
// flight) and combine() once per k-block in the same order, so producers keep incremental (tap, channel-block)
// counters instead of dividing -- the producer path is instruction-issue bound, every instruction counts.
constexpr int TC_ROWS = 2;

// Fast path: one NHWC source, channel count a multiple of 32, K order = (32-channel block, tap, channel)
// (tc_k_perm_taps) or a 1x1 conv / plain GEMM.  Per k-block: one LDS (tap offset), TC_ROWS predicated LDG.128.
struct TcConvFastAP {
    const float* src;
    int pitch, H, W, OH, OW, KW, stride, pad, taps;
    static constexpr int kExtraSmemBytes = 64;          // tap offset table (<= 9 taps used; 7x7 never comes here)
    static constexpr int RAW = TC_ROWS;
    static constexpr int DEPTH = 3;
    const float* rowp_[TC_ROWS];
    unsigned vmask_[TC_ROWS];
    const int* toff_;
    int tap_, cboff_;
    __device__ __forceinline__ void setup_once(unsigned char* extra) {       // tile-independent: tap offset table
        int* toff = reinterpret_cast<int*>(extra);
        if (threadIdx.x < taps) {
            int ky = threadIdx.x / KW, kx = threadIdx.x - ky * KW;
            toff[threadIdx.x] = ((ky - pad) * W + (kx - pad)) * pitch;
        }
        toff_ = toff;
        asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    __device__ __forceinline__ void init(int m0, int M, unsigned char*, int kb0) {
        const int KH = taps / KW;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            const int p = m0 + (threadIdx.x >> 3) + 64 * i;
            vmask_[i] = 0;
            rowp_[i] = src;
            if (p < M) {
                const int n = p / (OH * OW);
                const int r = p - n * (OH * OW);
                const int oy = (r / OW) * stride, ox = (r - (r / OW) * OW) * stride;
                rowp_[i] = src + (((size_t)n * H + oy) * W + ox) * pitch + (threadIdx.x & 7) * 4;
                unsigned bit = 1u;
                for (int ky = 0; ky < KH; ++ky)
                    for (int kx = 0; kx < KW; ++kx, bit <<= 1) {
                        const int iy = oy - pad + ky, ix = ox - pad + kx;
                        if (iy >= 0 && iy < H && ix >= 0 && ix < W) vmask_[i] |= bit;
                    }
            }
        }
        const int cb = kb0 / taps;
        tap_ = kb0 - cb * taps;
        cboff_ = cb * TC_BK;
    }
    __device__ __forceinline__ void issue(float4 (&raw)[RAW]) {
        const int off = toff_[tap_] + cboff_;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            raw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((vmask_[i] >> tap_) & 1u) raw[i] = ldg4(rowp_[i] + off);
        }
        if (++tap_ == taps) { tap_ = 0; cboff_ += TC_BK; }
    }
    __device__ __forceinline__ void combine(const float4 (&raw)[RAW], float4 (&v)[TC_ROWS]) {
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) v[i] = raw[i];
    }
};

// 3xFP16 variant of the fast path (NPASS = 6): k-block = 64 channels, K order (64-channel block, tap, channel); a thread's
// 16-byte chunk of an fp16 A row is 8 channels = two LDG.128 of fp32 per row.
struct TcConvFast16AP {
    const float* src;
    int pitch, H, W, OH, OW, KW, stride, pad, taps;
    static constexpr int kExtraSmemBytes = 64;
    static constexpr int RAW = TC_ROWS * 2;
    static constexpr int DEPTH = 2;
    const float* rowp_[TC_ROWS];
    unsigned vmask_[TC_ROWS];
    const int* toff_;
    int tap_, cboff_;
    __device__ __forceinline__ void setup_once(unsigned char* extra) {
        int* toff = reinterpret_cast<int*>(extra);
        if (threadIdx.x < taps) {
            int ky = threadIdx.x / KW, kx = threadIdx.x - ky * KW;
            toff[threadIdx.x] = ((ky - pad) * W + (kx - pad)) * pitch;
        }
        toff_ = toff;
        asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    __device__ __forceinline__ void init(int m0, int M, unsigned char*, int kb0) {
        const int KH = taps / KW;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            const int p = m0 + (threadIdx.x >> 3) + 64 * i;
            vmask_[i] = 0;
            rowp_[i] = src;
            if (p < M) {
                const int n = p / (OH * OW);
                const int r = p - n * (OH * OW);
                const int oy = (r / OW) * stride, ox = (r - (r / OW) * OW) * stride;
                rowp_[i] = src + (((size_t)n * H + oy) * W + ox) * pitch + (threadIdx.x & 7) * 8;
                unsigned bit = 1u;
                for (int ky = 0; ky < KH; ++ky)
                    for (int kx = 0; kx < KW; ++kx, bit <<= 1) {
                        const int iy = oy - pad + ky, ix = ox - pad + kx;
                        if (iy >= 0 && iy < H && ix >= 0 && ix < W) vmask_[i] |= bit;
                    }
            }
        }
        const int cb = kb0 / taps;
        tap_ = kb0 - cb * taps;
        cboff_ = cb * 64;
    }
    __device__ __forceinline__ void issue(float4 (&raw)[RAW]) {
        const int off = toff_[tap_] + cboff_;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            raw[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
            raw[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((vmask_[i] >> tap_) & 1u) { raw[2 * i] = ldg4(rowp_[i] + off); raw[2 * i + 1] = ldg4(rowp_[i] + off + 4); }
        }
        if (++tap_ == taps) { tap_ = 0; cboff_ += 64; }
    }
    // v: the row's 8 channels as two float4 per row (only used through the f16 store path of the kernel)
    __device__ __forceinline__ void combine(const float4 (&raw)[RAW], float4 (&v)[RAW]) {
#pragma unroll
        for (int i = 0; i < RAW; ++i) v[i] = raw[i];
    }
};

// General path: channel concat of up to 4 sources and/or channel counts that are not multiples of 32
// (Root 1x1 convs, the 16-channel full-resolution layers).  K order as tc_k_perm_taps says.
struct TcConvAP {
    ConvAP a;
    int K;
    int perm_taps;       // 0 = tap-major K order, else KH*KW with channel-block-major order (tc_k_perm_taps)
    static constexpr int kExtraSmemBytes = 0;
    static constexpr int RAW = TC_ROWS;
    static constexpr int DEPTH = 2;
    int n_[TC_ROWS], oy_[TC_ROWS], ox_[TC_ROWS];
    unsigned valid_;
    int kb_;
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void init(int m0, int M, unsigned char*, int kb0) {
        valid_ = 0;
        kb_ = kb0;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            int p = m0 + (threadIdx.x >> 3) + 64 * i;
            bool v = p < M;
            if (v) valid_ |= 1u << i;
            int q = v ? p : 0;
            n_[i] = q / (a.OH * a.OW);
            int r = q - n_[i] * (a.OH * a.OW);
            oy_[i] = (r / a.OW) * a.stride - a.pad;
            ox_[i] = (r - (r / a.OW) * a.OW) * a.stride - a.pad;
        }
    }
    __device__ __forceinline__ void issue(float4 (&raw)[RAW]) {
        const int kb = kb_++;
        int tap, cc;
        bool kok = true;
        if (perm_taps) {
            const int cb = kb / perm_taps;
            tap = kb - cb * perm_taps;
            cc = cb * TC_BK + (threadIdx.x & 7) * 4;
        } else {
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
        for (int i = 0; i < TC_ROWS; ++i) {
            int iy = oy_[i] + ky;
            int ix = ox_[i] + kx;
            raw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kok && ((valid_ >> i) & 1u) && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                raw[i] = ldg4(base + (((size_t)n_[i] * a.H + iy) * a.W + ix) * pitch + cc);
        }
    }
    __device__ __forceinline__ void combine(const float4 (&raw)[RAW], float4 (&v)[TC_ROWS]) {
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) v[i] = raw[i];
    }
};

// DCNv2: A[p, k] = mask_tap(p) * bilinear(x[:, c]; p + tap + offset_tap(p)); K order = (32-channel block, tap, channel)
// so that the 9 taps of a channel block re-read the same neighbourhood out of L1.  The 4 corner offsets/weights of
// the tile's 128 pixels x 9 taps are computed once into smem.
struct TcDcnAP {
    DcnAP a;
    int K;
    static constexpr int kExtraSmemBytes = TC_BM * 9 * 32;
    static constexpr int RAW = TC_ROWS * 4;
    static constexpr int DEPTH = 1;      // 8 float4 in flight per thread; a second set would spill under the 80-register cap
    int4* tabO_;
    float4* tabW_;
    const float* xc_;
    int tap_i_, cboff_, tap_c_;
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void init(int m0, int M, unsigned char* extra, int kb0) {
        tabO_ = reinterpret_cast<int4*>(extra);
        tabW_ = reinterpret_cast<float4*>(extra + TC_BM * 9 * 16);
        const int H = a.H, W = a.W;
        asm volatile("bar.sync 1, 512;" ::: "memory");     // every producer is done reading the previous tile's table
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
                    if (y0ok && x0ok) { o.x = (base + y0 * W + x0) * a.pitch; w.x = hy * hx * mk; }
                    if (y0ok && x1ok) { o.y = (base + y0 * W + x0 + 1) * a.pitch; w.y = hy * lx * mk; }
                    if (y1ok && x0ok) { o.z = (base + (y0 + 1) * W + x0) * a.pitch; w.z = ly * hx * mk; }
                    if (y1ok && x1ok) { o.w = (base + (y0 + 1) * W + x0 + 1) * a.pitch; w.w = ly * lx * mk; }
                }
            }
            tabO_[idx] = o;       // element offsets; corners with zero weight point at element 0: always a valid address
            tabW_[idx] = w;
        }
        const int cb = kb0 / 9;
        tap_i_ = tap_c_ = kb0 - cb * 9;
        cboff_ = cb * TC_BK;
        xc_ = a.x + (threadIdx.x & 7) * 4;
        asm volatile("bar.sync 1, 512;" ::: "memory");     // producers only: the table is private to them
    }
    __device__ __forceinline__ void issue(float4 (&raw)[RAW]) {
        const float* x = xc_ + cboff_;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            const int4 o = tabO_[((threadIdx.x >> 3) + 64 * i) * 9 + tap_i_];
            raw[i * 4 + 0] = ldg4(x + o.x);
            raw[i * 4 + 1] = ldg4(x + o.y);
            raw[i * 4 + 2] = ldg4(x + o.z);
            raw[i * 4 + 3] = ldg4(x + o.w);
        }
        if (++tap_i_ == 9) { tap_i_ = 0; cboff_ += TC_BK; }
    }
    __device__ __forceinline__ void combine(const float4 (&raw)[RAW], float4 (&v)[TC_ROWS]) {
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            const float4 w = tabW_[((threadIdx.x >> 3) + 64 * i) * 9 + tap_c_];
            float4 acc;
            acc.x = w.x * raw[i * 4].x; acc.y = w.x * raw[i * 4].y; acc.z = w.x * raw[i * 4].z; acc.w = w.x * raw[i * 4].w;
            acc.x += w.y * raw[i * 4 + 1].x; acc.y += w.y * raw[i * 4 + 1].y; acc.z += w.y * raw[i * 4 + 1].z; acc.w += w.y * raw[i * 4 + 1].w;
            acc.x += w.z * raw[i * 4 + 2].x; acc.y += w.z * raw[i * 4 + 2].y; acc.z += w.z * raw[i * 4 + 2].z; acc.w += w.z * raw[i * 4 + 2].w;
            acc.x += w.w * raw[i * 4 + 3].x; acc.y += w.w * raw[i * 4 + 3].y; acc.z += w.w * raw[i * 4 + 3].z; acc.w += w.w * raw[i * 4 + 3].w;
            v[i] = acc;
        }
        if (++tap_c_ == 9) tap_c_ = 0;
    }
};

// Affinity MLP layer 2: A[p, k] = relu(Au'[f(p), i(p), k] + Bv'[j(p), k]) over the live pairs.
struct TcPairAP {
    PairAP a;
    int K;
    static constexpr int kExtraSmemBytes = 0;
    static constexpr int RAW = TC_ROWS * 2;
    static constexpr int DEPTH = 2;
    const float* pa_[TC_ROWS];
    const float* pb_[TC_ROWS];
    int k_;
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void init(int m0, int M, unsigned char*, int kb0) {
        k_ = kb0 * TC_BK + (threadIdx.x & 7) * 4;
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            int p = m0 + (threadIdx.x >> 3) + 64 * i;
            pa_[i] = nullptr;
            pb_[i] = nullptr;
            if (p < M) {
                int f = 0;
                while (f < a.F - 1 && p >= a.off[f + 1]) ++f;
                int q = p - a.off[f];
                int ii = q / a.m, jj = q - ii * a.m;
                pa_[i] = a.au + ((size_t)f * a.P + ii) * a.ld;
                pb_[i] = a.bv + (size_t)f * a.bvf + (size_t)jj * a.ld;
            }
        }
    }
    __device__ __forceinline__ void issue(float4 (&raw)[RAW]) {
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            raw[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
            raw[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pa_[i] != nullptr && k_ < K) {
                raw[2 * i] = ldg4(pa_[i] + k_);
                raw[2 * i + 1] = ldg4(pb_[i] + k_);
            }
        }
        k_ += TC_BK;
    }
    __device__ __forceinline__ void combine(const float4 (&raw)[RAW], float4 (&v)[TC_ROWS]) {
#pragma unroll
        for (int i = 0; i < TC_ROWS; ++i) {
            v[i].x = fmaxf(raw[2 * i].x + raw[2 * i + 1].x, 0.f); v[i].y = fmaxf(raw[2 * i].y + raw[2 * i + 1].y, 0.f);
            v[i].z = fmaxf(raw[2 * i].z + raw[2 * i + 1].z, 0.f); v[i].w = fmaxf(raw[2 * i].w + raw[2 * i + 1].w, 0.f);
        }
    }
};

// ------------------------------------------------------------------------------------------------ epilogues
// 16 epilogue warps: TMEM lane quadrant = warp & 3, column group = warp >> 2.  A warp reads LW columns of its 32
// rows per tcgen05.ld (one wait per LW columns, not per 8).
// tmem_ld_issue: the load only; the caller batches several and pays ONE tcgen05.wait::ld (tmem_ld_wait) for all of them.
template <int LW>
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[LW]) {
    static_assert(LW == 8 || LW == 16 || LW == 32, "tmem_ld width");
    if constexpr (LW == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr) : "memory");
    } else if constexpr (LW == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
                     "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
                       "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
                       "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int LW>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[LW]) {
    uint32_t r[LW];
    tmem_ld_issue<LW>(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < LW; ++i) v[i] = __uint_as_float(r[i]);
}

// Epilogue functors, run by the 4 dedicated epilogue warps (TMEM lane quadrant = warp & 3; a thread owns one
// accumulator row): run<BN>(tmem_acc, smem, ewarp, lane, m0, M, ntile, z).
struct EpiStd {
    Epilogue e;
    float* partial;      // split-K scratch [splits][M][ldp] or nullptr
    int ldp;
    static constexpr int kSmemBytes = 0;
    template <int BN>
    __device__ __forceinline__ void run(uint32_t tmem_acc, unsigned char*, int ewarp, int lane, int m0, int M, int ntile, int z,
                                        uint32_t acc2_cols = 0u) const {
        const int p = m0 + ewarp * 32 + lane;
        run_p<BN>(tmem_acc, ewarp, p, p < M, M, ntile, z, acc2_cols);
    }
    // this thread's accumulator row belongs to output pixel p (ignored when !valid)
    template <int BN>
    __device__ __forceinline__ void run_p(uint32_t tmem_acc, int ewarp, int p, bool valid, int M, int ntile, int z, uint32_t acc2_cols) const {
        // acc2_cols != 0: a second partial accumulator sits acc2_cols columns further (TS kernel, one per MMA issuer)
        constexpr int LW = BN >= 32 ? 32 : BN;                // columns per tcgen05.ld
        const int lane_base = ewarp * 32;
#pragma unroll 1
        for (int c = 0; c < BN; c += LW) {
            float v[LW];
            tmem_ld<LW>(tmem_acc + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
            if (acc2_cols) {
                float v2[LW];
                tmem_ld<LW>(tmem_acc + acc2_cols + ((uint32_t)lane_base << 16) + (uint32_t)c, v2);
#pragma unroll
                for (int j = 0; j < LW; ++j) v[j] += v2[j];
            }
            const int co = ntile * BN + c;
            if (!valid) continue;
            if (partial != nullptr) {
                float* dst = partial + ((size_t)z * M + p) * ldp + co;
#pragma unroll
                for (int q = 0; q < LW / 4; ++q) *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                continue;
            }
#pragma unroll
            for (int q = 0; q < LW / 4; ++q) {
                const int cq = co + q * 4;
                if (cq < e.Cout) {
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
    static constexpr int kSmemBytes = HEADS_MAXC * 256 * 4 + 256 * 4 + 16;
    template <int BN>
    __device__ __forceinline__ void run(uint32_t tmem_acc, unsigned char* smem, int ewarp, int lane, int m0, int M, int ntile, int) const {
        static_assert(BN == 256, "EpiHeads needs one 256-wide tile per head");
        float* w2s = reinterpret_cast<float*>(smem);
        float* b1s = w2s + HEADS_MAXC * 256;
        int* cur_head = reinterpret_cast<int*>(b1s + 256);
        const int h = ntile;
        const int ch = c[h];
        const int etid = ewarp * 32 + lane;
        // (re)stage this head's 1x1 weights; named barrier 2 = the 128 epilogue threads
        asm volatile("bar.sync 2, 128;" ::: "memory");        // previous tile's readers are done
        for (int i = etid; i < ch * 256; i += TC_EWARPS * 32) w2s[i] = w2[(size_t)row0[h] * 256 + i];
        for (int i = etid; i < 256; i += TC_EWARPS * 32) b1s[i] = b1[h * 256 + i];
        asm volatile("bar.sync 2, 128;" ::: "memory");
        (void)cur_head;
        const int lane_base = ewarp * 32;
        const int p = m0 + lane_base + lane;
        float acc[HEADS_MAXC];
#pragma unroll
        for (int j = 0; j < HEADS_MAXC; ++j) acc[j] = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < 256; cb += 16) {
            float v[16];
            tmem_ld<16>(tmem_acc + ((uint32_t)lane_base << 16) + (uint32_t)cb, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + b1s[cb + j], 0.f);
#pragma unroll
            for (int o = 0; o < HEADS_MAXC; ++o) {
                if (o < ch) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(w2s + o * 256 + cb + q * 4);
                        acc[o] += v[q * 4] * w.x + v[q * 4 + 1] * w.y + v[q * 4 + 2] * w.z + v[q * 4 + 3] * w.w;
                    }
                }
            }
        }
        if (p < M) {
            const int n = p / HW, q = p - n * HW;
#pragma unroll
            for (int o = 0; o < HEADS_MAXC; ++o)
                if (o < ch) out[h][((size_t)n * ch + o) * HW + q] = acc[o] + b2[row0[h] + o];
        }
    }
};

// ------------------------------------------------------------------------------------------------ kernel
// Persistent: gridDim.x CTAs (<= #SMs) walk the tile list t = blockIdx.x, +gridDim.x, ...; tile = (m_tile, n_tile,
// k-split).  The smem ring and its phases run straight through tile boundaries; the accumulator is double-buffered
// in TMEM so the epilogue warps drain tile i while the tensor core already works on tile i+1.
struct TileSched {
    int m_tiles, n_tiles, splits, total, nkb_total, kb_per_split;
};
template <class AP, int BN, int NPASS, class EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_igemm_kernel(AP ap, const float* __restrict__ wblob, TileSched ts, int M, int stages, EPI epi,
                                                                 int dbg, long long* trace) {
    // trace (DEFT_TC_TRACE, perf bring-up only): CTA 0 records clock64() at protocol points, [role][k-block < 96][4]
    DEFT_PDL_PROLOGUE();
#define TC_TRACE(role, idx, slot) do { if (trace != nullptr && blockIdx.x == 0 && (idx) < 96) trace[((role) * 96 + (idx)) * 4 + (slot)] = clock64(); } while (0)
    constexpr int B_TILE = BN * 128;                      // bytes of one B part (hi or lo)
    constexpr bool F16 = NPASS == 6;                      // 3xFP16: 64 channels per k-block in the same 128-byte rows (see TsCfg)
    constexpr int PASSES = F16 ? 3 : NPASS;
    constexpr int STAGE_BYTES = PASSES == 3 ? 2 * (TC_A_TILE + B_TILE) : (TC_A_TILE + B_TILE);
    constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulators
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
    auto acc_full = [&](int b) { return bar_base + 8u * (24 + b); };
    auto acc_empty = [&](int b) { return bar_base + 8u * (26 + b); };
    const uint32_t tmem_slot = bar_base + 8u * 28;
    unsigned char* ap_extra = smem + (size_t)stages * STAGE_BYTES + 256;
    unsigned char* epi_extra = ap_extra + AP::kExtraSmemBytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_a(s), TC_PWARPS);              // one arrival per producer warp
            mbar_init(full_b(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), TC_EWARPS);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == TC_AUX_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (size_t)stages * STAGE_BYTES + 8 * 28);
    const int tiles_mn = ts.m_tiles * ts.n_tiles;

    if (warp < TC_PWARPS) {
        // ===================== A producers =====================
        ap.setup_once(ap_extra);
        uint32_t soff[TC_ROWS];
#pragma unroll
        for (int r = 0; r < TC_ROWS; ++r) soff[r] = swz((tid >> 3) + 64 * r, tid & 7);
        int s = 0, gk = 0;
        uint32_t ph = 0;
        // dbg bits (DEFT_TC_DBG, perf bring-up only): 1 no B copy, 2 no A global loads, 4 no A smem stores/fence, 8 no MMA, 16 no epilogue
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int m0 = (mn / ts.n_tiles) * TC_BM;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            ap.init(m0, M, ap_extra, kb0);
            float4 raw[AP::DEPTH][AP::RAW];
#pragma unroll
            for (int d = 0; d < AP::DEPTH; ++d)
                if (d < nkb) ap.issue(raw[d]);
            for (int it = 0; it < nkb; it += AP::DEPTH) {
#pragma unroll
                for (int d = 0; d < AP::DEPTH; ++d) {
                    const int i = it + d;                     // local k-block index
                    if (i < nkb) {
                        float4 v[AP::RAW >= TC_ROWS && F16 ? AP::RAW : TC_ROWS];
                        if (tid == 0) TC_TRACE(1, gk, 0);
                        ap.combine(raw[d], v);
                        if (i + AP::DEPTH < nkb && !(dbg & 2)) ap.issue(raw[d]);
                        if (tid == 0) TC_TRACE(1, gk, 1);
                        mbar_wait(empty(s), ph ^ 1u);
                        if (tid == 0) TC_TRACE(1, gk, 2);
                        unsigned char* a_hi = smem + (size_t)s * STAGE_BYTES;
                        if constexpr (F16) {
#pragma unroll
                            for (int r = 0; r < TC_ROWS; ++r) {         // 8 channels -> 8 fp16 hi + 8 fp16 lo = one 16-byte chunk each
                                const float f[8] = {v[2 * r].x, v[2 * r].y, v[2 * r].z, v[2 * r].w, v[2 * r + 1].x, v[2 * r + 1].y, v[2 * r + 1].z, v[2 * r + 1].w};
                                uint32_t hw[4], lw[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const __half2 hh = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                                    const float2 bk = __half22float2(hh);
                                    const __half2 ll = __floats2half2_rn(f[2 * j] - bk.x, f[2 * j + 1] - bk.y);
                                    hw[j] = *reinterpret_cast<const uint32_t*>(&hh);
                                    lw[j] = *reinterpret_cast<const uint32_t*>(&ll);
                                }
                                *reinterpret_cast<uint4*>(a_hi + soff[r]) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                *reinterpret_cast<uint4*>(a_hi + TC_A_TILE + soff[r]) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            }
                        } else
#pragma unroll
                        for (int r = 0; r < TC_ROWS; ++r) {
                            if (dbg & 4) break;
                            float4 h;
                            h.x = tf32_rna(v[r].x); h.y = tf32_rna(v[r].y); h.z = tf32_rna(v[r].z); h.w = tf32_rna(v[r].w);
                            *reinterpret_cast<float4*>(a_hi + soff[r]) = h;
                            if (NPASS == 3) {
                                float4 l;
                                l.x = v[r].x - h.x; l.y = v[r].y - h.y; l.z = v[r].z - h.z; l.w = v[r].w - h.w;
                                *reinterpret_cast<float4*>(a_hi + TC_A_TILE + soff[r]) = l;
                            }
                        }
                        if (!(dbg & 4)) fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_a(s)); // one arrival per warp: 16 per stage, not 512
                        if (tid == 0) TC_TRACE(1, gk, 3);
                        ++gk;
                        if (++s == stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp < TC_MMA_WARP) {
        // ===================== epilogue warps =====================
        const int ewarp = warp - TC_PWARPS;
        int iter = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++iter) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int m0 = (mn / ts.n_tiles) * TC_BM, ntile = mn % ts.n_tiles;
            const int b = iter & 1;
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 0);
            mbar_wait(acc_full(b), (uint32_t)(iter >> 1) & 1u);
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 1);
            tc_fence_after();
            if (!(dbg & 16)) epi.template run<BN>(tmem_base + (uint32_t)(b * BN), epi_extra, ewarp, lane, m0, M, ntile, z);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(b));
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 2);
        }
    } else if (warp == TC_MMA_WARP) {
        // ===================== MMA issuer (whole warp convergent, one elected lane issues) =====================
        {
            constexpr uint32_t idesc = F16 ? ((1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24)) : make_idesc(BN);
            int s = 0, iter = 0, gk = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++iter) {
                const int z = t / tiles_mn;
                const int kb0 = z * ts.kb_per_split;
                const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
                const int b = iter & 1;
                mbar_wait(acc_empty(b), ((uint32_t)(iter >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(b * BN);
                for (int i = 0; i < nkb; ++i, ++gk) {
                    if (lane == 0) TC_TRACE(0, gk, 0);
                    mbar_wait2(full_a(s), ph, full_b(s), ph, 5, 6, gk, nullptr, 0);     // both operands of the stage, one round trip
                    if (lane == 0) { TC_TRACE(0, gk, 1); TC_TRACE(0, gk, 2); }
                    tc_fence_after();
                    const uint32_t a_hi = sbase + (uint32_t)s * STAGE_BYTES;
                    const uint32_t a_lo = a_hi + TC_A_TILE;
                    const uint32_t b_hi = a_hi + (PASSES == 3 ? 2 * TC_A_TILE : TC_A_TILE);
                    const uint32_t b_lo = b_hi + B_TILE;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) {
                            if (dbg & 8) break;
                            const uint64_t da_hi = make_smem_desc(a_hi + k * 32), db_hi = make_smem_desc(b_hi + k * 32);
                            if constexpr (F16) {
                                const uint64_t da_lo = make_smem_desc(a_lo + k * 32), db_lo = make_smem_desc(b_lo + k * 32);
                                umma_f16(tmem_acc, da_lo, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                                umma_f16(tmem_acc, da_hi, db_lo, idesc, 1u);
                                umma_f16(tmem_acc, da_hi, db_hi, idesc, 1u);
                            } else if (NPASS == 3) {
                                // small terms first, so the big hi*hi product is added to an already-formed correction
                                const uint64_t da_lo = make_smem_desc(a_lo + k * 32), db_lo = make_smem_desc(b_lo + k * 32);
                                umma_tf32(tmem_acc, da_lo, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                                umma_tf32(tmem_acc, da_hi, db_lo, idesc, 1u);
                                umma_tf32(tmem_acc, da_hi, db_hi, idesc, 1u);
                            } else {
                                umma_tf32(tmem_acc, da_hi, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        umma_commit(empty(s));            // slot free once these MMAs have read it
                        if (i == nkb - 1) umma_commit(acc_full(b));   // accumulator complete
                    }
                    __syncwarp();
                    if (lane == 0) TC_TRACE(0, gk, 3);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == TC_AUX_WARP) {
        // ===================== weight-tile bulk copies (whole warp convergent, one elected lane issues) =====================
        {
            constexpr uint32_t bytes = PASSES == 3 ? 2 * B_TILE : B_TILE;
            int s = 0, gk = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
                const int z = t / tiles_mn, mn = t - z * tiles_mn;
                const int ntile = mn % ts.n_tiles;
                const int kb0 = z * ts.kb_per_split;
                const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
                const unsigned char* src = reinterpret_cast<const unsigned char*>(wblob) + ((size_t)ntile * ts.nkb_total + kb0) * (2 * B_TILE);
                for (int i = 0; i < nkb; ++i, ++gk) {
                    if (lane == 0) TC_TRACE(2, gk, 0);
                    mbar_wait(empty(s), ph ^ 1u);
                    if (lane == 0) TC_TRACE(2, gk, 1);
                    const uint32_t dst = sbase + (uint32_t)s * STAGE_BYTES + (PASSES == 3 ? 2 * TC_A_TILE : TC_A_TILE);
                    if (elect_one_sync()) {
                        if (dbg & 1) { mbar_arrive(full_b(s)); }
                        else {
                            mbar_arrive_expect_tx(full_b(s), bytes);
                            bulk_g2s(dst, src + (size_t)i * (2 * B_TILE), bytes, full_b(s));
                        }
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    }
    __syncthreads();
    if (warp == TC_AUX_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace deft
#include "tc_ts.cuh"
namespace deft {

// Deterministic split-K tail: out = epilogue(sum_z partial[z]).
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int ldp, Epilogue e) {
    DEFT_PDL_PROLOGUE();
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

// 3xFP16 weight images (TsCfg::F16): k-blocks of 64 channels, [n_tiles][nkb][hi|lo][BN rows x 128 B swizzled] of fp16,
// w = hi + lo + O(2^-22 w).  perm_taps > 0: K order (64-channel block, tap, channel).
__global__ void tc_pack_weights_f16_kernel(const float* __restrict__ wp, int ldw, int K, int BN, int nkb, int n_tiles, int perm_taps,
                                           __half* __restrict__ blob) {
    size_t total = (size_t)n_tiles * nkb * BN * 64;
    const int ctot = perm_taps ? K / perm_taps : 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int kk = (int)(i % 64);
        size_t t = i / 64;
        int n = (int)(t % BN);
        t /= BN;
        int kb = (int)(t % nkb);
        int nt = (int)(t / nkb);
        int k = kb * 64 + kk, co = nt * BN + n;
        if (perm_taps) {
            int cb = kb / perm_taps, tap = kb - cb * perm_taps;
            k = tap * ctot + cb * 64 + kk;
        }
        float w = (k < K && co < ldw) ? wp[(size_t)k * ldw + co] : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        size_t tile = ((size_t)nt * nkb + kb) * (size_t)(2 * BN * 64);       // halves per (hi + lo) tile pair
        size_t off = (size_t)n * 64 + (size_t)(((kk >> 3) ^ (n & 7)) << 3) + (kk & 7);
        blob[tile + off] = hi;
        blob[tile + (size_t)BN * 64 + off] = lo;
    }
}

// Which numeric mode a weight image is packed for -- a pure function of (K, taps, bn) so that the packer and every dispatcher
// agree without a flag travelling with the blob: 3xFP16 when enabled (DEFT_TC_F16) and the channel count per tap is a multiple of 64
// (TS-kernel tiles bn <= 128 and the fused heads' bn = 256); else 3xTF32.
bool tc_blob_is_f16(int K, int taps, int bn) {
    static int en = -1;
    if (en < 0) { const char* e = getenv("DEFT_TC_F16"); en = e ? atoi(e) : 1; }     // default ON (r2: 523 -> 569 fps at C2, parity 2.5e-4 -> 1.3e-4); 0 = 3xTF32 everywhere
    if (taps < 1) taps = 1;
    return en != 0 && (bn <= 128 || bn == 256) && K % taps == 0 && (K / taps) % 64 == 0;
}

// ------------------------------------------------------------------------------------------------ host
// Tile-N choice.  Depends only on (ldw, M, kind) so the packer and the dispatcher always agree; the chosen
// value also travels explicitly with every call (tc_bn).
// K permutation rule shared by the packer and every dispatcher: multi-tap convs whose channel count is a
// multiple of 32 run channel-block-major (returns the tap count), everything else tap-major (returns 0).
int tc_k_perm_taps(int K, int taps) { return (taps > 1 && K % taps == 0 && (K / taps) % TC_BK == 0) ? taps : 0; }

int tc_pick_bn(int ldw, int M, int kind) {
    if (kind == TC_KIND_HEADS) return 256;
    // DCN: the bilinear gather is the expensive side, so every gathered A tile should meet all of Cout (up to 128 columns;
    // r1 capped this at 64 and repeated the gather 2x / 4x for the Cout = 128 / 256 layers).  DEFT_DCN_MAXBN overrides.
    static int dcn_max = -1;
    if (dcn_max < 0) { const char* e = getenv("DEFT_DCN_MAXBN"); dcn_max = e ? atoi(e) : 128; if (dcn_max != 32 && dcn_max != 64) dcn_max = 128; }
    const int max_bn = (kind == TC_KIND_DCN) ? dcn_max : 128;
    const int mt = ceil_div(M, TC_BM);
    int best = 0;
    // DEFT_TC_WIDE=1 (default): always the widest dividing tile -- with several frames in flight (FramePipeline) idle SMs
    // are filled by the other frames' kernels, so fewer, longer CTAs that produce each A tile once win (+2.6 % measured);
    // 0 = prefer tiles narrow enough to give one launch >= 120 CTAs (better for a single frame in isolation).
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("DEFT_TC_WIDE"); wide = e ? atoi(e) : 1; }
    for (int bn = max_bn; bn >= 32; bn >>= 1) {
        if (ldw % bn != 0) continue;
        if (best == 0) best = bn;                        // largest dividing tile
        if (wide) return bn;
        if (mt * (ldw / bn) >= 120) return bn;           // enough CTAs for one wave
        if (bn == 64) break;                             // below 64 the redundant A gathers cost more than split-K
    }
    if (best == 0) best = (ldw % 32 == 0) ? 32 : 16;
    return best;                                         // too few CTAs either way: keep the widest tile, split K instead
}
static int tc_pick_splits(int ctas, int nkb, bool dcn = false) {
    // CTAs a split-K layer should reach (DEFT_TC_SPLIT_TARGET).  A third of the SMs by default: with several frames in flight
    // the rest is busy with the neighbouring frames' kernels, and every split costs a partial write + read.  Measured on one
    // box with the final r2 build, 4 frames in flight (tools/gpu_call25.sh): 1 -> 616.4, 24 -> 621.4, 37 -> 621.9, 50 -> 624.4,
    // 60 -> 623.4, 74 -> 617.6 / 619.0 fps (r1: 148 -> 487, 74 -> 497, 37 -> 486); one isolated frame prefers more splits
    // (depth 1: 74 -> 411.7, 50 -> 399.9 fps).
    // The DCN gather-GEMMs keep half the SMs (74): the metric also quotes their time alone (bench.py roofline: one eager
    // frame), where more splits are what the four few-tile DCN layers want; the convs take the throughput setting.
    static int target_conv = -1, target_dcn = -1;
    if (target_conv < 0) {
        const char* e = getenv("DEFT_TC_SPLIT_TARGET");
        target_conv = e ? atoi(e) : 50;
        target_dcn = e ? atoi(e) : TC_NUM_SMS / 2;
    }
    const int target = dcn ? target_dcn : target_conv;
    if (ctas >= 100 || ctas >= target) return 1;
    int s = target / ctas;
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
    const int splits = tc_pick_splits(ceil_div(M, TC_BM) * (ldw / bn), ceil_div(K, TC_BK), kind == TC_KIND_DCN);
    return splits > 1 ? align_up((size_t)splits * M * ldw * sizeof(float), 256) : 0;
}

template <class AP, int BN, int NPASS, class EPI>
static int tc_launch(AP ap, const float* blob, int nkb, int M, int n_tiles, int splits, EPI epi, cudaStream_t st) {
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = NPASS != 1 ? 2 * (TC_A_TILE + B_TILE) : (TC_A_TILE + B_TILE);
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + AP::kExtraSmemBytes + EPI::kSmemBytes;
    TileSched ts{};
    ts.kb_per_split = ceil_div(nkb, splits);
    ts.splits = ceil_div(nkb, ts.kb_per_split);          // no empty split
    ts.m_tiles = ceil_div(M, TC_BM);
    ts.n_tiles = n_tiles;
    ts.nkb_total = nkb;
    ts.total = ts.m_tiles * ts.n_tiles * ts.splits;
    int stages = (TC_SMEM_LIMIT - fixed) / STAGE_BYTES;
    if (stages > 6) stages = 6;
    if (stages < 2) return DEFT_EINVAL;
    size_t smem = (size_t)stages * STAGE_BYTES + fixed;
    static size_t granted[DEFT_MAX_DEVICES] = {};
    {
        const int rc_attr = ensure_dyn_smem(tc_igemm_kernel<AP, BN, NPASS, EPI>, TC_SMEM_LIMIT, granted, "tc_igemm_kernel");
        if (rc_attr != DEFT_OK) return rc_attr;
    }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DEFT_TC_DBG"); dbg = e ? atoi(e) : 0; }
    static int stage_cap = -1;
    if (stage_cap < 0) { const char* e = getenv("DEFT_TC_STAGES"); stage_cap = e ? atoi(e) : 0; }
    if (stage_cap >= 2 && stage_cap < stages) stages = stage_cap;
    const int grid = ts.total < TC_NUM_SMS ? ts.total : TC_NUM_SMS;
    long long* trace = debug_trace_ptr();
    launch_pdl(tc_igemm_kernel<AP, BN, NPASS, EPI>, dim3(grid), dim3(TC_THREADS), smem, st, ap, blob, ts, M, stages, epi, dbg, trace);
    DEFT_CHECK_LAUNCH("tc_igemm_kernel");
    return DEFT_OK;
}

template <class AP>
static int tc_dispatch(AP ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                       size_t scratch_bytes, cudaStream_t st) {
    if (blob == nullptr || bn <= 0 || ldw % bn != 0 || ep.Cout % 4 != 0) return DEFT_EINVAL;
    const int nkb = ceil_div(K, TC_BK), nt = ldw / bn;
    int splits = tc_pick_splits(ceil_div(M, TC_BM) * nt, nkb, std::is_same<AP, TcDcnAP>::value);
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
    launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float*>(scratch), splits, M, ldw, ep);
    DEFT_CHECK_LAUNCH("splitk_reduce");
    return DEFT_OK;
}

template <class AP>
static int ts_dispatch(AP ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                       size_t scratch_bytes, cudaStream_t st, int taps = 1) {
    if (blob == nullptr || bn <= 0 || bn > 128 || ldw % bn != 0 || ep.Cout % 4 != 0 || K % TC_BK != 0 || ldw > TS_SS_MAXC) return DEFT_EINVAL;
    const bool f16 = tc_blob_is_f16(K, taps, bn) && !(flags & DEFT_FLAG_TF32X1);
    if (tc_blob_is_f16(K, taps, bn) && !f16) return DEFT_EINVAL;      // the image is packed for 3xFP16: no single-pass TF32 variant of it
    const int nkb = K / (f16 ? 64 : TC_BK), nt = ldw / bn;
    int splits = tc_pick_splits(ceil_div(M, TC_BM) * nt, nkb, std::is_same<AP, TsDcnAP>::value);
    if (splits > 1 && (scratch == nullptr || scratch_bytes < (size_t)splits * M * ldw * sizeof(float))) splits = 1;
    if (splits > 1) splits = ceil_div(nkb, ceil_div(nkb, splits));
    EpiStd epi{ep, splits > 1 ? reinterpret_cast<float*>(scratch) : nullptr, ldw};
    const bool x1 = (flags & DEFT_FLAG_TF32X1) != 0;
    int rc = DEFT_EINVAL;
#define TS_CASE(BNV)                                                                          \
    case BNV:                                                                                 \
        rc = f16 ? ts_launch<AP, BNV, 6>(ap, blob, nkb, M, nt, splits, epi, st)               \
           : x1  ? ts_launch<AP, BNV, 1>(ap, blob, nkb, M, nt, splits, epi, st)               \
                 : ts_launch<AP, BNV, 3>(ap, blob, nkb, M, nt, splits, epi, st);              \
        break;
    switch (bn) {
        TS_CASE(128)
        TS_CASE(64)
        TS_CASE(32)
        TS_CASE(16)
    }
#undef TS_CASE
    if (rc != DEFT_OK || splits == 1) return rc;
    size_t total = (size_t)M * (ep.Cout / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > TC_NUM_SMS * 8) blocks = TC_NUM_SMS * 8;
    launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float*>(scratch), splits, M, ldw, ep);
    DEFT_CHECK_LAUNCH("splitk_reduce");
    return DEFT_OK;
}

static TcConvFastAP make_fast_ap(const ConvAP& ap) {
    TcConvFastAP f{};
    f.src = ap.src[0]; f.pitch = ap.pitch[0]; f.H = ap.H; f.W = ap.W; f.OH = ap.OH; f.OW = ap.OW;
    f.KW = ap.KW; f.stride = ap.stride; f.pad = ap.pad; f.taps = ap.KH * ap.KW;
    return f;
}

// ---- patch path (TsPatchAP): 3x3 / stride 1 / pad 1, one source, Cin % 32 == 0, and a map that tiles into 16x8
// patches without wasting more than ~15 % of the MMA rows.
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmapEncodeFn>(p);
    }
    return fn;
}
static bool patch_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DEFT_TC_PATCH"); v = e ? atoi(e) : 1; }
    return v != 0;
}
static bool patch_eligible(const ConvAP& ap) {
    if (!patch_enabled() || ap.nsrc != 1 || ap.KH != 3 || ap.KW != 3 || ap.stride != 1 || ap.pad != 1 || ap.Ctot % TC_BK != 0) return false;
    if (ap.OH != ap.H || ap.OW != ap.W || ap.pitch[0] % 4 != 0 || (reinterpret_cast<uintptr_t>(ap.src[0]) & 15) != 0) return false;
    const long long rows = (long long)ceil_div(ap.W, TP_PW) * ceil_div(ap.H, TP_PH) * TC_BM;
    return rows * 100 <= (long long)ap.H * ap.W * 115;
}
static int patch_dispatch(const ConvAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                          size_t scratch_bytes, cudaStream_t st) {
    if (blob == nullptr || bn <= 0 || bn > 128 || ldw % bn != 0 || ep.Cout % 4 != 0 || ldw > TS_SS_MAXC) return DEFT_EINVAL;
    TmapEncodeFn enc = tmap_encoder();
    if (enc == nullptr) return DEFT_EINVAL;
    TsPatchAP t{};
    t.H = ap.H; t.W = ap.W;
    t.tiles_x = ceil_div(ap.W, TP_PW);
    t.tiles_y = ceil_div(ap.H, TP_PH);
    const cuuint64_t gdim[4] = {(cuuint64_t)ap.Ctot, (cuuint64_t)ap.W, (cuuint64_t)ap.H, (cuuint64_t)ap.N};
    const cuuint64_t gstr[3] = {(cuuint64_t)ap.pitch[0] * 4, (cuuint64_t)ap.W * ap.pitch[0] * 4, (cuuint64_t)ap.H * ap.W * ap.pitch[0] * 4};
    const cuuint32_t box[4] = {TC_BK, TP_HW, TP_HH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ap.src[0]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return DEFT_EINVAL;
    const bool f16 = tc_blob_is_f16(K, 9, bn);
    if (f16 && (flags & DEFT_FLAG_TF32X1)) return DEFT_EINVAL;
    const int kch = f16 ? 64 : TC_BK;
    const int nkb = K / kch, nt = ldw / bn, ncb = ap.Ctot / kch;
    const int m_tiles = ap.N * t.tiles_x * t.tiles_y;
    int splits = tc_pick_splits(m_tiles * nt, nkb);
    if (splits > ncb) splits = ncb;
    if (splits > 1) splits = ceil_div(ncb, ceil_div(ncb, splits));     // whole channel blocks per split
    if (splits > 1 && (scratch == nullptr || scratch_bytes < (size_t)splits * M * ldw * sizeof(float))) splits = 1;
    EpiStd epi{ep, splits > 1 ? reinterpret_cast<float*>(scratch) : nullptr, ldw};
    const bool x1 = (flags & DEFT_FLAG_TF32X1) != 0;
    int rc = DEFT_EINVAL;
#define TP_CASE(BNV)                                                                                   \
    case BNV:                                                                                          \
        rc = f16 ? ts_launch<TsPatchAP, BNV, 6>(t, blob, nkb, M, nt, splits, epi, st, m_tiles)         \
           : x1  ? ts_launch<TsPatchAP, BNV, 1>(t, blob, nkb, M, nt, splits, epi, st, m_tiles)         \
                 : ts_launch<TsPatchAP, BNV, 3>(t, blob, nkb, M, nt, splits, epi, st, m_tiles);        \
        break;
    switch (bn) {
        TP_CASE(128)
        TP_CASE(64)
        TP_CASE(32)
        TP_CASE(16)
    }
#undef TP_CASE
    if (rc != DEFT_OK || splits == 1) return rc;
    size_t total = (size_t)M * (ep.Cout / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > TC_NUM_SMS * 8) blocks = TC_NUM_SMS * 8;
    launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float*>(scratch), splits, M, ldw, ep);
    DEFT_CHECK_LAUNCH("splitk_reduce");
    return DEFT_OK;
}

// ---- 16-channel patch path (TsPatch16AP): 3x3 / stride 1 / pad 1 over one 16-channel source, 3xFP16 image packed from K = 192
static int patch16_dispatch(const ConvAP& ap, const float* blob, int bn, int ldw, int M, const Epilogue& ep, unsigned flags, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("DEFT_TC_PATCH16"); on = e ? atoi(e) : 1; }
    if (!on || !patch_enabled() || blob == nullptr || (bn != 16 && bn != 32) || ldw % bn != 0 || ep.Cout % 4 != 0 || ldw > TS_SS_MAXC) return DEFT_EINVAL;
    if (!tc_blob_is_f16(192, 1, bn) || (flags & DEFT_FLAG_TF32X1)) return DEFT_EINVAL;
    if (ap.pitch[0] % 4 != 0 || (reinterpret_cast<uintptr_t>(ap.src[0]) & 15) != 0) return DEFT_EINVAL;
    const long long rows = (long long)ceil_div(ap.W, TP_PW) * ceil_div(ap.H, TP_PH) * TC_BM;
    if (rows * 100 > (long long)ap.H * ap.W * 115) return DEFT_EINVAL;
    TmapEncodeFn enc = tmap_encoder();
    if (enc == nullptr) return DEFT_EINVAL;
    TsPatch16AP t{};
    t.H = ap.H; t.W = ap.W;
    t.tiles_x = ceil_div(ap.W, TP_PW);
    t.tiles_y = ceil_div(ap.H, TP_PH);
    const cuuint64_t gdim[4] = {16, (cuuint64_t)ap.W, (cuuint64_t)ap.H, (cuuint64_t)ap.N};
    const cuuint64_t gstr[3] = {(cuuint64_t)ap.pitch[0] * 4, (cuuint64_t)ap.W * ap.pitch[0] * 4, (cuuint64_t)ap.H * ap.W * ap.pitch[0] * 4};
    const cuuint32_t box[4] = {16, TP_HW, TP_HH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ap.src[0]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return DEFT_EINVAL;
    const int m_tiles = ap.N * t.tiles_x * t.tiles_y;
    EpiStd epi{ep, nullptr, ldw};
    return bn == 16 ? ts_launch<TsPatch16AP, 16, 6>(t, blob, 3, M, ldw / bn, 1, epi, st, m_tiles)
                    : ts_launch<TsPatch16AP, 32, 6>(t, blob, 3, M, ldw / bn, 1, epi, st, m_tiles);
}

int tc_conv_dispatch(const ConvAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                     size_t scratch_bytes, cudaStream_t st) {
    for (int i = 0; i < ap.nsrc; ++i)
        if (ap.c[i] % 4 != 0) return DEFT_EINVAL;
    if (ts_enabled() && ap.nsrc == 1 && ap.c[0] == 16 && ap.KH == 3 && ap.KW == 3 && ap.pad == 1 && K == 144) {
        // the 16-channel full-resolution layers: tap-major pseudo channels, weight image packed from K = 192 (TC_C16_K) with taps = 1
        TsC16AP t{};
        t.src = ap.src[0]; t.pitch = ap.pitch[0]; t.N = ap.N; t.H = ap.H; t.W = ap.W; t.OH = ap.OH; t.OW = ap.OW; t.stride = ap.stride;
        if ((long long)ap.N * ap.H * ap.W >= 2147483647LL / 4) return DEFT_EINVAL;
        if (ap.stride == 1 && ap.OH == ap.H && ap.OW == ap.W) {
            const int rc16 = patch16_dispatch(ap, blob, bn, ldw, M, ep, flags, st);    // level0: operands from a TMA halo patch
            if (rc16 != DEFT_EINVAL) return rc16;
        }
        return ts_dispatch(t, blob, bn, ldw, 192, M, ep, flags, nullptr, 0, st, 1);
    }
    if (ts_enabled() && patch_eligible(ap) && K == 9 * ap.Ctot) {
        int rc = patch_dispatch(ap, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
        if (rc != DEFT_EINVAL) return rc;
    }
    if (ts_enabled() && bn <= 128 && ap.KH * ap.KW <= 9 && ap.Ctot % TC_BK == 0 && ap.Ctot / TC_BK <= 64) {
        bool ok = true;                                    // every source a multiple of 32 channels: a k-block never straddles sources
        for (int i = 0; i < ap.nsrc; ++i) ok = ok && (ap.c[i] % TC_BK == 0);
        if (ok) {
            TsConvAP t{};
            t.a = ap;
            t.taps = ap.KH * ap.KW;
            t.ncb = ap.Ctot / TC_BK;
            if (tc_blob_is_f16(K, t.taps, bn))                 // 64-channel k-blocks must not straddle sources
                for (int i = 0; i < ap.nsrc; ++i)
                    if (ap.c[i] % 64 != 0) return DEFT_EINVAL;
            return ts_dispatch(t, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st, t.taps);
        }
    }
    if (tc_blob_is_f16(K, ap.KH * ap.KW, bn)) return DEFT_EINVAL;   // the smem-operand kernel below reads 3xTF32 images only
    if (ap.nsrc == 1 && ap.Ctot % TC_BK == 0 && ap.KH * ap.KW <= 9) {
        // fast path: single source, channel-block-major K (what tc_k_perm_taps prescribes for Ctot % 32 == 0) or 1x1
        return tc_dispatch(make_fast_ap(ap), blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
    }
    TcConvAP t{};
    t.a = ap;
    t.K = K;
    t.perm_taps = tc_k_perm_taps(K, ap.KH * ap.KW);
    return tc_dispatch(t, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
}

// ---- DCN gather out of a TMA-loaded halo patch (TsDcnPatchAP): 3xFP16 images, N tile 64 or 128, maps that tile into 16x8
// patches without wasting more than ~15 % of the MMA rows (the 1/4 and 1/8 resolution layers; the deeper ones keep TsDcnAP).
static int dcn_patch_dispatch(const DcnAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                              size_t scratch_bytes, cudaStream_t st) {
    static int why = -1;
    const int on = dcn_patch_mode();      // default off: measured slower than the global gather (profiles/r2_dcn_patch_experiment.md)
    if (why < 0) { const char* e = getenv("DEFT_DEBUG_ENGINE"); why = e ? atoi(e) : 0; }
#define DP_DECLINE(reason) do { if (why) fprintf(stderr, "deft_b200: halo-patch DCN kernel declined (%s): Cin=%d ldw=%d bn=%d %dx%d\n", reason, ap.Cin, ldw, bn, ap.H, ap.W); return DEFT_EINVAL; } while (0)
    if (!on) return DEFT_EINVAL;
    if (blob == nullptr || (bn != 64 && bn != 128) || ldw % bn != 0 || ep.Cout % 4 != 0 || ldw > TS_SS_MAXC) DP_DECLINE("tile shape");
    if (!tc_blob_is_f16(K, 9, bn) || (flags & DEFT_FLAG_TF32X1) || K != 9 * ap.Cin || ap.Cin % 64 != 0) DP_DECLINE("not a 3xFP16 image");
    if (ap.pitch % 4 != 0 || (reinterpret_cast<uintptr_t>(ap.x) & 15) != 0) DP_DECLINE("alignment");
    const long long rows = (long long)ceil_div(ap.W, TP_PW) * ceil_div(ap.H, TP_PH) * TC_BM;
    if (on != 2 && rows * 100 > (long long)ap.H * ap.W * 115) DP_DECLINE("map wastes > 15 % of the patch rows");      // DEFT_DCN_PATCH=2: any map (tests)
    TmapEncodeFn enc = tmap_encoder();
    if (enc == nullptr) DP_DECLINE("no cuTensorMapEncodeTiled");
    TsDcnPatchAP t{};
    t.H = ap.H; t.W = ap.W; t.x = ap.x; t.pitch = ap.pitch; t.om = ap.om;
    t.tiles_x = ceil_div(ap.W, TP_PW);
    t.tiles_y = ceil_div(ap.H, TP_PH);
    const cuuint64_t gdim[4] = {(cuuint64_t)ap.Cin, (cuuint64_t)ap.W, (cuuint64_t)ap.H, (cuuint64_t)ap.N};
    const cuuint64_t gstr[3] = {(cuuint64_t)ap.pitch * 4, (cuuint64_t)ap.W * ap.pitch * 4, (cuuint64_t)ap.H * ap.W * ap.pitch * 4};
    const cuuint32_t box[4] = {TC_BK, DP_HW, DP_HH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult enc_rc = enc(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ap.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (enc_rc != CUDA_SUCCESS) {
        if (why) fprintf(stderr, "deft_b200: cuTensorMapEncodeTiled -> %d\n", (int)enc_rc);
        DP_DECLINE("tensor map encode failed");
    }
    const int nkb = K / 64, nt = ldw / bn, ncb = ap.Cin / 64;
    const int m_tiles = ap.N * t.tiles_x * t.tiles_y;
    int splits = tc_pick_splits(m_tiles * nt, nkb, true);
    if (splits > ncb) splits = ncb;
    if (splits > 1) splits = ceil_div(ncb, ceil_div(ncb, splits));     // whole channel blocks per split
    if (splits > 1 && (scratch == nullptr || scratch_bytes < (size_t)splits * M * ldw * sizeof(float))) splits = 1;
    EpiStd epi{ep, splits > 1 ? reinterpret_cast<float*>(scratch) : nullptr, ldw};
    int rc = bn == 64 ? ts_launch<TsDcnPatchAP, 64, 6>(t, blob, nkb, M, nt, splits, epi, st, m_tiles)
                      : ts_launch<TsDcnPatchAP, 128, 6>(t, blob, nkb, M, nt, splits, epi, st, m_tiles);
    if (rc == DEFT_EINVAL && why) fprintf(stderr, "deft_b200: halo-patch DCN kernel: launcher declined (bn=%d)\n", bn);
#undef DP_DECLINE
    if (rc == DEFT_OK) note_dcn_patch_launch();
    if (rc != DEFT_OK || splits == 1) return rc;
    size_t total = (size_t)M * (ep.Cout / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > TC_NUM_SMS * 8) blocks = TC_NUM_SMS * 8;
    launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float*>(scratch), splits, M, ldw, ep);
    DEFT_CHECK_LAUNCH("splitk_reduce");
    return DEFT_OK;
}

int tc_dcn_dispatch(const DcnAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, void* scratch,
                    size_t scratch_bytes, cudaStream_t st) {
    if (ap.Cin % 32 != 0 || bn > 128) return DEFT_EINVAL;
    static int ts_dcn = -1;
    if (ts_dcn < 0) { const char* e = getenv("DEFT_TC_TS_DCN"); ts_dcn = e ? atoi(e) : 1; }
    if (ts_dcn && ts_enabled() && (long long)M * ap.pitch < 2147483647LL) {
        const int rc_patch = dcn_patch_dispatch(ap, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
        if (rc_patch != DEFT_EINVAL) return rc_patch;
        TsDcnAP d{};
        d.a = ap;
        return ts_dispatch(d, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st, 9);
    }
    if (tc_blob_is_f16(K, 9, bn)) return DEFT_EINVAL;
    TcDcnAP t{};
    t.a = ap;
    t.K = K;
    return tc_dispatch(t, blob, bn, ldw, K, M, ep, flags, scratch, scratch_bytes, st);
}

int tc_pair_dispatch(const PairAP& ap, const float* blob, int bn, int ldw, int K, int M, const Epilogue& ep, unsigned flags, cudaStream_t st) {
    if (ts_enabled() && bn <= 128 && K % TC_BK == 0) {
        TsPairAP d{};
        d.a = ap;
        return ts_dispatch(d, blob, bn, ldw, K, M, ep, flags, nullptr, 0, st, 1);
    }
    if (tc_blob_is_f16(K, 1, bn)) return DEFT_EINVAL;
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
    if (ap.nsrc != 1 || ap.Ctot % TC_BK != 0 || ap.KH * ap.KW > 9) return DEFT_EINVAL;
    if (tc_blob_is_f16(K, ap.KH * ap.KW, 256)) {
        if (flags & DEFT_FLAG_TF32X1) return DEFT_EINVAL;
        TcConvFast16AP t16{};
        t16.src = ap.src[0]; t16.pitch = ap.pitch[0]; t16.H = ap.H; t16.W = ap.W; t16.OH = ap.OH; t16.OW = ap.OW;
        t16.KW = ap.KW; t16.stride = ap.stride; t16.pad = ap.pad; t16.taps = ap.KH * ap.KW;
        return tc_launch<TcConvFast16AP, 256, 6, EpiHeads>(t16, blob, K / 64, M, ha.nheads, 1, epi, st);
    }
    TcConvFastAP t = make_fast_ap(ap);
    const int nkb = ceil_div(K, TC_BK);
    if (flags & DEFT_FLAG_TF32X1) return tc_launch<TcConvFastAP, 256, 1, EpiHeads>(t, blob, nkb, M, ha.nheads, 1, epi, st);
    return tc_launch<TcConvFastAP, 256, 3, EpiHeads>(t, blob, nkb, M, ha.nheads, 1, epi, st);
}

int tc_pack_weights(const float* wp, int ldw, int K, int bn, int taps, float* blob, cudaStream_t st) {
    if (tc_blob_is_f16(K, taps, bn)) {
        const int nkb16 = K / 64, nt16 = ldw / bn;
        size_t total16 = (size_t)nt16 * nkb16 * bn * 64;
        int blocks16 = (int)((total16 + 255) / 256);
        if (blocks16 > TC_NUM_SMS * 8) blocks16 = TC_NUM_SMS * 8;
        tc_pack_weights_f16_kernel<<<blocks16, 256, 0, st>>>(wp, ldw, K, bn, nkb16, nt16, taps > 1 ? taps : 0, reinterpret_cast<__half*>(blob));
        DEFT_CHECK_LAUNCH("tc_pack_weights_f16");
        return DEFT_OK;
    }
    int nkb = ceil_div(K, TC_BK), nt = ldw / bn;
    size_t total = (size_t)nt * nkb * bn * TC_BK;
    int blocks = (int)((total + 255) / 256);
    if (blocks > TC_NUM_SMS * 8) blocks = TC_NUM_SMS * 8;
    tc_pack_weights_kernel<<<blocks, 256, 0, st>>>(wp, ldw, K, bn, nkb, nt, tc_k_perm_taps(K, taps), blob);
    DEFT_CHECK_LAUNCH("tc_pack_weights");
    return DEFT_OK;
}

}  // namespace deft
