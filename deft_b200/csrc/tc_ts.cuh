// tcgen05 "A through TMEM" kernel (TS-mode UMMA) for the narrow-N layers (N tile <= 128): included by tc_engine.cu.
//
// Measured on B200 (profiles/r1_tc_protocol_trace.md): with both operands in shared memory a 128x64x8 TF32 MMA
// takes ~70 cycles, because the 4 KB A tile is re-read from shared memory for every instruction -- the tensor pipe
// is operand-bandwidth bound, not math bound, whenever N is small.  Here the A tile never touches shared memory:
//   * producer thread = one accumulator row.  16 producer warps = 4 TMEM lane quadrants x 4 k-block groups; a warp
//     loads its 32 rows' 128 contiguous bytes per k-block (im2col tap / deformable gather / pair expansion), splits
//     them into tf32 hi + lo and writes them with tcgen05.st straight into TMEM columns (a ring of SA A-stages);
//   * two MMA warps issue tcgen05.mma with A from TMEM and only B (weights, bulk-copied) from shared memory; they
//     alternate k-blocks and own one accumulator each, which the epilogue adds in a fixed order (bit-deterministic);
//   * no generic->async proxy fence, no MEMBAR, no swizzle arithmetic on the A side;
//   * 3x3 / stride-1 convolutions take their operands from a TMA-loaded halo patch instead (TsPatchAP below);
//   * the epilogue transposes through shared memory so every global access is 128 contiguous bytes per row.
// Persistent tile loop, double-buffered accumulators, 3xTF32 split and split-K partials are as in tc_igemm_kernel.
#pragma once

namespace deft {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// 3xFP16 split of 16 consecutive K elements of one row: x = hi + lo + O(2^-22 x), hi = fp16(x), lo = fp16(x - hi).  A 32-bit
// TMEM cell holds two consecutive K elements, the EVEN one in the low half (measured: tools/exp_f16ts.py).
__device__ __forceinline__ void split_f16x3(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// cute::UMMA::InstrDescriptor for kind::f16: fp16 A and B (format 0), fp32 accumulate, K-major, M = 128, N = bn
__host__ __device__ constexpr uint32_t make_idesc_f16(int bn) { return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Programmatic dependent launch (common.cuh): here only what reads or writes ACTIVATIONS waits for the predecessor
// grid; barrier init, TMEM allocation, scale/shift staging and the weight ring start right away.  Every CTA has
// producer and epilogue warps that wait, so a CTA never retires before the predecessor grid has completed.
// ------------------------------------------------------------------------------------------------ producers
// A producer warp owns 32 rows (one TMEM lane quadrant) and every 4th k-block.  Per k-block it
//   1. fill(): loads the rows' 128 B with the COALESCED mapping (8 lanes x 16 B per row, 4 rows per LDG.128) -- the
//      thread=row mapping tcgen05.st needs would touch 32 cache lines per load instruction (8x the L1 wavefronts,
//      measured: no gain over the smem-operand kernel) -- and parks them in a private, padded smem stage [32][36];
//   2. the kernel reads the stage row-wise (lane = row, conflict-free LDS.128), splits hi/lo and tcgen05.st's.
// Interface:  kExtraSmemBytes; setup_once(extra); tile_init(m0, M, extra, quad, lane, meta) -- meta = this warp's
// private int4[32] row table in smem; fill(kb, quad, lane, meta, stage).
constexpr int TS_STAGE_LD = 36;                                   // floats per staged row (32 + 4 pad: conflict-free)
constexpr int TS_STAGE_FLOATS = 32 * TS_STAGE_LD;
constexpr int TS_WARP_SMEM = TS_STAGE_FLOATS * 4 + 32 * 16;       // stage + row table, per producer warp
constexpr int TS_COMMON_SMEM = TC_PWARPS * TS_WARP_SMEM;

// im2col over a channel concat of up to 4 NHWC sources, every channel count a multiple of 32.
// K order = (32-channel block, tap, channel) (tc_k_perm_taps) or 1x1.
struct TsSrcEntry { const float* p; int pitch; int coff; };
struct TsConvAP {
    static constexpr bool kPatch = false;
    ConvAP a;
    int taps, ncb;
    static constexpr int kExtraSmemBytes = 64 + 64 * 16;
    const int* toff_;
    const TsSrcEntry* cbsrc_;
    __device__ __forceinline__ int ktaps() const { return taps; }
    __device__ __forceinline__ void setup_once(unsigned char* extra) {
        int* toff = reinterpret_cast<int*>(extra);
        TsSrcEntry* cbsrc = reinterpret_cast<TsSrcEntry*>(extra + 64);
        if (threadIdx.x < taps) {
            int ky = threadIdx.x / a.KW, kx = threadIdx.x - ky * a.KW;
            toff[threadIdx.x] = (ky - a.pad) * a.W + (kx - a.pad);
        }
        if (threadIdx.x < ncb) {
            int cc = threadIdx.x * TC_BK, s = 0;
            while (s < a.nsrc - 1 && cc >= a.c[s]) { cc -= a.c[s]; ++s; }
            cbsrc[threadIdx.x].p = a.src[s];
            cbsrc[threadIdx.x].pitch = a.pitch[s];
            cbsrc[threadIdx.x].coff = cc;
        }
        toff_ = toff;
        cbsrc_ = cbsrc;
        asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char*, int quad, int lane, int4* meta) {
        const int p = m0 + quad * 32 + lane;
        int pix = 0;
        unsigned vmask = 0;
        if (p < M) {
            const int n = p / (a.OH * a.OW);
            const int r = p - n * (a.OH * a.OW);
            const int oy = (r / a.OW) * a.stride, ox = (r - (r / a.OW) * a.OW) * a.stride;
            pix = (n * a.H + oy) * a.W + ox;
            unsigned bit = 1u;
            for (int ky = 0; ky < a.KH; ++ky)
                for (int kx = 0; kx < a.KW; ++kx, bit <<= 1) {
                    const int iy = oy - a.pad + ky, ix = ox - a.pad + kx;
                    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) vmask |= bit;
                }
        }
        meta[lane] = make_int4(pix, (int)vmask, 0, 0);
        __syncwarp();
    }
    __device__ __forceinline__ void fill(int kb, int, int lane, const int4* meta, float* stage) const {
        const int cb = kb / taps, tap = kb - cb * taps;
        const TsSrcEntry e = cbsrc_[cb];
        const int toff = toff_[tap];
        const int ch = (lane & 7) * 4;
        const float* base = e.p + e.coff + ch;
        float4 r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int4 m = meta[(lane >> 3) + 4 * j];
            r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (((unsigned)m.y >> tap) & 1u) r[j] = ldg4(base + (size_t)(m.x + toff) * e.pitch);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stage + ((lane >> 3) + 4 * j) * TS_STAGE_LD + ch) = r[j];
    }
};

// 3x3 / pad 1 convolution over ONE 16-channel NHWC source (DLA level0 16->16 s1, level1 16->32 s2; r1 ran them as fp32 FMA kernels:
// 194 us of a 3.4 ms frame).  K is laid out tap-major exactly as the packed weight (k = tap * 16 + c), so a 32-wide pseudo
// channel block is two taps; the weight image is packed from a copy zero-padded to K = 192 (taps 9..11 do not exist: their
// operand rows are zeros too).  A row's 128 bytes = 2 taps x 64 B: lanes 0-3 of the row's 8 read tap 2b, lanes 4-7 tap 2b+1.
struct TsC16AP {
    static constexpr bool kPatch = false;
    const float* src;
    int pitch, N, H, W, OH, OW, stride;
    static constexpr int kExtraSmemBytes = 0;
    __device__ __forceinline__ int ktaps() const { return 1; }
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char*, int quad, int lane, int4* meta) {
        const int p = m0 + quad * 32 + lane;
        int pix = 0;
        unsigned vmask = 0;
        if (p < M) {
            const int n = p / (OH * OW);
            const int r = p - n * (OH * OW);
            const int oy = (r / OW) * stride, ox = (r - (r / OW) * OW) * stride;
            pix = (n * H + oy) * W + ox;
            unsigned bit = 1u;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx, bit <<= 1) {
                    const int iy = oy - 1 + ky, ix = ox - 1 + kx;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W) vmask |= bit;
                }
        }
        meta[lane] = make_int4(pix, (int)vmask, 0, 0);
        __syncwarp();
    }
    __device__ __forceinline__ void fill(int kb, int, int lane, const int4* meta, float* stage) const {
        const int tap = 2 * kb + ((lane & 7) >> 2);
        const int ky = tap / 3, kx = tap - ky * 3;
        const int toff = (ky - 1) * W + (kx - 1);
        const int ch = (lane & 3) * 4;
        float4 r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int4 m = meta[(lane >> 3) + 4 * j];
            r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tap < 9 && (((unsigned)m.y >> tap) & 1u)) r[j] = ldg4(src + (size_t)(m.x + toff) * pitch + ch);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stage + ((lane >> 3) + 4 * j) * TS_STAGE_LD + (lane & 7) * 4) = r[j];
    }
};

// DCNv2 bilinear gather; K order = (32-channel block, tap, channel).  The corner table of the tile's 128 rows x 9 taps
// lives in smem (built once per tile by all producers), 20 bytes per (row, tap):
//   tabB: element offset of the clamped top-left corner (a multiple of 4) | dx | dy << 1   (dx, dy: 1 = the right /
//         lower neighbour exists, 0 = it coincides with the clamped corner and carries zero weight or IS that row/col)
//   tabW: {wy0, wy1, wx0, wx1} with the validity of each corner row / column and the sigmoid mask folded in:
//         w(corner y, corner x) = wy * wx.
// Compact on purpose: the gather runs through L1, and every KB of shared memory the CTA does not claim is L1
// (r1 ncu: 56 % L1 hit rate with 225 KB of smem configured; the corner neighbourhood of a tile is ~27 KB per
// channel block).
struct TsDcnAP {
    static constexpr bool kPatch = false;
    DcnAP a;
    static constexpr int kExtraSmemBytes = TC_BM * 9 * 20;
    const int* tabB_;
    const float4* tabW_;
    int rowpitch_;
    __device__ __forceinline__ int ktaps() const { return 9; }
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char* extra, int, int, int4*) {
        float4* tabW = reinterpret_cast<float4*>(extra);
        int* tabB = reinterpret_cast<int*>(extra + TC_BM * 9 * 16);
        const int H = a.H, W = a.W;
        asm volatile("bar.sync 1, 512;" ::: "memory");     // every producer is done with the previous tile's table
        for (int idx = threadIdx.x; idx < TC_BM * 9; idx += TC_PRODUCERS) {
            int r9 = idx / 9, tap = idx - r9 * 9;
            int p = m0 + r9;
            int code = 0;
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
                    const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= W - 1;
                    w.x = y0ok ? hy : 0.f;
                    w.y = y1ok ? ly : 0.f;
                    w.z = x0ok ? hx * mk : 0.f;
                    w.w = x1ok ? lx * mk : 0.f;
                    const int yc = y0ok ? y0 : 0, xc = x0ok ? x0 : 0;     // y0 = -1: the lower row IS row 0 (dy = 0)
                    code = ((n * H + yc) * W + xc) * a.pitch | ((x0ok && x1ok) ? 1 : 0) | ((y0ok && y1ok) ? 2 : 0);
                }
            }
            tabB[idx] = code;    // all-zero weights point at element 0: always a valid address
            tabW[idx] = w;
        }
        tabB_ = tabB;
        tabW_ = tabW;
        rowpitch_ = W * a.pitch;
        asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    __device__ __forceinline__ void fill(int kb, int quad, int lane, const int4*, float* stage) const {
        const int cb = kb / 9, tap = kb - cb * 9;
        const int ch = (lane & 7) * 4;
        const float* x = a.x + cb * TC_BK + ch;
        const int pitch = a.pitch, rowpitch = rowpitch_;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {                   // two rows (8 corner loads) in flight at a time
            const int r0 = (lane >> 3) + 8 * jj, r1 = r0 + 4;
            const int c0 = tabB_[(quad * 32 + r0) * 9 + tap], c1 = tabB_[(quad * 32 + r1) * 9 + tap];
            const float4 w0 = tabW_[(quad * 32 + r0) * 9 + tap], w1 = tabW_[(quad * 32 + r1) * 9 + tap];
            const float* p0 = x + (c0 & ~3);
            const float* p1 = x + (c1 & ~3);
            const int dx0 = (c0 & 1) ? pitch : 0, dy0 = (c0 & 2) ? rowpitch : 0;
            const int dx1 = (c1 & 1) ? pitch : 0, dy1 = (c1 & 2) ? rowpitch : 0;
            const float4 a0 = ldg4(p0), a1 = ldg4(p0 + dx0), a2 = ldg4(p0 + dy0), a3 = ldg4(p0 + dy0 + dx0);
            const float4 b0 = ldg4(p1), b1 = ldg4(p1 + dx1), b2 = ldg4(p1 + dy1), b3 = ldg4(p1 + dy1 + dx1);
            const float u00 = w0.x * w0.z, u01 = w0.x * w0.w, u10 = w0.y * w0.z, u11 = w0.y * w0.w;
            const float v00 = w1.x * w1.z, v01 = w1.x * w1.w, v10 = w1.y * w1.z, v11 = w1.y * w1.w;
            float4 u, v;
            u.x = u00 * a0.x + u01 * a1.x + u10 * a2.x + u11 * a3.x; u.y = u00 * a0.y + u01 * a1.y + u10 * a2.y + u11 * a3.y;
            u.z = u00 * a0.z + u01 * a1.z + u10 * a2.z + u11 * a3.z; u.w = u00 * a0.w + u01 * a1.w + u10 * a2.w + u11 * a3.w;
            v.x = v00 * b0.x + v01 * b1.x + v10 * b2.x + v11 * b3.x; v.y = v00 * b0.y + v01 * b1.y + v10 * b2.y + v11 * b3.y;
            v.z = v00 * b0.z + v01 * b1.z + v10 * b2.z + v11 * b3.z; v.w = v00 * b0.w + v01 * b1.w + v10 * b2.w + v11 * b3.w;
            *reinterpret_cast<float4*>(stage + r0 * TS_STAGE_LD + ch) = u;
            *reinterpret_cast<float4*>(stage + r1 * TS_STAGE_LD + ch) = v;
        }
    }
};

// Affinity MLP layer 2: A[p, k] = relu(Au'[f(p), i(p), k] + Bv'[j(p), k]) over the live pairs.
struct TsPairAP {
    static constexpr bool kPatch = false;
    PairAP a;
    static constexpr int kExtraSmemBytes = 0;
    __device__ __forceinline__ int ktaps() const { return 1; }
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char*, int quad, int lane, int4* meta) {
        const int p = m0 + quad * 32 + lane;
        long long offa = -1, offb = 0;                     // element offsets into au / bv; -1 = dead row
        if (p < M) {
            int f = 0;
            while (f < a.F - 1 && p >= a.off[f + 1]) ++f;
            int q = p - a.off[f];
            int ii = q / a.m, jj = q - ii * a.m;
            offa = ((long long)f * a.P + ii) * a.ld;
            offb = (long long)f * a.bvf + (long long)jj * a.ld;
        }
        meta[lane] = make_int4((int)(offa & 0xffffffffLL), (int)(offa >> 32), (int)(offb & 0xffffffffLL), (int)(offb >> 32));
        __syncwarp();
    }
    __device__ __forceinline__ void fill(int kb, int, int lane, const int4* meta, float* stage) const {
        const int k = kb * TC_BK + (lane & 7) * 4;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            float4 xa[4], xb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int4 m = meta[(lane >> 3) + 4 * (jj * 4 + j)];
                const long long offa = ((long long)m.y << 32) | (unsigned)m.x, offb = ((long long)m.w << 32) | (unsigned)m.z;
                xa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                xb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (offa >= 0) { xa[j] = ldg4(a.au + offa + k); xb[j] = ldg4(a.bv + offb + k); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 v;
                v.x = fmaxf(xa[j].x + xb[j].x, 0.f); v.y = fmaxf(xa[j].y + xb[j].y, 0.f);
                v.z = fmaxf(xa[j].z + xb[j].z, 0.f); v.w = fmaxf(xa[j].w + xb[j].w, 0.f);
                *reinterpret_cast<float4*>(stage + ((lane >> 3) + 4 * (jj * 4 + j)) * TS_STAGE_LD + (lane & 7) * 4) = v;
            }
        }
    }
};

// 3x3 / stride 1 / pad 1 convolution over ONE NHWC source with Cin % 32 == 0: the "patch" producer.
// An M tile is a 16 x 8 pixel patch of the output (not 128 consecutive pixels), so the 9 taps of a 32-channel block
// all read the same (16+2) x (8+2) input halo patch.  One TMA tensor copy (box {32 ch, 18, 10, 1}, SWIZZLE_128B,
// out-of-bounds = zero fill = the conv's zero padding) per (tile, channel block) lands it in shared memory; producer
// thread = output pixel reads its tap's 128 B with conflict-free LDS.128 (the TMA swizzle spreads 8 consecutive
// pixels over all banks), splits hi/lo and tcgen05.st's.  No address tables, no global-load latency in the producers,
// and the input is fetched from L2 ~1.4x instead of 9x.
constexpr int TP_PW = 16, TP_PH = 8;                               // output patch
constexpr int TP_HW = TP_PW + 2, TP_HH = TP_PH + 2;                // halo patch
constexpr int TP_BOX_BYTES = TP_HW * TP_HH * 128;                  // 23040: what one TMA copy delivers
constexpr int TP_STAGE_BYTES = (TP_BOX_BYTES + 1023) / 1024 * 1024;   // ring stages stay 1024 B aligned (swizzle atom)
constexpr int TP_MAX_STAGES = 4;
struct TsPatchAP {
    static constexpr bool kPatch = true;
    static constexpr bool kDcn = false;
    static constexpr int kHW = TP_HW, kHH = TP_HH, kHalo = 1, kBoxBytes = TP_BOX_BYTES, kStageBytes = TP_STAGE_BYTES;
    static constexpr int kExtraSmemBytes = 0;
    alignas(64) CUtensorMap tmap;                                  // dims {C, W, H, N} of the NHWC source
    int H, W, tiles_x, tiles_y;
    // interface stubs (the patch path has its own producer loop)
    __device__ __forceinline__ int ktaps() const { return 9; }
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int, int, unsigned char*, int, int, int4*) {}
    __device__ __forceinline__ void fill(int, int, int, const int4*, float*) const {}
};

// DCNv2 gather out of a halo patch (round 2).  The global-memory gather (TsDcnAP above) is bound by the latency of its
// corner loads: 4 serial rounds of 8 LDG.128 per lane and k-block, nearly every round containing an L1 miss (the 100 KB
// neighbourhood of a 128-pixel row tile does not fit the ~50 KB of L1 left beside the operand rings).  Here an M tile is
// a 16 x 8 pixel patch as in TsPatchAP, and ONE TMA tensor copy per 32-channel block lands the (16+2+2R) x (8+2+2R) halo
// in shared memory (R = DP_R pixels of margin for the learned offsets).  Producer thread = output pixel: it recomputes
// its sampling position from the offset/mask row (3 cached loads per tap; no corner table, no table-build barrier),
// reads the four corners with LDS.128 (the TMA swizzle keeps neighbouring pixels on different banks), blends, splits and
// tcgen05.st's -- no transposition stage either.  A corner pair that leaves the halo (|offset| >= R) falls back to the
// global loads of the old path for that pixel and tap: any offset is handled, small ones are fast.
constexpr int DP_R = 1;
constexpr int DP_HW = TP_PW + 2 + 2 * DP_R, DP_HH = TP_PH + 2 + 2 * DP_R;       // 20 x 12
constexpr int DP_BOX_BYTES = DP_HW * DP_HH * 128;                               // 30720
constexpr int DP_STAGE_BYTES = (DP_BOX_BYTES + 1023) / 1024 * 1024;
struct TsDcnPatchAP {
    static constexpr bool kPatch = true;
    static constexpr bool kDcn = true;
    static constexpr int kHW = DP_HW, kHH = DP_HH, kHalo = 1 + DP_R, kBoxBytes = DP_BOX_BYTES, kStageBytes = DP_STAGE_BYTES;
    static constexpr int kExtraSmemBytes = 0;
    alignas(64) CUtensorMap tmap;                                  // dims {C, W, H, N} of the NHWC input
    int H, W, tiles_x, tiles_y;
    const float* x;                                                // the same input through plain pointers (out-of-halo corners)
    int pitch;
    const float* om;                                               // [P, 32] offsets (2k = dy, 2k+1 = dx) and mask logits (18 + k)
    __device__ __forceinline__ int ktaps() const { return 9; }
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int, int, unsigned char*, int, int, int4*) {}
    __device__ __forceinline__ void fill(int, int, int, const int4*, float*) const {}
};
// 3x3 / stride 1 / pad 1 convolution over ONE 16-channel NHWC source (DLA level0 at full resolution) through the halo patch:
// the same 18 x 10 patch, 64 bytes per pixel (TMA box {16 ch, 18, 10, 1}, SWIZZLE_64B), ONE patch per tile for all nine
// taps.  K is tap-major as in TsC16AP (k = tap * 16 + c, the weight image packed from K = 192): a 64-wide k-block is four
// taps, three k-blocks per tile, taps 9..11 are zeros.  Producer thread = output pixel reads 4 x 64 B per k-block with
// conflict-free LDS.128 instead of TsC16AP's global gathers (16 LDG.128 per lane and k-block, 8 cache lines each, behind a
// staging transposition): the level0 layer was producer-latency bound at 4.9 k cycles per 128-pixel tile.
constexpr int TP16_BOX_BYTES = TP_HW * TP_HH * 64;                 // 11520
constexpr int TP16_STAGE_BYTES = (TP16_BOX_BYTES + 1023) / 1024 * 1024;
struct TsPatch16AP {
    static constexpr bool kPatch = true;
    static constexpr bool kDcn = false;
    static constexpr bool kC16 = true;
    static constexpr int kHW = TP_HW, kHH = TP_HH, kHalo = 1, kBoxBytes = TP16_BOX_BYTES, kStageBytes = TP16_STAGE_BYTES;
    static constexpr int kExtraSmemBytes = 0;
    alignas(64) CUtensorMap tmap;                                  // dims {16, W, H, N} of the NHWC source
    int H, W, tiles_x, tiles_y;
    __device__ __forceinline__ int ktaps() const { return 1; }
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int, int, unsigned char*, int, int, int4*) {}
    __device__ __forceinline__ void fill(int, int, int, const int4*, float*) const {}
};
template <class AP, bool P = AP::kPatch>
struct TsPG {                                                      // patch geometry of an A producer (none for the gather producers)
    static constexpr bool DCN = false, C16 = false;
    static constexpr int HW = 1, HALO = 0, BOX = 0, STAGE = 0, KPP = 1;
};
template <class AP, class = void> struct TsIsC16 { static constexpr bool value = false; };
template <class AP> struct TsIsC16<AP, typename std::enable_if<AP::kC16>::type> { static constexpr bool value = true; };
template <class AP>
struct TsPG<AP, true> {
    static constexpr bool DCN = AP::kDcn, C16 = TsIsC16<AP>::value;
    static constexpr int HW = AP::kHW, HALO = AP::kHalo, BOX = AP::kBoxBytes, STAGE = AP::kStageBytes;
    static constexpr int KPP = C16 ? 3 : 9;                         // k-blocks that share one patch (one tile's taps / one channel block's taps)
};
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ kernel
// dynamic shared memory the TS kernel may ask for: the 227 KB opt-in limit minus its 64 bytes of static shared memory (progress board)
constexpr int TS_SMEM_LIMIT = TC_SMEM_LIMIT - 128;
constexpr int TS_KGROUPS = TC_PWARPS / 4;       // k-block interleave across producer warp groups
// Two MMA issuer warps alternate k-blocks (tcgen05.mma issue blocks the issuing thread, so one issuer leaves the tensor
// pipe idle during its own barrier waits and commits: 53 -> 43 us on the 64->64 3x3 layer).  Each issuer accumulates
// into its OWN TMEM accumulator and the epilogue adds the two in a fixed order: with a shared accumulator the order in
// which the two issuers' MMAs land is timing dependent and results stop being bit-identical run to run
// (tests/test_gpu_frame.py::test_c2_full_size_properties caught exactly that).
constexpr int TS_ISSUERS = 2;
#ifndef DEFT_TS_SETMAXNREG
#define DEFT_TS_SETMAXNREG 0  // tried in r2: ptxas keeps the producers' spill count and the 56-register epilogue gains 60 reloads; off
#endif
// 16 producer warps x 32 x (96 - 80) = 8192 registers taken = 4 x 32 x (80 - 56) + 4 x 32 x (80 - 40) released
#define DEFT_TS_PROD_REGS 96
#define DEFT_TS_EPI_REGS 56
#define DEFT_TS_AUX_REGS 40
#ifndef DEFT_TS_CHAINS
#define DEFT_TS_CHAINS 0      // 1 = several accumulate chains per issuer at N <= 32 (TsCfg::CH): parity-green, measured neutral (the issuer's own waits bound N <= 64), off
#endif
#ifndef DEFT_TS_PAIR32
#define DEFT_TS_PAIR32 0     // 1 = the N = 32 instantiation issues [hi|lo] as one N = 64 MMA (see TsCfg::PAIR); measured neutral in r2 (producer-bound), off
#endif
#ifndef DEFT_TS_ISS128
#define DEFT_TS_ISS128 2      // issuer warps of the N = 128 instantiations (A/B builds: make EXTRA=-DDEFT_TS_ISS128=1 OUT=...)
#endif

template <int BN, int NPASS>
struct TsCfg {
    // NPASS: 1 = single TF32 pass, 3 = 3xTF32 (k-block = 32 channels), 6 = 3xFP16 (kind::f16, k-block = 64 channels: the same
    // 64 TMEM columns and the same 128-byte weight rows hold twice the K, every MMA covers K = 16 instead of 8).
    static constexpr bool F16 = NPASS == 6;
    static constexpr int PASSES = F16 ? 3 : NPASS;
    static constexpr int KCH = F16 ? 64 : 32;                          // channels per k-block
    static constexpr int ACOLS = PASSES == 3 ? 64 : 32;                // TMEM columns of one A stage (hi [+ lo])
    // Issuer warps: two, alternating k-blocks (a 128x64x8 MMA executes in ~32 cycles but costs ~46 to issue, and an
    // issuer also pays its own barrier waits and commits serially).  r2 tried ONE issuer for N = 128 so that the two
    // accumulator sets fit twice in TMEM (double-buffered epilogue): the pipe, at its floor with two issuers (766 cycles
    // per k-block = 12 x 64), fell to 1351 cycles per k-block -- 43.0 vs 35.8 us on the 128->128 3x3 layer
    // (profiles/r2_engine_experiments.md).  DEFT_TS_ISS128 keeps the variant buildable.
    static constexpr int ISS = BN >= 128 ? DEFT_TS_ISS128 : TS_ISSUERS;
    // N = 32 (the DCN offset/mask convs), 3xTF32: a TS-mode MMA costs ~the same to issue at N = 32 as at N = 64, so the
    // two products that share A_hi are issued as ONE N = 64 instruction against the stage's [B_hi ; B_lo] rows (they
    // are contiguous and form a valid 64-row swizzled tile): 2 MMAs per k-step instead of 3.  The accumulator of an
    // issuer is then 2 x BN columns ([hi*hi + lo*hi | hi*lo]) and the epilogue adds the column blocks.
    static constexpr bool PAIR = (BN == 32 && NPASS == 3 && DEFT_TS_PAIR32);
    // Accumulate chains per issuer (3xFP16, narrow tiles).  A tcgen05.mma into an accumulator that the previous MMA is still
    // updating waits for it: 96.8 cycles per MMA whatever N is (profiles/r2_engine_experiments.md section 5), so at N <= 32 --
    // where the MMA itself takes 8-16 cycles -- the two issuers' two chains cap a k-block at 12 x 97 / 2 = 581 cycles while the
    // producers need ~330.  The k-steps of a k-block therefore rotate over CH accumulators per issuer (N = 32: 2, N = 16: 4;
    // the accumulator columns are there: 2 sets x 2 issuers x CH x BN = 256); the epilogue adds the ISS x CH partial tiles in a
    // fixed order.  Measured (r2, tools/gpu_call16.sh): tests green, frame 593 vs 597 fps -- the N = 32 issuer spends 1715 cycles per
    // own k-block of which 542 issue MMAs, the rest are its barrier waits and commits, so the chain is not what binds.  Default off.
    static constexpr int CH = (F16 && DEFT_TS_CHAINS && BN <= 32) ? (BN == 32 ? 2 : 4) : 1;
    static constexpr int ACCW = PAIR ? 2 * BN : CH * BN;                // accumulator columns of one issuer
    static constexpr int NBLK = (PAIR ? 2 : CH);                        // BN-wide column blocks per issuer
    static constexpr int NBUF = (2 * ISS * ACCW <= 256) ? 2 : 1;        // accumulator sets (double-buffered across tiles when they fit)
    static constexpr int A0 = NBUF * ISS * ACCW;                        // first A column (after the accumulators)
    static constexpr int SA_RAW = (512 - A0) / ACOLS;
    static constexpr int SA = SA_RAW > 6 ? 6 : SA_RAW;                  // A ring depth
    static constexpr int B_TILE = BN * 128;
    static constexpr int B_STAGE = PASSES == 3 ? 2 * B_TILE : B_TILE;
};

// Accumulator tile -> output, 32 (or BN) columns at a time.  tcgen05.ld hands every lane one accumulator ROW; storing
// rows from there touches 32 different lines per store instruction and fills only half of every 32-byte sector.  So
// each epilogue warp parks its 32 x LW block in a padded smem stage and re-reads it with LW/4 lanes per row: every
// global access (output, residual, split-K partial) is then 128 contiguous bytes per row.  Scale / shift sit in smem.
constexpr int TS_EPI_LD = 36;                                       // floats per staged row (32 + 4 pad)
constexpr int TS_SS_MAXC = 512;                                     // widest layer whose scale/shift are staged in smem
constexpr int TS_EPI_SMEM = TC_EWARPS * 32 * TS_EPI_LD * 4 + 2 * TS_SS_MAXC * 4;    // transpose stages + scale + shift
struct TsRowMap {                                                   // accumulator row (0..127) of this tile -> output pixel
    int base, H, W, y0, x0, M;                                      // patch: y0/x0/H/W, base = n*H*W; linear: base = m0, W = 0
    __device__ __forceinline__ int pixel(int row, bool& valid) const {
        if (W == 0) { const int p = base + row; valid = p < M; return p; }
        const int y = y0 + (row >> 4), x = x0 + (row & 15);
        valid = y < H && x < W;
        return base + y * W + x;
    }
};
template <int BN>
__device__ __forceinline__ void ts_epilogue(const EpiStd& epi, uint32_t tmem_acc, int nblk, float gain, int ewarp, int lane, const TsRowMap& rm, int ntile,
                                            int z, float* stage, const float* ss) {   // nblk consecutive BN-wide partial accumulators are added in order
    const Epilogue& e = epi.e;
    constexpr int LW = BN >= 32 ? 32 : BN;
    constexpr int LPR = LW / 4;                                     // lanes per row in the coalesced mapping
    constexpr int RPI = 32 / LPR;                                   // rows per instruction
    const uint32_t lane_addr = (uint32_t)(ewarp * 32) << 16;
    const int cq = (lane % LPR) * 4, rsub = lane / LPR;
#pragma unroll 1
    for (int c = 0; c < BN; c += LW) {
        __syncwarp();                                               // the previous block's readers are done with the stage
        constexpr int LH = LW >= 16 ? 16 : LW;                      // 16 columns at a time keeps the register count down
#pragma unroll
        for (int h = 0; h < LW; h += LH) {
            float v[LH];
            if ((nblk & 1) == 0) {                                  // the common case (two issuers [x chains]): two loads per wait
#pragma unroll 1
                for (int q = 0; q < nblk; q += 2) {
                    uint32_t r0[LH], r1[LH];
                    tmem_ld_issue<LH>(tmem_acc + (uint32_t)(q * BN) + lane_addr + (uint32_t)(c + h), r0);
                    tmem_ld_issue<LH>(tmem_acc + (uint32_t)((q + 1) * BN) + lane_addr + (uint32_t)(c + h), r1);
                    tmem_ld_wait();
                    if (q == 0) {
#pragma unroll
                        for (int j = 0; j < LH; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < LH; ++j) v[j] = (v[j] + __uint_as_float(r0[j])) + __uint_as_float(r1[j]);
                    }
                }
            } else {
                tmem_ld<LH>(tmem_acc + lane_addr + (uint32_t)(c + h), v);
#pragma unroll 1
                for (int q = 1; q < nblk; ++q) {                    // the other partial accumulators, added in a fixed order
                    float v2[LH];
                    tmem_ld<LH>(tmem_acc + (uint32_t)(q * BN) + lane_addr + (uint32_t)(c + h), v2);
#pragma unroll
                    for (int j = 0; j < LH; ++j) v[j] += v2[j];
                }
            }
#pragma unroll
            for (int q = 0; q < LH / 4; ++q)
                *reinterpret_cast<float4*>(stage + lane * TS_EPI_LD + h + q * 4) =
                    make_float4(v[q * 4] * gain, v[q * 4 + 1] * gain, v[q * 4 + 2] * gain, v[q * 4 + 3] * gain);
        }
        __syncwarp();
        const int col = ntile * BN + c + cq;                        // this lane's 4 output channels
        if (epi.partial != nullptr) {                               // split-K: raw partial sums, reduced by splitk_reduce_kernel
#pragma unroll
            for (int j = 0; j < 32 / RPI; ++j) {
                const int row = rsub + RPI * j;
                bool valid;
                const int p = rm.pixel(ewarp * 32 + row, valid);
                if (valid) *reinterpret_cast<float4*>(epi.partial + ((size_t)z * rm.M + p) * epi.ldp + col) = *reinterpret_cast<const float4*>(stage + row * TS_EPI_LD + cq);
            }
            continue;
        }
        const bool col_ok = col < e.Cout;
        const float4 sc = *reinterpret_cast<const float4*>(ss + (col_ok ? col : 0));
        const float4 sh = *reinterpret_cast<const float4*>(ss + TS_SS_MAXC + (col_ok ? col : 0));
        constexpr int NJ = 32 / RPI, JB = NJ >= 4 ? 4 : NJ;        // rows of this lane, JB at a time: loads first, then the stores
#pragma unroll
        for (int j0 = 0; j0 < NJ; j0 += JB) {
            float4 a[JB], rr[JB];
            int pp[JB];
            bool ok[JB];
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const int row = rsub + RPI * (j0 + j);
                a[j] = *reinterpret_cast<const float4*>(stage + row * TS_EPI_LD + cq);
                bool valid;
                pp[j] = rm.pixel(ewarp * 32 + row, valid);
                ok[j] = valid && col_ok;
                rr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[j] && e.residual) rr[j] = ldg4(e.residual + (size_t)pp[j] * e.res_pitch + col);
            }
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                if (!ok[j]) continue;
                float4 o;
                o.x = a[j].x * sc.x + sh.x + rr[j].x; o.y = a[j].y * sc.y + sh.y + rr[j].y;
                o.z = a[j].z * sc.z + sh.z + rr[j].z; o.w = a[j].w * sc.w + sh.w + rr[j].w;
                if (e.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(e.out + (size_t)pp[j] * e.out_pitch + col) = o;
            }
        }
    }
}

template <class AP, int BN, int NPASS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_ts_kernel(const __grid_constant__ AP ap_param, const float* __restrict__ wblob, TileSched ts,
                                                              int M, int sb, int ps, EpiStd epi, int dbg, long long* trace, float rz_comp) {
    using Cfg = TsCfg<BN, NPASS>;
    constexpr int SA = Cfg::SA;
    constexpr bool kPatch = AP::kPatch;
    using PG = TsPG<AP>;
    constexpr int P_SUB = PG::STAGE;                               // one 32-channel (C16: 16-channel) halo patch
    constexpr int P_STAGE = (Cfg::F16 && !PG::C16) ? 2 * P_SUB : P_SUB;   // halo patch of one k-block's channels
    constexpr int KPP = PG::KPP;
    // the gather producers keep per-thread state in their AP (private copy); the patch AP holds the TMA descriptor,
    // which must stay in param space
    typename std::conditional<kPatch, const AP&, AP>::type ap = ap_param;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    unsigned char* smem = smem_raw + pad;
    // layout: [patch ring: ps x TP_STAGE_BYTES (patch producer only)] [sb x B_STAGE] [barriers + tmem slot: 512 B] [epilogue stages]
    //         [producer stages (gather producers only)] [AP extra]
    const uint32_t patch_bytes = kPatch ? (uint32_t)ps * P_STAGE : 0u;
    const uint32_t sbase = raw_addr + pad + patch_bytes;  // weight ring
    const uint32_t bar_base = sbase + (uint32_t)sb * Cfg::B_STAGE;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (8 + s); };
    // b_full: TWO barriers per weight stage, used by the even / odd uses of the stage.  With an odd ring depth (3 stages is what
    // fits at N = 128) consecutive uses of a stage belong to different issuers, bulk copies land out of order, and an issuer
    // that ran ahead saw its one-bit parity wait satisfied by the PREVIOUS use's phase (consumed a stage whose copy was still
    // in flight, then desynchronised the ring: the second stall of round 2).  Split by use parity, every barrier is owned by
    // one issuer for the whole kernel (uses u, u+2, ... are 2*sb k-blocks apart: same parity), so its waits are sequential.
    auto b_full = [&](int s, int u) { return bar_base + 8u * (48 + 2 * s + (u & 1)); };
    auto b_empty = [&](int s) { return bar_base + 8u * (24 + s); };
    auto acc_full = [&](int b) { return bar_base + 8u * (32 + b); };
    auto acc_empty = [&](int b) { return bar_base + 8u * (34 + b); };
    auto p_full = [&](int s) { return bar_base + 8u * (40 + s); };
    auto p_empty = [&](int s) { return bar_base + 8u * (44 + s); };
    const uint32_t tmem_slot = bar_base + 8u * 38;
    // progress board for stall reports (static shared memory: its address costs no register): see mbar_wait.  Every wait of every
    // role leader then starts with a shared-memory store, through an LSU the producers keep busy: 0.7 % of the C2 frame rate
    // (609.1 / 609.6 fps with, 613.5 / 613.8 without, same box).  So the board is armed on request (DEFT_TC_DBG bit 32); a
    // stall still reports its own site and traps either way -- reproduce with the bit set to see where the other roles stood.
    __shared__ int prog_board[16];
    volatile int* const prog = (dbg & 32) ? prog_board : nullptr;
    unsigned char* epi_stage = smem + patch_bytes + (size_t)sb * Cfg::B_STAGE + 512;     // epilogue warps' transpose stages, scale, shift
    unsigned char* ap_common = epi_stage + TS_EPI_SMEM;                                  // producer warps' transpose stages + row tables
    unsigned char* ap_extra = ap_common + TS_COMMON_SMEM;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(a_full(s), 4);                      // the 4 quadrant warps of the producing k-group
            mbar_init(a_empty(s), 1);
            mbar_init(b_full(s, 0), 1);
            mbar_init(b_full(s, 1), 1);
            mbar_init(b_empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), Cfg::ISS);             // one commit (or plain arrive) per issuer warp
            mbar_init(acc_empty(b), TC_EWARPS);
        }
        for (int s = 0; s < TP_MAX_STAGES; ++s) {
            mbar_init(p_full(s), 1);
            mbar_init(p_empty(s), TC_PWARPS);
        }
        for (int q = 0; q < 16; ++q) prog_board[q] = 0;
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == TC_AUX_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + patch_bytes + (size_t)sb * Cfg::B_STAGE + 8 * 38);
    const int tiles_mn = ts.m_tiles * ts.n_tiles;
    pdl_launch_dependents();                              // the next launch may take SMs as our CTAs retire

    // Register rebalancing between the warpgroups (setmaxnreg): the kernel launches with 80 registers per thread
    // (768 threads x 80 = 60 K of the 64 K file), which the A producers overflow -- their spills go through an L1 that the
    // ~200 KB shared-memory carve-out leaves ~30 KB of, i.e. every spill reload is an L2 round trip in the hot loop.  The
    // issuer / copy warpgroup needs a few dozen registers and the epilogue warpgroup ~64: they release theirs, the four
    // producer warpgroups take them.  (Warpgroup = 4 consecutive warps; every warp of one executes the same setmaxnreg.)
#if DEFT_TS_SETMAXNREG
    if (warp < TC_PWARPS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DEFT_TS_PROD_REGS));
    else if (warp < TC_MMA_WARP) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DEFT_TS_EPI_REGS));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DEFT_TS_AUX_REGS));
#endif

    if (warp < TC_PWARPS) {
        // ===================== A producers: warp = (TMEM lane quadrant, k-group) =====================
        pdl_wait();                                       // activations (and the DCN offset map) come from the predecessor
        if constexpr (kPatch && PG::DCN) {
            // ----- DCN patch producer (3xFP16 only): thread = output pixel; bilinear corners out of the halo patch
            static_assert(!PG::DCN || Cfg::F16, "the DCN patch producer is written for 64-channel (3xFP16) k-blocks");
            const int quad = warp & 3, grp = warp >> 2;
            const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
            const int r = quad * 32 + lane, ry = r >> 4, rx = r & 15;
            const int H = ap.H, W = ap.W, pitch = ap.pitch, rowpitch = W * pitch;
            const int per_img = ap.tiles_x * ap.tiles_y;
            int gk0 = 0, gc0 = 0;
            for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
                const int z = t / tiles_mn, mn = t - z * tiles_mn;
                const int mt = mn / ts.n_tiles;
                const int n = mt / per_img, tr = mt - n * per_img;
                const int ty0 = (tr / ap.tiles_x) * TP_PH, tx0 = (tr % ap.tiles_x) * TP_PW;
                const int y = ty0 + ry, xx = tx0 + rx;
                const bool pix_ok = y < H && xx < W;
                const float* omrow = ap.om + ((size_t)(n * H + (pix_ok ? y : 0)) * W + (pix_ok ? xx : 0)) * 32;
                const int hy0 = ty0 - PG::HALO, hx0 = tx0 - PG::HALO;                   // image coordinates of halo element (0, 0)
                const int kb0 = z * ts.kb_per_split;
                const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;      // whole channel blocks: multiple of 9
                const int c64_0 = kb0 / 9;
                int last_c = -1;
                const float* pbase = nullptr;
                int i = (grp - (gk0 % TS_KGROUPS) + TS_KGROUPS) % TS_KGROUPS;
                for (; i < nkb; i += TS_KGROUPS) {
                    const int c = i / 9, tap = i - c * 9;
                    if (tid == grp * 128) TC_TRACE(1, gk0 + i, 0);
                    // sampling position of this pixel and tap (dcn_v2: zero outside (-1, H) x (-1, W), per-corner bounds)
                    float u00 = 0.f, u01 = 0.f, u10 = 0.f, u11 = 0.f;
                    int pidx = 0, dxs = 0, dys = 0;                                  // patch index of the top-left corner, strides to the others
                    const float* gp = nullptr;                                       // != nullptr: corners leave the halo, read them from global
                    int gdx = 0, gdy = 0;
                    if (pix_ok) {
                        const float dy = __ldg(omrow + 2 * tap), dx = __ldg(omrow + 2 * tap + 1);
                        const float mk = sigmoidf_(__ldg(omrow + 18 + tap));
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const float py = (float)(y - 1 + ky) + dy, px = (float)(xx - 1 + kx) + dx;
                        if (py > -1.f && py < (float)H && px > -1.f && px < (float)W) {
                            const float fy = floorf(py), fx = floorf(px);
                            const int y0 = (int)fy, x0 = (int)fx;
                            const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
                            const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= H - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= W - 1;
                            const float wy0 = y0ok ? hy : 0.f, wy1 = y1ok ? ly : 0.f, wx0 = x0ok ? hx * mk : 0.f, wx1 = x1ok ? lx * mk : 0.f;
                            u00 = wy0 * wx0; u01 = wy0 * wx1; u10 = wy1 * wx0; u11 = wy1 * wx1;
                            const int yc = y0ok ? y0 : 0, xc = x0ok ? x0 : 0;         // y0 = -1: the lower row IS row 0 (its weight sits in u1x)
                            const int ex = (x0ok && x1ok) ? 1 : 0, ey = (y0ok && y1ok) ? 1 : 0;
                            const int cy = yc - hy0, cx = xc - hx0;
                            if (cy >= 0 && cx >= 0 && cy + ey < AP::kHH && cx + ex < PG::HW) {
                                pidx = cy * PG::HW + cx; dxs = ex; dys = ey * PG::HW;
                            } else {
                                gp = ap.x + (size_t)((n * H + yc) * W + xc) * pitch;
                                gdx = ex ? pitch : 0; gdy = ey ? rowpitch : 0;
                            }
                        }
                    }
                    if (c != last_c) {
                        if (last_c >= 0) {                 // done with the previous channel block's patch
                            __syncwarp();
                            if (lane == 0) mbar_arrive(p_empty((gc0 + last_c) % ps));
                        }
                        const int gc = gc0 + c;
                        mbar_wait(p_full(gc % ps), (uint32_t)(gc / ps) & 1u, 1, gc, quad == 0 ? prog : nullptr, grp);
                        pbase = reinterpret_cast<const float*>(smem + (size_t)(gc % ps) * P_STAGE);
                        last_c = c;
                    }
                    const int gk = gk0 + i;
                    const int sa = gk % SA;
                    const uint32_t pa = (uint32_t)(gk / SA) & 1u;
                    const uint32_t acol = tmem_base + lane_addr + (uint32_t)(Cfg::A0 + sa * Cfg::ACOLS);
                    const int i00 = pidx, i01 = pidx + dxs, i10 = pidx + dys, i11 = pidx + dys + dxs;
                    const float* gx = gp != nullptr ? gp + (size_t)(c64_0 + c) * 64 : nullptr;
                    float4 x[8];
                    auto blend32 = [&](int h) {            // 32 channels (half h of the k-block) of this pixel's sample -> x[0..7]
                        if (gx == nullptr) {
                            const float* ph = pbase + h * (P_SUB / 4);
                            const float* q00 = ph + i00 * 32; const float* q01 = ph + i01 * 32;
                            const float* q10 = ph + i10 * 32; const float* q11 = ph + i11 * 32;
                            const int s00 = i00 & 7, s01 = i01 & 7, s10 = i10 & 7, s11 = i11 & 7;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 a0 = *reinterpret_cast<const float4*>(q00 + ((q ^ s00) << 2));
                                const float4 a1 = *reinterpret_cast<const float4*>(q01 + ((q ^ s01) << 2));
                                const float4 a2 = *reinterpret_cast<const float4*>(q10 + ((q ^ s10) << 2));
                                const float4 a3 = *reinterpret_cast<const float4*>(q11 + ((q ^ s11) << 2));
                                x[q].x = u00 * a0.x + u01 * a1.x + u10 * a2.x + u11 * a3.x; x[q].y = u00 * a0.y + u01 * a1.y + u10 * a2.y + u11 * a3.y;
                                x[q].z = u00 * a0.z + u01 * a1.z + u10 * a2.z + u11 * a3.z; x[q].w = u00 * a0.w + u01 * a1.w + u10 * a2.w + u11 * a3.w;
                            }
                        } else {
                            const float* g = gx + h * 32;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 a0 = ldg4(g + q * 4), a1 = ldg4(g + gdx + q * 4), a2 = ldg4(g + gdy + q * 4), a3 = ldg4(g + gdy + gdx + q * 4);
                                x[q].x = u00 * a0.x + u01 * a1.x + u10 * a2.x + u11 * a3.x; x[q].y = u00 * a0.y + u01 * a1.y + u10 * a2.y + u11 * a3.y;
                                x[q].z = u00 * a0.z + u01 * a1.z + u10 * a2.z + u11 * a3.z; x[q].w = u00 * a0.w + u01 * a1.w + u10 * a2.w + u11 * a3.w;
                            }
                        }
                    };
                    blend32(0);
                    if (tid == grp * 128) TC_TRACE(1, gk, 1);
                    mbar_wait(a_empty(sa), pa ^ 1u, 2, gk, quad == 0 ? prog : nullptr, grp);
                    __syncwarp();                              // converged before the .aligned tcgen05.st
                    tc_fence_after();
                    if (tid == grp * 128) TC_TRACE(1, gk, 2);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1) { blend32(1); __syncwarp(); }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float v[16] = {x[4 * q].x, x[4 * q].y, x[4 * q].z, x[4 * q].w, x[4 * q + 1].x, x[4 * q + 1].y, x[4 * q + 1].z, x[4 * q + 1].w,
                                                 x[4 * q + 2].x, x[4 * q + 2].y, x[4 * q + 2].z, x[4 * q + 2].w, x[4 * q + 3].x, x[4 * q + 3].y, x[4 * q + 3].z, x[4 * q + 3].w};
                            uint32_t hi[8], lo[8];
                            split_f16x3(v, hi, lo);
                            tmem_st8u(acol + (uint32_t)(h * 16 + q * 8), hi);
                            tmem_st8u(acol + 32u + (uint32_t)(h * 16 + q * 8), lo);
                        }
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full(sa));
                    if (tid == grp * 128) TC_TRACE(1, gk, 3);
                }
                if (last_c >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_empty((gc0 + last_c) % ps));
                }
                gk0 += nkb;
                gc0 += nkb / 9;
            }
        } else if constexpr (kPatch) {
            // ----- patch producer: thread = output pixel (ry, rx) of the 16x8 patch, operands from the TMA-filled halo patch
            const int quad = warp & 3, grp = warp >> 2;
            const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
            const int r = quad * 32 + lane, ry = r >> 4, rx = r & 15;
            int gk0 = 0, gc0 = 0;
            // p_empty expects all 16 producer warps per patch.  With three k-blocks per patch (C16) one of the four k-groups has no
            // k-block in it: its warps release the patch too -- after waiting for p_full, so that the arrival cannot land in the
            // phase of the stage's previous use (the loader issues use u+1 only once use u has been released by everyone).
            auto skip_patch = [&](int gc) {
                mbar_wait(p_full(gc % ps), (uint32_t)(gc / ps) & 1u, 1, gc, quad == 0 ? prog : nullptr, grp);
                __syncwarp();
                if (lane == 0) mbar_arrive(p_empty(gc % ps));
            };
            for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
                const int z = t / tiles_mn;
                const int kb0 = z * ts.kb_per_split;
                const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;      // whole channel blocks: multiple of 9
                int last_c = -1;
                const float* pbase = nullptr;
                int i = (grp - (gk0 % TS_KGROUPS) + TS_KGROUPS) % TS_KGROUPS;
                for (; i < nkb; i += TS_KGROUPS) {
                    const int c = i / KPP, tap = i - c * KPP;                 // C16: tap = k-block of the tile (taps 4 tap .. 4 tap + 3)
                    if (tid == grp * 128) TC_TRACE(1, gk0 + i, 0);
                    if (c != last_c) {
                        if (last_c >= 0) {                 // done with the previous channel block's patch
                            __syncwarp();
                            if (lane == 0) mbar_arrive(p_empty((gc0 + last_c) % ps));
                        }
                        for (int pc = last_c + 1; pc < c; ++pc) skip_patch(gc0 + pc);     // patches without a k-block of this group (KPP < 4)
                        const int gc = gc0 + c;
                        mbar_wait(p_full(gc % ps), (uint32_t)(gc / ps) & 1u, 1, gc, quad == 0 ? prog : nullptr, grp);
                        pbase = reinterpret_cast<const float*>(smem + (size_t)(gc % ps) * P_STAGE);
                        last_c = c;
                    }
                    const int gk = gk0 + i;
                    const int sa = gk % SA;
                    const uint32_t pa = (uint32_t)(gk / SA) & 1u;
                    const uint32_t acol = tmem_base + lane_addr + (uint32_t)(Cfg::A0 + sa * Cfg::ACOLS);
                    const int ky = tap / 3, kx = tap - ky * 3;
                    const int idx = (ry + ky) * TP_HW + rx + kx;
                    const float* prow = pbase + idx * 32;
                    const int sw = idx & 7;
                    float4 x[8];
                    // C16: half h of the k-block = taps 4 tap + 2 h, + 1: 2 x 64 B of the 16-channel patch (SWIZZLE_64B: 16-byte chunk ^ ((pixel >> 1) & 3))
                    auto load16 = [&](int h) {
#pragma unroll
                        for (int tq = 0; tq < 2; ++tq) {
                            const int t16 = 4 * tap + 2 * h + tq;
                            if (t16 < 9) {
                                const int ky16 = t16 / 3, kx16 = t16 - ky16 * 3;
                                const int i16 = (ry + ky16) * TP_HW + rx + kx16;
                                const float* p16 = pbase + i16 * 16;
                                const int s16 = (i16 >> 1) & 3;
#pragma unroll
                                for (int q = 0; q < 4; ++q) x[tq * 4 + q] = *reinterpret_cast<const float4*>(p16 + ((q ^ s16) << 2));
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; ++q) x[tq * 4 + q] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                    };
                    if constexpr (PG::C16) {
                        load16(0);
                    } else if (!(dbg & 2)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) x[q] = *reinterpret_cast<const float4*>(prow + ((q ^ sw) << 2));
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (tid == grp * 128) TC_TRACE(1, gk, 1);
                    mbar_wait(a_empty(sa), pa ^ 1u, 2, gk, quad == 0 ? prog : nullptr, grp);
                    __syncwarp();                              // converged before the .aligned tcgen05.st
                    tc_fence_after();
                    if (tid == grp * 128) TC_TRACE(1, gk, 2);
                    if constexpr (Cfg::F16) {
                        // 64 channels per k-block: the two 32-channel halo sub-patches of the stage, one after the other
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h == 1) {
                                if constexpr (PG::C16) {
                                    load16(1);
                                } else {
                                    const float* prow1 = prow + P_SUB / 4;
#pragma unroll
                                    for (int q = 0; q < 8; ++q) x[q] = *reinterpret_cast<const float4*>(prow1 + ((q ^ sw) << 2));
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const float v[16] = {x[4 * q].x, x[4 * q].y, x[4 * q].z, x[4 * q].w, x[4 * q + 1].x, x[4 * q + 1].y, x[4 * q + 1].z, x[4 * q + 1].w,
                                                     x[4 * q + 2].x, x[4 * q + 2].y, x[4 * q + 2].z, x[4 * q + 2].w, x[4 * q + 3].x, x[4 * q + 3].y, x[4 * q + 3].z, x[4 * q + 3].w};
                                uint32_t hi[8], lo[8];
                                split_f16x3(v, hi, lo);
                                tmem_st8u(acol + (uint32_t)(h * 16 + q * 8), hi);
                                tmem_st8u(acol + 32u + (uint32_t)(h * 16 + q * 8), lo);
                            }
                        }
                    } else if (!(dbg & 4)) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float v[8] = {x[2 * q].x, x[2 * q].y, x[2 * q].z, x[2 * q].w, x[2 * q + 1].x, x[2 * q + 1].y, x[2 * q + 1].z, x[2 * q + 1].w};
                            float hi[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) hi[j] = tf32_rna(v[j]);
                            tmem_st8(acol + (uint32_t)(q * 8), hi);
                            if (NPASS == 3) {
                                float lo[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) lo[j] = v[j] - hi[j];
                                tmem_st8(acol + 32u + (uint32_t)(q * 8), lo);
                            }
                        }
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full(sa));
                    if (tid == grp * 128) TC_TRACE(1, gk, 3);
                }
                if (last_c >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_empty((gc0 + last_c) % ps));
                }
                for (int pc = last_c + 1; pc < nkb / KPP; ++pc) skip_patch(gc0 + pc);
                gk0 += nkb;
                gc0 += nkb / KPP;
            }
        } else {
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        float* stage = reinterpret_cast<float*>(ap_common + (size_t)warp * TS_WARP_SMEM);
        int4* meta = reinterpret_cast<int4*>(ap_common + (size_t)warp * TS_WARP_SMEM + TS_STAGE_FLOATS * 4);
        ap.setup_once(ap_extra);
        int gk0 = 0;                                      // global k-block counter at the start of the tile
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int m0 = (mn / ts.n_tiles) * TC_BM;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            ap.tile_init(m0, M, ap_extra, quad, lane, meta);
            // this group's k-blocks: global index gk with gk % TS_KGROUPS == grp
            int i = (grp - (gk0 % TS_KGROUPS) + TS_KGROUPS) % TS_KGROUPS;
            for (; i < nkb; i += TS_KGROUPS) {
                const int gk = gk0 + i;
                const int sa = gk % SA;
                const uint32_t pa = (uint32_t)(gk / SA) & 1u;
                const uint32_t acol = tmem_base + lane_addr + (uint32_t)(Cfg::A0 + sa * Cfg::ACOLS);
                if (tid == grp * 128) TC_TRACE(1, gk, 0);
                if constexpr (Cfg::F16) {
                    // k-block = 64 channels = two 32-channel fills through the same private stage.  K order is (64-channel block,
                    // tap, channel); fill() counts 32-channel blocks: (2 c + h) * taps + tap.
                    const int taps = ap.ktaps();
                    const int kbg = kb0 + i, c64 = kbg / taps, tap = kbg - c64 * taps;
                    ap.fill((2 * c64) * taps + tap, quad, lane, meta, stage);
                    __syncwarp();
                    if (tid == grp * 128) TC_TRACE(1, gk, 1);
                    mbar_wait(a_empty(sa), pa ^ 1u, 2, gk, quad == 0 ? prog : nullptr, grp);
                    __syncwarp();                              // converged before the .aligned tcgen05.st
                    tc_fence_after();
                    if (tid == grp * 128) TC_TRACE(1, gk, 2);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1) {
                            __syncwarp();                      // everyone has read the first half out of the stage
                            ap.fill((2 * c64 + 1) * taps + tap, quad, lane, meta, stage);
                            __syncwarp();
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {          // lane = row: 16 columns at a time
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 t4 = *reinterpret_cast<const float4*>(stage + lane * TS_STAGE_LD + q * 16 + j * 4);
                                v[4 * j] = t4.x; v[4 * j + 1] = t4.y; v[4 * j + 2] = t4.z; v[4 * j + 3] = t4.w;
                            }
                            uint32_t hi[8], lo[8];
                            split_f16x3(v, hi, lo);
                            tmem_st8u(acol + (uint32_t)(h * 16 + q * 8), hi);
                            tmem_st8u(acol + 32u + (uint32_t)(h * 16 + q * 8), lo);
                        }
                    }
                } else {
                if (!(dbg & 2)) ap.fill(kb0 + i, quad, lane, meta, stage);     // coalesced loads -> private smem stage
                __syncwarp();
                if (tid == grp * 128) TC_TRACE(1, gk, 1);
                mbar_wait(a_empty(sa), pa ^ 1u, 2, gk, quad == 0 ? prog : nullptr, grp);
                    __syncwarp();                              // converged before the .aligned tcgen05.st
                tc_fence_after();
                if (tid == grp * 128) TC_TRACE(1, gk, 2);
                if (!(dbg & 4)) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {          // lane = row: 8 columns at a time
                        const float4 x0 = *reinterpret_cast<const float4*>(stage + lane * TS_STAGE_LD + q * 8);
                        const float4 x1 = *reinterpret_cast<const float4*>(stage + lane * TS_STAGE_LD + q * 8 + 4);
                        const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                        float hi[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hi[j] = tf32_rna(v[j]);
                        tmem_st8(acol + (uint32_t)(q * 8), hi);
                        if (NPASS == 3) {
                            float lo[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) lo[j] = v[j] - hi[j];
                            tmem_st8(acol + 32u + (uint32_t)(q * 8), lo);
                        }
                    }
                }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();                              // also: everyone has read the stage before the next fill
                if (lane == 0) mbar_arrive(a_full(sa));
                if (tid == grp * 128) TC_TRACE(1, gk, 3);
            }
            gk0 += nkb;
        }
        }  // !kPatch
    } else if (warp < TC_MMA_WARP) {
        // ===================== epilogue warps =====================
        const int ewarp = warp - TC_PWARPS;
        const int etid = ewarp * 32 + lane;
        // scale / shift of every output channel, once per CTA (defaults 1 / 0; columns >= Cout are never stored)
        float* ss = reinterpret_cast<float*>(epi_stage) + TC_EWARPS * 32 * TS_EPI_LD;
        for (int i = etid; i < TS_SS_MAXC; i += TC_EWARPS * 32) {
            ss[i] = (epi.e.scale != nullptr && i < epi.e.Cout) ? __ldg(epi.e.scale + i) : 1.f;
            ss[TS_SS_MAXC + i] = (epi.e.shift != nullptr && i < epi.e.Cout) ? __ldg(epi.e.shift + i) : 0.f;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        pdl_wait();                                       // residual reads, output / scratch writes
        int iter = 0, gk0 = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++iter) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int m0 = (mn / ts.n_tiles) * TC_BM, ntile = mn % ts.n_tiles;
            const int b = iter % Cfg::NBUF;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 0);
            mbar_wait(acc_full(b), (uint32_t)(iter / Cfg::NBUF) & 1u, 3, iter, ewarp == 0 ? prog : nullptr, 7);
            __syncwarp();                                  // converged before the .aligned tcgen05.ld
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 1);
            tc_fence_after();
            // Issuer w owns the k-blocks whose GLOBAL index is congruent to w (see the issuer loop), so with a single k-block
            // only the accumulator of issuer (gk0 % ISS) holds data; otherwise all of them do.
            const int first_iss = (Cfg::ISS > 1 && nkb == 1) ? (gk0 % Cfg::ISS) : 0;
            const int nblk = ((Cfg::ISS > 1 && nkb > 1) ? Cfg::ISS : 1) * Cfg::NBLK;
            // The tensor core adds every MMA into the fp32 accumulator with round-toward-zero (measured: tools/tc_bias.py, a
            // gain of -1.6e-8 per accumulate step, e.g. -6.8e-6 at K = 2304).  An accumulator of this tile took
            // 4 * NPASS * ceil(nkb / ISS) steps; the expected shrink is undone here (DEFT_TC_RZ_COMP=0 disables it).
            const float rz_gain = 1.0f + rz_comp * (float)((4 / Cfg::CH) * Cfg::PASSES * ((nkb + Cfg::ISS - 1) / Cfg::ISS));
            TsRowMap rm{};
            rm.M = M;
            if constexpr (kPatch) {
                const int mt = mn / ts.n_tiles, per_img = ap.tiles_x * ap.tiles_y;
                const int n = mt / per_img, tr = mt - n * per_img;
                rm.base = n * ap.H * ap.W; rm.H = ap.H; rm.W = ap.W;
                rm.y0 = (tr / ap.tiles_x) * TP_PH; rm.x0 = (tr % ap.tiles_x) * TP_PW;
            } else {
                rm.base = m0;                              // W = 0: rows are consecutive pixels
            }
            if (!(dbg & 16))
                ts_epilogue<BN>(epi, tmem_base + (uint32_t)((b * Cfg::ISS + first_iss) * Cfg::ACCW), nblk, rz_gain, ewarp, lane, rm, ntile, z,
                                reinterpret_cast<float*>(epi_stage) + ewarp * 32 * TS_EPI_LD, ss);
            gk0 += nkb;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(b));
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 2);
        }
    } else if (warp < TC_AUX_WARP) {
        // ===================== two MMA issuer warps (A from TMEM, B from smem), alternating k-blocks =====================
        // tcgen05.mma issue blocks the issuing thread for roughly the MMA's duration (measured ~46 cycles for 128x64x8).
        // Issuer w owns k-blocks w, w+2, ... of every tile and accumulator column block w of the tile's set.
        constexpr uint32_t idesc = Cfg::F16 ? make_idesc_f16(BN) : make_idesc(BN);
        const int w = warp - TC_MMA_WARP;
        int iter = 0, gk0 = 0;
        for (int t = (w < Cfg::ISS) ? (int)blockIdx.x : ts.total; t < ts.total; t += gridDim.x, ++iter) {
            const int z = t / tiles_mn;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            const int b = iter % Cfg::NBUF;
            const uint32_t tile_par = (uint32_t)(iter / Cfg::NBUF) & 1u;
            mbar_wait(acc_empty(b), tile_par ^ 1u, 4, iter, prog, 4 + w);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)((b * Cfg::ISS + w) * Cfg::ACCW);
            // Ownership by GLOBAL k-block index: issuer w takes gk with gk % ISS == w.  The A ring depth (4 or 6) and ISS (2) are
            // both even, so a ring stage is then consumed by the same issuer for the whole kernel.  r1 split by the tile-local
            // index; with an odd k-block count per tile (Cin = 32 3x3: 9) stage ownership flipped from tile to tile, an issuer
            // could run two uses of a stage ahead of the other one, and its parity wait on a_full was then satisfied by the
            // WRONG phase (stale operand, then a desynchronised ring: the C4 stall of round 2, the stream-K hang of round 1).
            const int i_first = (w - (gk0 % Cfg::ISS) + Cfg::ISS) % Cfg::ISS;
            for (int i = i_first; i < nkb; i += Cfg::ISS) {
                const int gk = gk0 + i;
                const int sa = gk % SA;
                const uint32_t pa = (uint32_t)(gk / SA) & 1u;
                const int s_b = gk % sb, u_b = gk / sb;                 // weight stage and its use index
                if (lane == 0 && w == 0) TC_TRACE(0, gk, 0);
                mbar_wait2(a_full(sa), pa, b_full(s_b, u_b), (uint32_t)(u_b >> 1) & 1u, 5, 6, gk, prog, 4 + w);
                if (lane == 0 && w == 0) { TC_TRACE(0, gk, 1); TC_TRACE(0, gk, 2); }
                tc_fence_after();
                const uint32_t a_hi = tmem_base + (uint32_t)(Cfg::A0 + sa * Cfg::ACOLS);
                const uint32_t a_lo = a_hi + 32u;
                const uint32_t b_hi = sbase + (uint32_t)s_b * Cfg::B_STAGE;
                const uint32_t b_lo = b_hi + Cfg::B_TILE;
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        if (dbg & 8) break;
                        const uint64_t db_hi = make_smem_desc(b_hi + k * 32);
                        if constexpr (Cfg::PAIR) {
                            // A_hi x [B_hi ; B_lo] as one N = 2*BN MMA (initialises / accumulates both column blocks),
                            // then the A_lo x B_hi correction into the first block
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_hi, make_idesc(2 * BN), (i > i_first || k > 0) ? 1u : 0u);
                            umma_tf32_ts(tmem_acc, a_lo + k * 8, db_hi, idesc, 1u);
                        } else if constexpr (Cfg::F16) {
                            const uint64_t db_lo = make_smem_desc(b_lo + k * 32);      // 16 fp16 = 32 bytes per k-step, 8 TMEM cells
                            const uint32_t d = tmem_acc + (uint32_t)((k % Cfg::CH) * BN);   // this k-step's chain
                            umma_f16_ts(d, a_lo + k * 8, db_hi, idesc, (i > i_first || k >= Cfg::CH) ? 1u : 0u);
                            umma_f16_ts(d, a_hi + k * 8, db_lo, idesc, 1u);
                            umma_f16_ts(d, a_hi + k * 8, db_hi, idesc, 1u);
                        } else if (NPASS == 3) {
                            const uint64_t db_lo = make_smem_desc(b_lo + k * 32);
                            umma_tf32_ts(tmem_acc, a_lo + k * 8, db_hi, idesc, (i > i_first || k > 0) ? 1u : 0u);
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_lo, idesc, 1u);
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_hi, idesc, 1u);
                        } else {
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_hi, idesc, (i > i_first || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(a_empty(sa));
                    umma_commit(b_empty(s_b));
                    if (i + Cfg::ISS >= nkb) umma_commit(acc_full(b));       // this issuer's share of the tile is in flight
                }
                __syncwarp();
                if (lane == 0 && w == 0) TC_TRACE(0, gk, 3);
            }
            if (i_first >= nkb && lane == 0) mbar_arrive(acc_full(b));    // no k-block for this issuer in the tile
            gk0 += nkb;
        }
    } else if (warp == TC_AUX_WARP) {
        // ===================== weight-tile bulk copies =====================
        int s_b = 0, gk = 0, u_b = 0;
        uint32_t ph_b = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int ntile = mn % ts.n_tiles;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(wblob) + ((size_t)ntile * ts.nkb_total + kb0) * (2 * Cfg::B_TILE);
            for (int i = 0; i < nkb; ++i, ++gk) {
                if (lane == 0) TC_TRACE(2, gk, 0);
                mbar_wait(b_empty(s_b), ph_b ^ 1u, 7, gk, prog, 6);
                if (lane == 0) TC_TRACE(2, gk, 1);
                const uint32_t dst = sbase + (uint32_t)s_b * Cfg::B_STAGE;
                if (elect_one_sync()) {
                    if (dbg & 1) { mbar_arrive(b_full(s_b, u_b)); }
                    else {
                        mbar_arrive_expect_tx(b_full(s_b, u_b), (uint32_t)Cfg::B_STAGE);
                        bulk_g2s(dst, src + (size_t)i * (2 * Cfg::B_TILE), (uint32_t)Cfg::B_STAGE, b_full(s_b, u_b));
                    }
                }
                __syncwarp();
                if (++s_b == sb) { s_b = 0; ph_b ^= 1u; ++u_b; }
            }
        }
    } else if (warp == TC_AUX_WARP + 1) {
        // ===================== halo-patch TMA loader (patch producer only) =====================
        if constexpr (kPatch) {
            pdl_wait();                                   // the TMA reads the predecessor's output
            int gc = 0;
            const int per_img = ap.tiles_x * ap.tiles_y;
            for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
                const int z = t / tiles_mn, mn = t - z * tiles_mn;
                const int mt = mn / ts.n_tiles;
                const int n = mt / per_img, tr = mt - n * per_img;
                const int y0 = (tr / ap.tiles_x) * TP_PH - PG::HALO, x0 = (tr % ap.tiles_x) * TP_PW - PG::HALO;
                const int kb0 = z * ts.kb_per_split;
                const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
                const int c0 = kb0 / KPP, nch = nkb / KPP;
                for (int c = 0; c < nch; ++c, ++gc) {
                    const int s = gc % ps;
                    mbar_wait(p_empty(s), ((uint32_t)(gc / ps) & 1u) ^ 1u, 8, gc, prog, 8);
                    if (elect_one_sync()) {
                        if constexpr (PG::C16) {           // the tile's one 16-channel patch
                            mbar_arrive_expect_tx(p_full(s), (uint32_t)PG::BOX);
                            tma_load_4d(raw_addr + pad + (uint32_t)s * P_STAGE, &ap_param.tmap, 0, x0, y0, n, p_full(s));
                        } else if constexpr (Cfg::F16) {   // 64 channels per k-block: two 32-channel boxes into the two halves of the stage
                            mbar_arrive_expect_tx(p_full(s), (uint32_t)(2 * PG::BOX));
                            tma_load_4d(raw_addr + pad + (uint32_t)s * P_STAGE, &ap_param.tmap, (c0 + c) * 64, x0, y0, n, p_full(s));
                            tma_load_4d(raw_addr + pad + (uint32_t)s * P_STAGE + P_SUB, &ap_param.tmap, (c0 + c) * 64 + 32, x0, y0, n, p_full(s));
                        } else {
                            mbar_arrive_expect_tx(p_full(s), (uint32_t)PG::BOX);
                            tma_load_4d(raw_addr + pad + (uint32_t)s * P_STAGE, &ap_param.tmap, (c0 + c) * TC_BK, x0, y0, n, p_full(s));
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (lane == 0) {                                       // role finished its tile loop (shows as site 0xee on the progress board)
        const int slot = warp < TC_PWARPS ? ((warp & 3) == 0 ? (warp >> 2) : -1) : warp < TC_MMA_WARP ? (warp == TC_PWARPS ? 7 : -1)
                         : warp < TC_AUX_WARP ? 4 + (warp - TC_MMA_WARP) : warp == TC_AUX_WARP ? 6 : 8;
        if (slot >= 0 && prog != nullptr) prog[slot] = 0xee << 24;
    }
    __syncthreads();
    if (warp == TC_AUX_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <class AP, int BN, int NPASS>
static int ts_launch(const AP& ap, const float* blob, int nkb, int M, int n_tiles, int splits, EpiStd epi, cudaStream_t st, int m_tiles_override = 0) {
    using Cfg = TsCfg<BN, NPASS>;
    constexpr int ps = AP::kPatch ? ((Cfg::F16 && !TsPG<AP>::C16) ? 2 : TP_MAX_STAGES) : 0;
    constexpr int P_STAGE = ((Cfg::F16 && !TsPG<AP>::C16) ? 2 : 1) * TsPG<AP>::STAGE;
    const int fixed = 1024 /*align*/ + 512 /*barriers*/ + TS_EPI_SMEM + (AP::kPatch ? ps * P_STAGE : TS_COMMON_SMEM) + AP::kExtraSmemBytes;
    TileSched ts{};
    ts.kb_per_split = ceil_div(nkb, splits);
    if (AP::kPatch) ts.kb_per_split = ceil_div(ts.kb_per_split, TsPG<AP>::KPP) * TsPG<AP>::KPP;      // whole channel blocks (patches) per split
    ts.splits = ceil_div(nkb, ts.kb_per_split);
    ts.m_tiles = m_tiles_override > 0 ? m_tiles_override : ceil_div(M, TC_BM);
    ts.n_tiles = n_tiles;
    ts.nkb_total = nkb;
    ts.total = ts.m_tiles * ts.n_tiles * ts.splits;
    int sb = (TS_SMEM_LIMIT - fixed) / Cfg::B_STAGE;
    if (sb > 8) sb = 8;
    if (std::is_same<AP, TsDcnAP>::value) {
        // The DCN gather goes through L1: claim only as much shared memory as the weight ring needs (DEFT_DCN_SB
        // stages, default 3) so that the unified L1/smem array keeps >= 60 KB of cache for the corner neighbourhoods.
        static int cap = -1;
        if (cap < 0) { const char* e = getenv("DEFT_DCN_SB"); cap = e ? atoi(e) : 3; if (cap < 2) cap = 2; }
        if (sb > cap) sb = cap;
    }
    if (sb < 2) return DEFT_EINVAL;
    size_t smem = (size_t)sb * Cfg::B_STAGE + fixed;
    static size_t granted[DEFT_MAX_DEVICES] = {};
    {
        const int rc_attr = ensure_dyn_smem(tc_ts_kernel<AP, BN, NPASS>, TS_SMEM_LIMIT, granted, "tc_ts_kernel");
        if (rc_attr != DEFT_OK) return rc_attr;
    }
    if (std::is_same<AP, TsDcnAP>::value) {
        // ask for exactly the shared memory this launch uses; the rest of the 256 KB array is the L1 the gather wants
        static int carve_done[DEFT_MAX_DEVICES] = {};
        int d = 0;
        if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < DEFT_MAX_DEVICES && carve_done[d] != (int)smem) {
            int pct = (int)((smem + 1024) * 100 / (228 * 1024)) + 1;
            if (pct > 100) pct = 100;
            cudaFuncSetAttribute(tc_ts_kernel<AP, BN, NPASS>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            carve_done[d] = (int)smem;
        }
    }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DEFT_TC_DBG"); dbg = e ? atoi(e) : 0; }
    long long* trace = debug_trace_ptr();
    const int grid = ts.total < TC_NUM_SMS ? ts.total : TC_NUM_SMS;
    static float rz = -1.f;
    if (rz < 0.f) { const char* e = getenv("DEFT_TC_RZ_COMP"); rz = (e && atoi(e) == 0) ? 0.f : 1.6e-8f; }
    if (launch_pdl(tc_ts_kernel<AP, BN, NPASS>, dim3(grid), dim3(TC_THREADS), smem, st, ap, blob, ts, M, sb, ps, epi, dbg, trace, rz) != cudaSuccess) {
        set_error("tc_ts_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return DEFT_ECUDA;
    }
    DEFT_CHECK_LAUNCH("tc_ts_kernel");
    return DEFT_OK;
}

// 1 = route narrow-N problems through the A-via-TMEM kernel (default); DEFT_TC_TS=0 keeps the smem-operand kernel.
static bool ts_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DEFT_TC_TS"); v = e ? atoi(e) : 1; }
    return v != 0;
}

}  // namespace deft
