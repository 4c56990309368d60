// tcgen05 "A through TMEM" kernel (TS-mode UMMA) for the narrow-N layers (N tile <= 128): included by tc_engine.cu.
//
// Measured on B200 (profiles/r1_tc_protocol_trace.md): with both operands in shared memory a 128x64x8 TF32 MMA
// takes ~70 cycles, because the 4 KB A tile is re-read from shared memory for every instruction -- the tensor pipe
// is operand-bandwidth bound, not math bound, whenever N is small.  Here the A tile never touches shared memory:
//   * producer thread = one accumulator row.  16 producer warps = 4 TMEM lane quadrants x 4 k-block groups; a warp
//     loads its 32 rows' 128 contiguous bytes per k-block (im2col tap / deformable gather / pair expansion), splits
//     them into tf32 hi + lo and writes them with tcgen05.st straight into TMEM columns (a ring of SA A-stages);
//   * the MMA warp issues tcgen05.mma with A from TMEM and only B (weights, bulk-copied) from shared memory;
//   * no generic->async proxy fence, no MEMBAR, no swizzle arithmetic, and shared memory holds only weights.
// Everything else (persistent tile loop, double-buffered accumulators, dedicated epilogue warps, 3xTF32 split,
// split-K partials) is as in tc_igemm_kernel.
#pragma once

namespace deft {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ------------------------------------------------------------------------------------------------ producers
// Interface (thread = row `row` of the tile, 0..127):
//   kExtraSmemBytes, BATCH (8-column groups fetched per load() call; 4/BATCH calls per k-block)
//   setup_once(extra)                      once per kernel, all 512 producer threads (may use named barrier 1)
//   tile_init(m0, M, extra, row)           once per tile
//   load(kb, bi, v[BATCH*8])               A[row, kb*32 + bi*BATCH*8 ...] (absolute k-block index kb)

// im2col over a channel concat of up to 4 NHWC sources, every channel count a multiple of 32.
// K order = (32-channel block, tap, channel) (tc_k_perm_taps) or 1x1.
struct TsSrcEntry { const float* p; int pitch; int coff; };
struct TsConvAP {
    ConvAP a;
    int taps, ncb;
    static constexpr int kExtraSmemBytes = 64 + 64 * 16;
    static constexpr int BATCH = 4;
    const int* toff_;
    const TsSrcEntry* cbsrc_;
    int pix_;
    unsigned vmask_;
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
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char*, int row) {
        const int p = m0 + row;
        vmask_ = 0;
        pix_ = 0;
        if (p < M) {
            const int n = p / (a.OH * a.OW);
            const int r = p - n * (a.OH * a.OW);
            const int oy = (r / a.OW) * a.stride, ox = (r - (r / a.OW) * a.OW) * a.stride;
            pix_ = (n * a.H + oy) * a.W + ox;
            unsigned bit = 1u;
            for (int ky = 0; ky < a.KH; ++ky)
                for (int kx = 0; kx < a.KW; ++kx, bit <<= 1) {
                    const int iy = oy - a.pad + ky, ix = ox - a.pad + kx;
                    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) vmask_ |= bit;
                }
        }
    }
    __device__ __forceinline__ void load(int kb, int, float (&v)[32]) const {
        const int cb = kb / taps, tap = kb - cb * taps;
        if ((vmask_ >> tap) & 1u) {
            const TsSrcEntry e = cbsrc_[cb];
            const float* ptr = e.p + (size_t)(pix_ + toff_[tap]) * e.pitch + e.coff;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = ldg4(ptr + 4 * j);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
    }
};

// DCNv2 bilinear gather; K order = (32-channel block, tap, channel).  The 4 corner element-offsets / weights of the
// tile's 128 rows x 9 taps live in smem (built once per tile by all producers, read by the row's 4 k-group threads).
struct TsDcnAP {
    DcnAP a;
    static constexpr int kExtraSmemBytes = TC_BM * 9 * 32;
    static constexpr int BATCH = 1;
    const int4* tabO_;
    const float4* tabW_;
    int row_;
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char* extra, int row) {
        int4* tabO = reinterpret_cast<int4*>(extra);
        float4* tabW = reinterpret_cast<float4*>(extra + TC_BM * 9 * 16);
        const int H = a.H, W = a.W;
        asm volatile("bar.sync 1, 512;" ::: "memory");     // every producer is done with the previous tile's table
        for (int idx = threadIdx.x; idx < TC_BM * 9; idx += TC_PRODUCERS) {
            int r9 = idx / 9, tap = idx - r9 * 9;
            int p = m0 + r9;
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
            tabO[idx] = o;       // element offsets; corners with zero weight point at element 0 (a valid address)
            tabW[idx] = w;
        }
        tabO_ = tabO;
        tabW_ = tabW;
        row_ = row;
        asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    __device__ __forceinline__ void load(int kb, int bi, float (&v)[8]) const {
        const int cb = kb / 9, tap = kb - cb * 9;
        const int4 o = tabO_[row_ * 9 + tap];
        const float4 w = tabW_[row_ * 9 + tap];
        const float* x = a.x + cb * TC_BK + bi * 8;
        const float4 c00 = ldg4(x + o.x), c01 = ldg4(x + o.x + 4);
        const float4 c10 = ldg4(x + o.y), c11 = ldg4(x + o.y + 4);
        const float4 c20 = ldg4(x + o.z), c21 = ldg4(x + o.z + 4);
        const float4 c30 = ldg4(x + o.w), c31 = ldg4(x + o.w + 4);
        v[0] = w.x * c00.x + w.y * c10.x + w.z * c20.x + w.w * c30.x;
        v[1] = w.x * c00.y + w.y * c10.y + w.z * c20.y + w.w * c30.y;
        v[2] = w.x * c00.z + w.y * c10.z + w.z * c20.z + w.w * c30.z;
        v[3] = w.x * c00.w + w.y * c10.w + w.z * c20.w + w.w * c30.w;
        v[4] = w.x * c01.x + w.y * c11.x + w.z * c21.x + w.w * c31.x;
        v[5] = w.x * c01.y + w.y * c11.y + w.z * c21.y + w.w * c31.y;
        v[6] = w.x * c01.z + w.y * c11.z + w.z * c21.z + w.w * c31.z;
        v[7] = w.x * c01.w + w.y * c11.w + w.z * c21.w + w.w * c31.w;
    }
};

// Affinity MLP layer 2: A[p, k] = relu(Au'[f(p), i(p), k] + Bv'[j(p), k]) over the live pairs.
struct TsPairAP {
    PairAP a;
    static constexpr int kExtraSmemBytes = 0;
    static constexpr int BATCH = 2;
    const float* pa_;
    const float* pb_;
    __device__ __forceinline__ void setup_once(unsigned char*) {}
    __device__ __forceinline__ void tile_init(int m0, int M, unsigned char*, int row) {
        const int p = m0 + row;
        pa_ = nullptr;
        pb_ = nullptr;
        if (p < M) {
            int f = 0;
            while (f < a.F - 1 && p >= a.off[f + 1]) ++f;
            int q = p - a.off[f];
            int ii = q / a.m, jj = q - ii * a.m;
            pa_ = a.au + ((size_t)f * a.P + ii) * a.ld;
            pb_ = a.bv + (size_t)jj * a.ld;
        }
    }
    __device__ __forceinline__ void load(int kb, int bi, float (&v)[16]) const {
        if (pa_ != nullptr) {
            const int k = kb * TC_BK + bi * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 x = ldg4(pa_ + k + 4 * j), y = ldg4(pb_ + k + 4 * j);
                v[4 * j] = fmaxf(x.x + y.x, 0.f); v[4 * j + 1] = fmaxf(x.y + y.y, 0.f);
                v[4 * j + 2] = fmaxf(x.z + y.z, 0.f); v[4 * j + 3] = fmaxf(x.w + y.w, 0.f);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
    }
};

// ------------------------------------------------------------------------------------------------ kernel
constexpr int TS_KGROUPS = TC_PWARPS / 4;       // k-block interleave across producer warp groups

template <int BN, int NPASS>
struct TsCfg {
    static constexpr int ACOLS = NPASS == 3 ? 64 : 32;                 // TMEM columns of one A stage (hi [+ lo])
    static constexpr int A0 = 2 * BN;                                   // first A column (after the 2 accumulators)
    static constexpr int SA_RAW = (512 - A0) / ACOLS;
    static constexpr int SA = SA_RAW > 6 ? 6 : SA_RAW;                  // A ring depth
    static constexpr int B_TILE = BN * 128;
    static constexpr int B_STAGE = NPASS == 3 ? 2 * B_TILE : B_TILE;
};

template <class AP, int BN, int NPASS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_ts_kernel(AP ap, const float* __restrict__ wblob, TileSched ts, int M, int sb, EpiStd epi,
                                                              int dbg, long long* trace) {
    using Cfg = TsCfg<BN, NPASS>;
    constexpr int SA = Cfg::SA;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    unsigned char* smem = smem_raw + pad;
    const uint32_t sbase = raw_addr + pad;
    // layout: [sb x B_STAGE] [barriers + tmem slot: 512 B] [AP extra]
    const uint32_t bar_base = sbase + (uint32_t)sb * Cfg::B_STAGE;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (8 + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (16 + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (24 + s); };
    auto acc_full = [&](int b) { return bar_base + 8u * (32 + b); };
    auto acc_empty = [&](int b) { return bar_base + 8u * (34 + b); };
    auto tile_go = [&](int b) { return bar_base + 8u * (36 + b); };
    const uint32_t tmem_slot = bar_base + 8u * 38;
    unsigned char* ap_extra = smem + (size_t)sb * Cfg::B_STAGE + 512;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(a_full(s), 4);                      // the 4 quadrant warps of the producing k-group
            mbar_init(a_empty(s), 1);
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), TC_MMA_WARPS);         // one commit (or plain arrive) per issuer warp
            mbar_init(acc_empty(b), TC_EWARPS);
            mbar_init(tile_go(b), 1);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == TC_AUX_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (size_t)sb * Cfg::B_STAGE + 8 * 38);
    const int tiles_mn = ts.m_tiles * ts.n_tiles;

    if (warp < TC_PWARPS) {
        // ===================== A producers: thread = row, warp = (quadrant, k-group) =====================
        const int quad = warp & 3, grp = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        ap.setup_once(ap_extra);
        int gk0 = 0;                                      // global k-block counter at the start of the tile
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int m0 = (mn / ts.n_tiles) * TC_BM;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            ap.tile_init(m0, M, ap_extra, row);
            // this group's k-blocks: global index gk with gk % TS_KGROUPS == grp
            int i = (grp - (gk0 % TS_KGROUPS) + TS_KGROUPS) % TS_KGROUPS;
            for (; i < nkb; i += TS_KGROUPS) {
                const int gk = gk0 + i;
                const int sa = gk % SA;
                const uint32_t pa = (uint32_t)(gk / SA) & 1u;
                const uint32_t acol = tmem_base + lane_addr + (uint32_t)(Cfg::A0 + sa * Cfg::ACOLS);
                if (tid == grp * 128) TC_TRACE(1, gk, 0);
                bool waited = false;
#pragma unroll
                for (int bi = 0; bi < 4 / AP::BATCH; ++bi) {
                    float v[AP::BATCH * 8];
                    if (!(dbg & 2)) ap.load(kb0 + i, bi, v);
                    if (!waited) {                         // the first loads are in flight while we wait for the slot
                        if (tid == grp * 128) TC_TRACE(1, gk, 1);
                        mbar_wait(a_empty(sa), pa ^ 1u);
                        tc_fence_after();
                        if (tid == grp * 128) TC_TRACE(1, gk, 2);
                        waited = true;
                    }
                    if (dbg & 4) continue;
#pragma unroll
                    for (int q = 0; q < AP::BATCH; ++q) {
                        float hi[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hi[j] = tf32_rna(v[q * 8 + j]);
                        tmem_st8(acol + (uint32_t)((bi * AP::BATCH + q) * 8), hi);
                        if (NPASS == 3) {
                            float lo[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) lo[j] = v[q * 8 + j] - hi[j];
                            tmem_st8(acol + 32u + (uint32_t)((bi * AP::BATCH + q) * 8), lo);
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(sa));
                if (tid == grp * 128) TC_TRACE(1, gk, 3);
            }
            gk0 += nkb;
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
            if (!(dbg & 16)) epi.template run<BN>(tmem_base + (uint32_t)(b * BN), nullptr, ewarp, lane, m0, M, ntile, z);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(b));
            if (ewarp == 0 && lane == 0) TC_TRACE(3, iter, 2);
        }
    } else if (warp < TC_AUX_WARP) {
        // ===================== two MMA issuer warps (A from TMEM, B from smem), alternating k-blocks =====================
        // tcgen05.mma issue blocks the issuing thread for roughly the MMA's duration (measured ~46 cycles for 128x64x8),
        // so a single issuer leaves the tensor pipe idle during its own barrier waits / commits; two issuers hide that.
        // Accumulation order between k-blocks is irrelevant except for the zero-initialising first MMA of a tile, which
        // issuer 0 owns; issuer 1 waits on tile_go until that k-block has completed.
        constexpr uint32_t idesc = make_idesc(BN);
        const int w = warp - TC_MMA_WARP;
        int iter = 0, gk0 = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x, ++iter) {
            const int z = t / tiles_mn;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            const int b = iter & 1;
            const uint32_t tile_par = (uint32_t)(iter >> 1) & 1u;
            mbar_wait(acc_empty(b), tile_par ^ 1u);
            if (w == 1 && nkb > 1) mbar_wait(tile_go(b), tile_par);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(b * BN);
            for (int i = w; i < nkb; i += TC_MMA_WARPS) {
                const int gk = gk0 + i;
                const int sa = gk % SA;
                const uint32_t pa = (uint32_t)(gk / SA) & 1u;
                const int s_b = gk % sb;
                const uint32_t ph_b = (uint32_t)(gk / sb) & 1u;
                if (lane == 0 && w == 0) TC_TRACE(0, gk, 0);
                mbar_wait(a_full(sa), pa);
                if (lane == 0 && w == 0) TC_TRACE(0, gk, 1);
                mbar_wait(b_full(s_b), ph_b);
                if (lane == 0 && w == 0) TC_TRACE(0, gk, 2);
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
                        if (NPASS == 3) {
                            const uint64_t db_lo = make_smem_desc(b_lo + k * 32);
                            umma_tf32_ts(tmem_acc, a_lo + k * 8, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_lo, idesc, 1u);
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_hi, idesc, 1u);
                        } else {
                            umma_tf32_ts(tmem_acc, a_hi + k * 8, db_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(a_empty(sa));
                    umma_commit(b_empty(s_b));
                    if (i == 0) umma_commit(tile_go(b));                 // issuer 1 may start accumulating
                    if (i + TC_MMA_WARPS >= nkb) umma_commit(acc_full(b));   // this issuer's share of the tile is in flight
                }
                __syncwarp();
                if (lane == 0 && w == 0) TC_TRACE(0, gk, 3);
            }
            if (w == 1 && nkb < 2 && lane == 0) mbar_arrive(acc_full(b));   // no k-block for issuer 1 in this tile
            gk0 += nkb;
        }
    } else if (warp == TC_AUX_WARP) {
        // ===================== weight-tile bulk copies =====================
        int s_b = 0, gk = 0;
        uint32_t ph_b = 0;
        for (int t = blockIdx.x; t < ts.total; t += gridDim.x) {
            const int z = t / tiles_mn, mn = t - z * tiles_mn;
            const int ntile = mn % ts.n_tiles;
            const int kb0 = z * ts.kb_per_split;
            const int nkb = min(ts.nkb_total, kb0 + ts.kb_per_split) - kb0;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(wblob) + ((size_t)ntile * ts.nkb_total + kb0) * (2 * Cfg::B_TILE);
            for (int i = 0; i < nkb; ++i, ++gk) {
                if (lane == 0) TC_TRACE(2, gk, 0);
                mbar_wait(b_empty(s_b), ph_b ^ 1u);
                if (lane == 0) TC_TRACE(2, gk, 1);
                const uint32_t dst = sbase + (uint32_t)s_b * Cfg::B_STAGE;
                if (elect_one_sync()) {
                    if (dbg & 1) { mbar_arrive(b_full(s_b)); }
                    else {
                        mbar_arrive_expect_tx(b_full(s_b), (uint32_t)Cfg::B_STAGE);
                        bulk_g2s(dst, src + (size_t)i * (2 * Cfg::B_TILE), (uint32_t)Cfg::B_STAGE, b_full(s_b));
                    }
                }
                __syncwarp();
                if (++s_b == sb) { s_b = 0; ph_b ^= 1u; }
            }
        }
    }
    __syncthreads();
    if (warp == TC_AUX_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <class AP, int BN, int NPASS>
static int ts_launch(AP ap, const float* blob, int nkb, int M, int n_tiles, int splits, EpiStd epi, cudaStream_t st) {
    using Cfg = TsCfg<BN, NPASS>;
    const int fixed = 1024 /*align*/ + 512 /*barriers*/ + AP::kExtraSmemBytes;
    TileSched ts{};
    ts.kb_per_split = ceil_div(nkb, splits);
    ts.splits = ceil_div(nkb, ts.kb_per_split);
    ts.m_tiles = ceil_div(M, TC_BM);
    ts.n_tiles = n_tiles;
    ts.nkb_total = nkb;
    ts.total = ts.m_tiles * ts.n_tiles * ts.splits;
    int sb = (TC_SMEM_LIMIT - fixed) / Cfg::B_STAGE;
    if (sb > 8) sb = 8;
    if (sb < 2) return DEFT_EINVAL;
    size_t smem = (size_t)sb * Cfg::B_STAGE + fixed;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc_ts_kernel<AP, BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT) != cudaSuccess) {
            set_error("tc_engine: cannot raise dynamic smem limit");
            return DEFT_ECUDA;
        }
        attr_done = true;
    }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DEFT_TC_DBG"); dbg = e ? atoi(e) : 0; }
    long long* trace = nullptr;
    if (const char* e = getenv("DEFT_TC_TRACE")) trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 10));
    const int grid = ts.total < TC_NUM_SMS ? ts.total : TC_NUM_SMS;
    tc_ts_kernel<AP, BN, NPASS><<<grid, TC_THREADS, smem, st>>>(ap, blob, ts, M, sb, epi, dbg, trace);
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
