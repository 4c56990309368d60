// Micro-experiment (GPU box; built by tools/exp_f16ts.py): tcgen05.mma kind::f16 with the A operand in TMEM (TS mode).
// Question 1: how are 16-bit A elements laid out in TMEM columns -- which half of a 32-bit cell holds the even k?
// Question 2: issue cost of kind::f16 (K = 16 per instruction) against kind::tf32 (K = 8) at N = 64 / 128.
// One CTA, 128 threads.  A [128 x 64] fp16 -> 32 TMEM columns via tcgen05.st (packing selected by `pack_mode`),
// B [N x 64] fp16 in smem, K-major, SWIZZLE_128B (128-byte rows), D [128 x N] fp32 -> global.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16: c_format F32 (1 << 4), a_format / b_format F16 = 0; kind::tf32: formats = 2
__device__ __forceinline__ uint32_t make_idesc(int n, int fmt) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

// mode 0: correctness (one k-block of 64 = 4 MMAs of K = 16); mode 1/2: timing loops (f16 / tf32), `iters` MMAs
__global__ void __launch_bounds__(128, 1) exp_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int N, int pack_mode,
                                                     int mode, int iters, long long* cycles) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    unsigned char* smem = smem_raw + pad;
    const uint32_t sb = raw + pad;                       // B tile: N rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // B -> smem, swizzled: element (n, k) at n*128 + ((k/8 ^ (n&7)) << 4) + (k%8)*2
    for (int i = tid; i < N * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        *reinterpret_cast<__half*>(smem + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    // A row = this thread (TMEM lane = tid); 64 halves -> 32 cells at columns 256..287
    {
        uint32_t cell[32];
        for (int j = 0; j < 32; ++j) {
            const unsigned short lo = __half_as_ushort(A[tid * 64 + 2 * j]), hi = __half_as_ushort(A[tid * 64 + 2 * j + 1]);
            cell[j] = pack_mode == 0 ? ((uint32_t)hi << 16) | lo : ((uint32_t)lo << 16) | hi;
        }
        for (int q = 0; q < 4; ++q)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_addr + 256u + q * 8), "r"(cell[q * 8]),
                         "r"(cell[q * 8 + 1]), "r"(cell[q * 8 + 2]), "r"(cell[q * 8 + 3]), "r"(cell[q * 8 + 4]), "r"(cell[q * 8 + 5]), "r"(cell[q * 8 + 6]),
                         "r"(cell[q * 8 + 7])
                         : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = make_idesc(N, mode == 2 ? 2 : 0);
        long long t0 = clock64();
        if (mode == 0) {
            for (int k = 0; k < 4; ++k) mma_f16_ts(tmem, tmem + 256u + k * 8, make_smem_desc(sb + k * 32), idesc, k > 0);
        } else {
            for (int it = 0; it < iters; ++it) {
                const int k = it & 3;
                if (mode == 1) mma_f16_ts(tmem, tmem + 256u + k * 8, make_smem_desc(sb + k * 32), idesc, 1);
                else mma_tf32_ts(tmem, tmem + 256u + k * 8, make_smem_desc(sb + k * 32), idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (cycles) cycles[0] = t1 - t0;
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            if (clock64() - t0 > 2000000000LL) { if (tid == 0) printf("exp_f16ts: MMA never completed\n"); __trap(); }
        }
        if (tid == 0 && cycles) cycles[1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0) {
        for (int c = 0; c < N; c += 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(tmem + lane_addr + c)
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

extern "C" int exp_f16ts_run(const void* A, const void* B, float* D, int N, int pack_mode, int mode, int iters, long long* cycles) {
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    exp_kernel<<<1, 128, 40 * 1024>>>((const __half*)A, (const __half*)B, D, N, pack_mode, mode, iters, cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("exp_f16ts: %s\n", cudaGetErrorString(e)); return -1; }
    return 0;
}
