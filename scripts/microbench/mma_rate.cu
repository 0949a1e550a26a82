// mma_rate.cu -- how long does a tcgen05.mma kind::f16 M=128 x N x K=16 take, with both operands in shared memory
// (SS) or A in tensor memory (TS), alone and while other warps stream writes into shared memory?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

template <int N, bool TS>
__global__ void __launch_bounds__(256, 1) k(int iters, int bg, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tbase;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    uint8_t* A = smem;                 // 128 rows x 64 fp16 (16 KiB)
    uint8_t* B = smem + 16384;         // N rows x 64 fp16
    uint8_t* bgbuf = smem + 16384 + 32768;  // 64 KiB for the background writers
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(bgbuf) = 0u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tbase;
    if (warp == 0) {
        if (threadIdx.x == 0) {
            const uint64_t ad = desc_sw128(smem_u32(A)), bd = desc_sw128(smem_u32(B));
            const uint32_t id = idesc_f16(128, N);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (TS)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(t + (uint32_t)((i & 1) * N)),
                                     "r"(t + 256 + 8 * j), "l"(bd + 2 * j), "r"(id), "r"(1)
                                     : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(t + (uint32_t)((i & 1) * N)),
                                     "l"(ad + 2 * j), "l"(bd + 2 * j), "r"(id), "r"(1)
                                     : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
            *reinterpret_cast<volatile uint32_t*>(bgbuf) = 1u;  // stop flag for the writers
        }
    } else if (warp <= bg) {
        // background shared-memory traffic: each warp stores 512 bytes per instruction into its own 8 KiB region
        uint4* dst = reinterpret_cast<uint4*>(bgbuf + 1024 + (warp - 1) * 8192) + (threadIdx.x & 31);
        uint4 v = make_uint4(warp, 2, 3, 4);
        int n = 0;
        while (*reinterpret_cast<volatile uint32_t*>(bgbuf) == 0u && n < (1 << 22)) {
#pragma unroll
            for (int u = 0; u < 16; ++u) dst[u * 32] = v;
            ++n;
        }
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[warp] = (unsigned long long)n * 16 * 512;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
}

template <int N, bool TS>
void run(const char* name, unsigned long long* out) {
    const int iters = 8192, smem = 16384 + 32768 + 1024 + 7 * 8192;
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int bg = 0; bg <= 7; bg += (bg == 0 ? 1 : (bg == 1 ? 1 : (bg == 2 ? 2 : 3)))) {
        cudaMemset(out, 0, 64);
        k<N, TS><<<148, 256, smem>>>(iters, bg, out);
        k<N, TS><<<148, 256, smem>>>(iters, bg, out);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[8];
        cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
        unsigned long long wb = 0;
        for (int i = 1; i <= bg; ++i) wb += h[i];
        printf("%-28s background writer warps=%d: %.1f cycles per MMA (ideal %d), writers %.1f B/clk  %s\n", name, bg,
               (double)h[0] / (iters * 4), N / 2, (double)wb / h[0], cudaGetErrorString(e));
    }
}

int main() {
    unsigned long long* out;
    cudaMalloc(&out, 64);
    run<128, false>("SS M=128 N=128 K=16", out);
    run<64, false>("SS M=128 N=64  K=16", out);
    run<256, false>("SS M=128 N=256 K=16", out);
    run<128, true>("TS M=128 N=128 K=16", out);
    run<64, true>("TS M=128 N=64  K=16", out);
    return 0;
}
