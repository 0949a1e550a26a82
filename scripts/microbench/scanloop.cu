// Microbenchmark of the epilogue's inner loop in isolation (registers in, no TMEM): cycles per 32 columns
// per warp for 1 / 2 / 4 warps per SM sub-partition.  nvcc -arch=sm_100a -O3 -o scanloop scanloop.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define BIG 3.0e38f
__device__ __forceinline__ void merge_pair(float a, float b, float& m1, float& m2, float& m3, int one) {
    const float lo = fminf(a, b);
    const int t_ = __float_as_int(a) * one + __float_as_int(b);
    const float hi = __int_as_float(t_ - __float_as_int(lo) * one);
    const float n3 = fminf(fminf(m3, fmaxf(m2, lo)), fmaxf(m1, hi));
    const float n2 = fminf(fminf(m2, fmaxf(m1, lo)), hi);
    m1 = fminf(m1, lo);
    m2 = n2;
    m3 = n3;
}
__device__ __forceinline__ void scan16(const uint32_t (&v)[16], const float* nptr, float na, uint32_t colpack,
                                       float (&A)[3], float (&B)[3], int one) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 nn = *reinterpret_cast<const float4*>(nptr + j);
        const float s0 = fmaf(na, nn.x, __uint_as_float(v[j + 0]));
        const float s1 = fmaf(na, nn.y, __uint_as_float(v[j + 1]));
        const float s2 = fmaf(na, nn.z, __uint_as_float(v[j + 2]));
        const float s3 = fmaf(na, nn.w, __uint_as_float(v[j + 3]));
        const uint32_t cp = colpack + (uint32_t)j * 0x01010101u;
        const float p0 = __uint_as_float(__byte_perm(__float_as_uint(s0), cp, 0x3214));
        const float p1 = __uint_as_float(__byte_perm(__float_as_uint(s1), cp, 0x3215));
        const float p2 = __uint_as_float(__byte_perm(__float_as_uint(s2), cp, 0x3216));
        const float p3 = __uint_as_float(__byte_perm(__float_as_uint(s3), cp, 0x3217));
        merge_pair(p0, p1, A[0], A[1], A[2], one);
        merge_pair(p2, p3, B[0], B[1], B[2], one);
    }
}
__global__ void k(float* out, long long* cyc, const float* norms_g, int one, int iters) {
    __shared__ float norms[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) norms[i] = norms_g[i];
    __syncthreads();
    uint32_t va[16], vb[16];
    for (int i = 0; i < 16; ++i) { va[i] = __float_as_uint(1.f + threadIdx.x * 0.01f + i); vb[i] = __float_as_uint(2.f + threadIdx.x * 0.02f - i); }
    float A[3] = {BIG, BIG, BIG}, B[3] = {BIG, BIG, BIG};
    const float na = 0.5f + one;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t colpack = 0x03020100u;
#pragma unroll 1
        for (int cb = 0; cb < 256; cb += 32) {
            scan16(va, norms + cb, na, colpack, A, B, one);
            scan16(vb, norms + cb + 16, na, colpack + 0x10101010u, A, B, one);
            colpack += 0x20202020u;
            for (int i = 0; i < 16; ++i) { va[i] += 0x1234 * one; vb[i] ^= 0x77 * one; }  // keep inputs changing
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = A[0] + A[1] + A[2] + B[0] + B[1] + B[2];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float *out, *norms; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&norms, 1024); cudaMemset(norms, 0, 1024);
    for (int w : {1, 2, 3, 4, 6}) {
        const int iters = 64;
        k<<<148, w * 128>>>(out, cyc, norms, 1, iters); k<<<148, w * 128>>>(out, cyc, norms, 1, iters);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double per32 = (double)h[0] / (iters * 8.0);
        printf("warps/SMSP=%d cycles per 32 columns per warp = %.1f  -> SMSP throughput %.3f columns/cycle (scores/cycle/lane)\n", w, per32, 32.0 * w / per32);
    }
    return 0;
}
