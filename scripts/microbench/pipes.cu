// Pipe-rate microbenchmark for the epilogue design: warp-instructions per cycle per SM sub-partition
// for the ops the argmin epilogue is made of.  nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    float b = seed * 0.5f, c = seed * 0.25f;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fminf(a[i], b);                                      // FMNMX
            if (OP == 1) a[i] = fminf(fminf(a[i], b), c);                            // FMNMX3
            if (OP == 2) a[i] = __uint_as_float(__byte_perm(__float_as_uint(a[i]), __float_as_uint(b), 0x3214));  // PRMT
            if (OP == 3) a[i] = fmaf(a[i], b, c);                                    // FFMA
            if (OP == 4) a[i] = __uint_as_float((__float_as_uint(a[i]) & 0xFFFFFF00u) | 7u);  // LOP3
            if (OP == 5) a[i] = __int_as_float(max(__float_as_int(a[i]), __float_as_int(b)));  // VIMNMX
            if (OP == 6) a[i] = a[i] + b;                                            // FADD
            if (OP == 7) { a[i] = fminf(a[i], b); a[i] = fmaf(a[i], b, c); }          // FMNMX + FFMA (dual pipe)
            if (OP == 8) a[i] = __int_as_float(__viaddmin_s32(__float_as_int(a[i]), 3, __float_as_int(b)));  // VIADDMNMX
        }
        b += 1e-7f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// the epilogue's merge network: 8 FMNMX-class per pair
__global__ void kmerge(float* out, long long* cyc, float seed, int nchain) {
    float m[4][3];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) m[i][j] = 1e30f;
    float p = seed + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ch = i % nchain;
            const float x = p + i, y = p * 0.5f - i;
            const float lo = fminf(x, y), hi = fmaxf(x, y);
            const float n3 = fminf(fminf(m[ch][2], fmaxf(m[ch][1], lo)), fmaxf(m[ch][0], hi));
            const float n2 = fminf(fminf(m[ch][1], fmaxf(m[ch][0], lo)), hi);
            m[ch][0] = fminf(m[ch][0], lo);
            m[ch][1] = n2;
            m[ch][2] = n3;
        }
        p += 0.37f;
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) s += m[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps_per_smsp, int ops_per_iter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int threads = warps_per_smsp * 4 * 32;
    k<OP><<<148, threads>>>(out, cyc, 1.5f); k<OP><<<148, threads>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = h[0];
    double inst = (double)ITERS * ops_per_iter * warps_per_smsp;   // warp-instr per SMSP
    printf("%-22s warps/SMSP=%d  cycles=%8.0f  warp-instr/cycle/SMSP=%.3f\n", name, warps_per_smsp, c, inst / c);
    cudaFree(out); cudaFree(cyc);
}
void runmerge(int warps_per_smsp, int nchain) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int threads = warps_per_smsp * 4 * 32;
    kmerge<<<148, threads>>>(out, cyc, 1.5f, nchain); kmerge<<<148, threads>>>(out, cyc, 1.5f, nchain);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = h[0];
    printf("merge_pair chains=%d       warps/SMSP=%d  cycles=%8.0f  cycles/pair/warp=%.2f  pairs/cycle/SMSP=%.4f\n", nchain,
           warps_per_smsp, c, c / (ITERS * 4.0), ITERS * 4.0 * warps_per_smsp / c);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 2, 4}) {
        run<0>("FMNMX", w, 8); run<1>("FMNMX3", w, 8); run<2>("PRMT", w, 8); run<3>("FFMA", w, 8);
        run<4>("LOP3(and|or)", w, 8); run<5>("VIMNMX", w, 8); run<6>("FADD", w, 8); run<7>("FMNMX+FFMA", w, 16);
        run<8>("VIADDMNMX", w, 8);
    }
    for (int w : {1, 2, 4}) for (int ch : {1, 2, 4}) runmerge(w, ch);
    return 0;
}
