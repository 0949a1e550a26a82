// Second pipe-rate microbenchmark: can alu-pipe and fma-pipe instructions issue together, and how fast are
// immediate forms?  nvcc -arch=sm_100a -O3 -o pipes2 pipes2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    float a[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; e[i] = seed * 3 + i; }
    float b = seed * 0.5f, c = seed * 0.25f;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { a[i] = fminf(a[i], b); e[i] = fmaf(e[i], b, c); }        // independent FMNMX + FFMA(3reg)
            if (OP == 1) a[i] = fmaf(a[i], 1.0009765625f, c);                        // FFMA imm multiplier
            if (OP == 2) a[i] = a[i] + 1.25f;                                        // FADD imm
            if (OP == 3) { a[i] = fminf(a[i], b); e[i] = e[i] + 1.25f; }             // independent FMNMX + FADD imm
            if (OP == 4) { a[i] = fminf(a[i], b); e[i] = __int_as_float(__float_as_int(e[i]) * 3 + 7); }  // FMNMX + IMAD
            if (OP == 5) a[i] = __int_as_float(__float_as_int(a[i]) * 3 + 7);        // IMAD imm
            if (OP == 6) { a[i] = fminf(a[i], 1.25f); }                              // FMNMX imm
            if (OP == 7) { a[i] = fminf(a[i], b); e[i] = fminf(e[i], c); }           // 2 independent FMNMX
        }
        b += 1e-7f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int w, int ops) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, w * 128>>>(out, cyc, 1.5f); k<OP><<<148, w * 128>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s warps/SMSP=%d cycles=%8lld warp-instr/cycle/SMSP=%.3f\n", name, w, h[0], (double)ITERS * ops * w / h[0]);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {2, 4, 8}) {
        run<0>("FMNMX || FFMA3reg", w, 16); run<1>("FFMA imm", w, 8); run<2>("FADD imm", w, 8);
        run<3>("FMNMX || FADDimm", w, 16); run<4>("FMNMX || IMAD", w, 16); run<5>("IMAD imm", w, 8);
        run<6>("FMNMX imm", w, 8); run<7>("FMNMX x2", w, 16);
    }
    return 0;
}
