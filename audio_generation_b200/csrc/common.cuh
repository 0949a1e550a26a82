// common.cuh -- constants, error plumbing and small device helpers shared by the RVQ kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/rvq_sm100a.h"

namespace rvq {

constexpr int TILE_M = 128;     // frames per tile (UMMA M)
constexpr int CHUNK_N = 256;    // codes per MMA / TMEM accumulator buffer (UMMA N)
constexpr int KSLICE = 64;      // fp16 elements per 128-byte swizzle row (one TMA box column span)
constexpr int META_STRIDE = 8;  // floats of per-stage metadata
constexpr int MAX_D = 512;

// fp16 operand windows: 2^b * max|c| in [2^8, 2^9) (so |-2 * 2^b c| < 2^10), 2^a * max|r| in [2^8, 2^9),
// a <= b + 3 so that 2^(a-b) * (2^(2b) ||c||^2) stays far below fp32 overflow.
constexpr int SCALE_TARGET_EXP = 8;
constexpr int SCALE_EXP_CLAMP = 60;
constexpr int ROW_OVER_CODE_MAX = 3;
constexpr int ROW_UNDER_CODE_MAX = 40;
constexpr float PAD_NORM = 1.2676506e30f;  // 2^100: padding codes can never be selected

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Layout of the cb_norm buffer written by rvq_prepare_codebooks for `nq` prepared stages (floats unless noted):
//   norm [nq, Kpad] | norm slices [nq, Kpad / 128, 4096 bytes] | xc [nq, Kpad] | xb bytes [nq, Kpad] | xflag int [nq, Kpad / 256]
// Allowance X_k = rs xc_k + na x2_k of the codes above the stage's norm cap (k0_bound in rvq_aux.cu): xc = its first
// factor (the second is folded into the norm), xb = one byte per code with X_k <= xb_k (rs U1 + na U2), U in the
// stage's meta[7], meta[5]; xflag = the 256-code chunk holds such a code.
struct NormLayout {
    uint8_t *slices, *xb;
    float* xc;
    int* xflag;
    __host__ __device__ NormLayout(const float* base, int nq, int Kpad) {
        float* b = const_cast<float*>(base);
        slices = reinterpret_cast<uint8_t*>(b + (size_t)nq * Kpad);
        xc = b + (size_t)nq * Kpad * 9;
        xb = reinterpret_cast<uint8_t*>(b + (size_t)nq * Kpad * 10);
        xflag = reinterpret_cast<int*>(xb + (size_t)nq * Kpad);
    }
    static size_t bytes(int nq, int Kpad) {
        return (size_t)nq * Kpad * 41 + (((size_t)nq * (Kpad / 256) * sizeof(int) + 15) & ~(size_t)15);
    }
};

// host-side error state (thread local)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RVQ_CUDA(call)                                        \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);   \
    } while (0)

// exponent e such that 2^e <= v < 2^(e+1) for finite v > 0
__device__ __forceinline__ int ilog2f_floor(float v) {
    // exponent field of a positive finite float (subnormals report -127: they are treated as tiny)
    return ((__float_as_int(v) >> 23) & 0xFF) - 127;
}

// exact power of two as float for |e| <= 126
__device__ __forceinline__ float exp2i(int e) { return __int_as_float((e + 127) << 23); }

// per-stage operand scale exponent b from max|c|
__device__ __forceinline__ int code_scale_exp(float cmax) {
    if (!(cmax > 0.f) || !isfinite(cmax)) return 0;
    int b = SCALE_TARGET_EXP - ilog2f_floor(cmax);
    return max(-SCALE_EXP_CLAMP, min(SCALE_EXP_CLAMP, b));
}
// per-row operand scale exponent a from max|r| and the stage's b
__device__ __forceinline__ int row_scale_exp(float amax, int b) {
    int a = b + ROW_OVER_CODE_MAX;
    if (amax > 0.f && isfinite(amax)) a = min(a, SCALE_TARGET_EXP - ilog2f_floor(amax));
    a = max(a, b - ROW_UNDER_CODE_MAX);
    return max(-SCALE_EXP_CLAMP - ROW_UNDER_CODE_MAX, min(SCALE_EXP_CLAMP + ROW_OVER_CODE_MAX, a));
}

}  // namespace rvq
