// exact.cuh -- the ONE definition of the exact fp32 score used by every GPU path
// (re-rank of tensor-core candidates, in-kernel fallback scan, and the exact-scan kernel),
// so that all of them agree bit for bit with each other.
//
//   score(r, c) = ||c||^2 - 2 r.c      (the form of oracle/rvq_oracle.py:stage_scores)
//
// An 8-lane group owns one (row, code) pair.  Lane j of the group owns the float4 pieces
// j, j+8, j+16, ... of the d-vector, accumulates dot and norm with fmaf in that order and
// the 8 partials are combined by an xor-butterfly (1, 2, 4).
#pragma once
#include "ptx.cuh"

namespace rvq {

struct ScoreIdx {
    float s;
    int k;
};

__device__ __forceinline__ bool better(float s, int k, float bs, int bk) {
    return (s < bs) || (s == bs && k < bk);  // lowest index wins exact ties (torch argmin on CPU)
}

// r: row vector (generic pointer: shared or global), c: code vector (global). d % 32 == 0.
// All 8 lanes of the group return the same value.
__device__ __forceinline__ float exact_score8(const float* __restrict__ r, const float* __restrict__ c, int d,
                                              int sub) {
    float dot = 0.f, nrm = 0.f;
    for (int p = sub * 4; p < d; p += 32) {
        const float4 rv = *reinterpret_cast<const float4*>(r + p);
        const float4 cv = ldg_nc_v4(c + p);
        dot = fmaf(rv.x, cv.x, dot);
        dot = fmaf(rv.y, cv.y, dot);
        dot = fmaf(rv.z, cv.z, dot);
        dot = fmaf(rv.w, cv.w, dot);
        nrm = fmaf(cv.x, cv.x, nrm);
        nrm = fmaf(cv.y, cv.y, nrm);
        nrm = fmaf(cv.z, cv.z, nrm);
        nrm = fmaf(cv.w, cv.w, nrm);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    }
    return fmaf(-2.f, dot, nrm);
}

// Exact argmin over codes [k0, k1) for one row by ONE WARP (4 groups of 8 lanes, one code each per step).
// Every lane returns the warp-wide best.
__device__ __forceinline__ ScoreIdx exact_scan_warp(const float* __restrict__ r, const float* __restrict__ cbq,
                                                    int d, int k0, int k1, int lane) {
    const int sub = lane & 7, grp = lane >> 3;
    float bs = __int_as_float(0x7f800000);
    int bk = 0x7fffffff;
    for (int kb = k0; kb < k1; kb += 4) {
        const int k = kb + grp;
        const bool valid = k < k1;
        const float s = exact_score8(r, cbq + (size_t)(valid ? k : k0) * d, d, sub);
        if (valid && better(s, k, bs, bk)) {
            bs = s;
            bk = k;
        }
    }
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (better(os, ok, bs, bk)) {
            bs = os;
            bk = ok;
        }
    }
    return ScoreIdx{bs, bk};
}

}  // namespace rvq
