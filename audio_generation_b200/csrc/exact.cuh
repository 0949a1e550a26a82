// exact.cuh -- the ONE definition of the exact fp32 score used by every GPU path
// (re-rank of tensor-core candidates, in-kernel fallback scan, and the exact-scan kernel),
// so that all of them agree bit for bit with each other.
//
//   score(r, c) = ||c||^2 - 2 r.c      (the form of oracle/rvq_oracle.py:stage_scores)
//
// An 8-lane group owns one (row, code) pair.  Lane j of the group owns the float4 pieces
// j, j+8, j+16, ... of the d-vector, accumulates dot and norm with fmaf in that order and
// the 8 partials are combined by an xor-butterfly (1, 2, 4).  The batched variant scores NC codes
// against one row with all loads of a 128-feature segment issued before the arithmetic (memory-level
// parallelism); per code the operation order is identical, so results are bit-identical.
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

// One segment of NP float4 pieces per lane (NP*32 features) starting at feature p0 (lane offset included).
template <int NC, int NP>
__device__ __forceinline__ void exact_seg(const float* r, const float* const* c, int p0, float* dot,
                                          float* nrm) {
    float4 rv[NP], cv[NC][NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        rv[i] = *reinterpret_cast<const float4*>(r + p0 + i * 32);
#pragma unroll
        for (int j = 0; j < NC; ++j) cv[j][i] = ldg_nc_v4(c[j] + p0 + i * 32);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            dot[j] = fmaf(rv[i].x, cv[j][i].x, dot[j]);
            dot[j] = fmaf(rv[i].y, cv[j][i].y, dot[j]);
            dot[j] = fmaf(rv[i].z, cv[j][i].z, dot[j]);
            dot[j] = fmaf(rv[i].w, cv[j][i].w, dot[j]);
            nrm[j] = fmaf(cv[j][i].x, cv[j][i].x, nrm[j]);
            nrm[j] = fmaf(cv[j][i].y, cv[j][i].y, nrm[j]);
            nrm[j] = fmaf(cv[j][i].z, cv[j][i].z, nrm[j]);
            nrm[j] = fmaf(cv[j][i].w, cv[j][i].w, nrm[j]);
        }
    }
}

// r: row vector (generic pointer: shared or global), c[j]: code vectors (global). d % 64 == 0.
// All 8 lanes of the group obtain the same out[j].
template <int NC>
__device__ __forceinline__ void exact_score8_n(const float* r, const float* const* c, int d, int sub,
                                               float* out) {
    float dot[NC], nrm[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) dot[j] = nrm[j] = 0.f;
    int p0 = sub * 4;
#pragma unroll 1
    for (; p0 + 128 <= d + sub * 4; p0 += 128) exact_seg<NC, 4>(r, c, p0, dot, nrm);
    if (d & 64) exact_seg<NC, 2>(r, c, p0, dot, nrm);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            dot[j] += __shfl_xor_sync(0xffffffffu, dot[j], o);
            nrm[j] += __shfl_xor_sync(0xffffffffu, nrm[j], o);
        }
        out[j] = fmaf(-2.f, dot[j], nrm[j]);
    }
}

// NC independent (row, code) pairs scored together by one 8-lane group: the code loads of all pairs (global,
// long latency) are in flight at once, the rows come from shared memory piece by piece.  Per pair the
// arithmetic is exact_score8_n's (same fmaf chain, same butterfly), hence bit-identical scores.
template <int NC, int NP>
__device__ __forceinline__ void exact_seg_pairs(const float* const* r, const float* const* c, int p0, float* dot,
                                                float* nrm) {
    float4 cv[NC][NP];
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int i = 0; i < NP; ++i) cv[j][i] = ldg_nc_v4(c[j] + p0 + i * 32);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float4 rv = *reinterpret_cast<const float4*>(r[j] + p0 + i * 32);
            dot[j] = fmaf(rv.x, cv[j][i].x, dot[j]);
            dot[j] = fmaf(rv.y, cv[j][i].y, dot[j]);
            dot[j] = fmaf(rv.z, cv[j][i].z, dot[j]);
            dot[j] = fmaf(rv.w, cv[j][i].w, dot[j]);
            nrm[j] = fmaf(cv[j][i].x, cv[j][i].x, nrm[j]);
            nrm[j] = fmaf(cv[j][i].y, cv[j][i].y, nrm[j]);
            nrm[j] = fmaf(cv[j][i].z, cv[j][i].z, nrm[j]);
            nrm[j] = fmaf(cv[j][i].w, cv[j][i].w, nrm[j]);
        }
    }
}
template <int NC>
__device__ __forceinline__ void exact_score8_pairs(const float* const* r, const float* const* c, int d, int sub,
                                                   float* out) {
    float dot[NC], nrm[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) dot[j] = nrm[j] = 0.f;
    int p0 = sub * 4;
#pragma unroll 1
    for (; p0 + 128 <= d + sub * 4; p0 += 128) exact_seg_pairs<NC, 4>(r, c, p0, dot, nrm);
    if (d & 64) exact_seg_pairs<NC, 2>(r, c, p0, dot, nrm);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            dot[j] += __shfl_xor_sync(0xffffffffu, dot[j], o);
            nrm[j] += __shfl_xor_sync(0xffffffffu, nrm[j], o);
        }
        out[j] = fmaf(-2.f, dot[j], nrm[j]);
    }
}

__device__ __forceinline__ float exact_score8(const float* r, const float* __restrict__ c, int d,
                                              int sub) {
    const float* const cc[1] = {c};
    float o[1];
    exact_score8_n<1>(r, cc, d, sub, o);
    return o[0];
}

// Exact argmin over codes [k0, k1) for one row by ONE WARP: 4 groups of 8 lanes, 4 codes per group per step
// (16 codes per warp step, all loads of a step in flight together).  Every lane returns the warp-wide best.
__device__ __forceinline__ ScoreIdx exact_scan_warp(const float* r, const float* __restrict__ cbq,
                                                    int d, int k0, int k1, int lane) {
    const int sub = lane & 7, grp = lane >> 3;
    float bs = __int_as_float(0x7f800000);
    int bk = 0x7fffffff;
    constexpr int NC = 4;
#pragma unroll 1
    for (int kb = k0; kb < k1; kb += 4 * NC) {
        int k[NC];
        const float* cc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            k[j] = kb + j * 4 + grp;
            cc[j] = cbq + (size_t)(k[j] < k1 ? k[j] : k0) * d;
        }
        float s[NC];
        exact_score8_n<NC>(r, cc, d, sub, s);
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (k[j] < k1 && better(s[j], k[j], bs, bk)) {
                bs = s[j];
                bk = k[j];
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


// Exact argmin by ONE WARP over the codes {16 * it + j : it0 <= it < it1, bit j of cols set, code < Kv}
// (4 groups of 8 lanes, 4 codes per group per step).  Every lane returns the warp-wide best.
__device__ __forceinline__ ScoreIdx exact_scan_cols(const float* r, const float* __restrict__ cbq, int d, int it0,
                                                    int it1, uint32_t cols, int Kv, int lane) {
    const int sub = lane & 7, grp = lane >> 3;
    float bs = __int_as_float(0x7f800000);
    int bk = 0x7fffffff;
    constexpr int NC = 4;
    const int pc = __popc(cols);
    const int total = (it1 - it0) * pc;
#pragma unroll 1
    for (int eb = 0; eb < total; eb += 4 * NC) {
        int k[NC];
        const float* cc[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int e = eb + j * 4 + grp;
            const int ec = e < total ? e : 0;
            const int a = ec / pc, jj = ec - a * pc;
            const int kk = (it0 + a) * 16 + (int)__fns(cols, 0, jj + 1);
            k[j] = (e < total && kk < Kv) ? kk : -1;
            cc[j] = cbq + (size_t)(k[j] >= 0 ? k[j] : 0) * d;
        }
        float s[NC];
        exact_score8_n<NC>(r, cc, d, sub, s);
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (k[j] >= 0 && better(s[j], k[j], bs, bk)) {
                bs = s[j];
                bk = k[j];
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

// The same score computed by ONE lane (bit-identical to exact_score8: accumulator j plays lane j of the group,
// the final sums follow the xor-butterfly order).  r must be readable by every lane (shared memory).
__device__ __forceinline__ float exact_score_lane(const float* r, const float* __restrict__ c, int d) {
    float dot[8], nrm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dot[j] = nrm[j] = 0.f;
#pragma unroll 1
    for (int p = 0; p < d; p += 32) {
        float4 cv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cv[j] = ldg_nc_v4(c + p + j * 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 rv = *reinterpret_cast<const float4*>(r + p + j * 4);
            dot[j] = fmaf(rv.x, cv[j].x, dot[j]);
            dot[j] = fmaf(rv.y, cv[j].y, dot[j]);
            dot[j] = fmaf(rv.z, cv[j].z, dot[j]);
            dot[j] = fmaf(rv.w, cv[j].w, dot[j]);
            nrm[j] = fmaf(cv[j].x, cv[j].x, nrm[j]);
            nrm[j] = fmaf(cv[j].y, cv[j].y, nrm[j]);
            nrm[j] = fmaf(cv[j].z, cv[j].z, nrm[j]);
            nrm[j] = fmaf(cv[j].w, cv[j].w, nrm[j]);
        }
    }
    const float d01 = dot[0] + dot[1], d23 = dot[2] + dot[3], d45 = dot[4] + dot[5], d67 = dot[6] + dot[7];
    const float n01 = nrm[0] + nrm[1], n23 = nrm[2] + nrm[3], n45 = nrm[4] + nrm[5], n67 = nrm[6] + nrm[7];
    const float dt = (d01 + d23) + (d45 + d67);
    const float nt = (n01 + n23) + (n45 + n67);
    return fmaf(-2.f, dt, nt);
}

// Exact argmin over codes [k0, k1) for one row by ONE WARP, one code per lane per step (row vector in shared
// memory).  Every lane returns the warp-wide best.  Same scores, hence same winner, as exact_scan_warp.
__device__ __forceinline__ ScoreIdx exact_scan_warp_lanes(const float* r_smem,
                                                          const float* __restrict__ cbq, int d, int k0, int k1,
                                                          int lane) {
    float bs = __int_as_float(0x7f800000);
    int bk = 0x7fffffff;
#pragma unroll 1
    for (int k = k0 + lane; k < k1; k += 32) {
        const float s = exact_score_lane(r_smem, cbq + (size_t)k * d, d);
        if (better(s, k, bs, bk)) {
            bs = s;
            bk = k;
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
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
