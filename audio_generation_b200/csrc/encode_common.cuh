// encode_common.cuh -- pieces shared by the two fused encode kernels (rvq_encode_tc.cu: residual in an
// L2-resident scratch, any d; rvq_encode_tr.cu: residual resident in tensor memory, d <= 128):
// frame addressing, the UMMA A-tile layout, operand scaling and the proven error bound, the two-dimensional
// running-minimum scan, candidate-set enumeration and the job sequence every warp role walks.
#pragma once
#include <cstdint>

#include "common.cuh"
#include "ptx.cuh"

namespace rvq {

constexpr uint32_t A_SLICE_BYTES = TILE_M * KSLICE * 2;  // 16 KiB: [128 rows x 64 fp16], SWIZZLE_128B
constexpr float BIG = 3.0e38f;

struct RowAddrT {
    long long L, sb, sl, sd;
    __device__ __forceinline__ long long row(long long n) const { return (n / L) * sb + (n % L) * sl; }
};


// byte offset of fp16 element (row, col) inside the A tile (d/64 slices of [128 rows x 128 B], SWIZZLE_128B)
// (slice_bytes < A_SLICE_BYTES: tiles that hold fewer than 128 frames keep only those rows of every slice; the MMA
// still reads 128 rows from a slice's base - the lanes beyond the tile see the following slices' bytes and are ignored)
__device__ __forceinline__ uint32_t a_tile_offset(int row, int col, uint32_t slice_bytes = A_SLICE_BYTES) {
    const int slice = col >> 6, c = col & 63;
    return (uint32_t)slice * slice_bytes + (uint32_t)row * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)row & 7u)) << 4) +
           (((uint32_t)c & 7u) << 1);
}

// per-row constants of a stage from the residual's squared norm and the operand scales:
//   na    = 2^(a-b): factor of the scaled code norms in this row's score unit
//   delta = 2.1 x (proven bound on |approximate - exact| score), see DESIGN.md section 3
//   rs    = 2^a ||r||_2 (upper bound): what the allowance of a code above the stage's norm cap is proportional to
__device__ __forceinline__ void row_consts(int d, float sq, bool force_exact, int a, int b, float sb, float cnmax,
                                           float& na_out, float& delta_out, float& rs_out) {
    const float sa = exp2i(a);
    const float na = exp2i(max(-120, min(120, a - b)));
    const float rs = sqrtf(sq) * 1.00002f * sa;  // scaled ||r||_2 (upper bound)
    const float cs = cnmax * sb;                 // scaled max ||c||_2
    // |approx - exact| (scaled units) <= 2^-9(1+..) rs cs  [fp16 rounding of both operands, Cauchy-Schwarz]
    //   + 2^-15 |score|                                   [slack for low mantissa bits used as tags]
    //   + d 2^-14                                          [fp16 subnormal absolute error, accumulate slack]
    const float mag = na * cs * cs + 2.f * rs * cs;  // >= |scaled score| and >= sum of |terms| of the exact scorer
    const float E = 1.02f * 0.001953125f * rs * cs + 3.0517578125e-5f * mag + (float)d * 6.103515625e-5f;
    // the exact fp32 scorer (exact.cuh: d/8 sequential fmaf per accumulator, 3 butterfly adds, one final fmaf) is
    // itself within E32 = (d/8 + 4) 2^-24 mag of the real score; the winner under IT must stay inside best + delta:
    //   s~(k^) <= best + 2 E + 2 E32          
    const float E32 = (float)(d / 8 + 4) * 5.9604645e-8f * mag;
    float delta = 2.1f * E + 2.f * E32;
    if (force_exact || !isfinite(delta)) delta = __int_as_float(0x7f800000);
    na_out = na;
    delta_out = delta;
    rs_out = isfinite(rs) ? rs : 0.f;
}

// operand scale exponent for a row whose entries are bounded by amax_bound, given the stage's b
__device__ __forceinline__ int pick_row_exp(float amax_bound, int b, bool& force_exact) {
    int a = b + ROW_OVER_CODE_MAX;
    if (!isfinite(amax_bound)) force_exact = true;
    if (amax_bound > 0.f && isfinite(amax_bound)) a = min(a, SCALE_TARGET_EXP - ilog2f_floor(amax_bound));
    if (a < b - ROW_UNDER_CODE_MAX) force_exact = true;  // frame >= 2^40 x larger than the codes: no fp16 window
    a = max(a, b - ROW_UNDER_CODE_MAX);
    return max(-100, min(100, a));
}


// Candidate set of one scan group for one frame, factorised: up to three loads (`it`, 9 bits each) x a 16-bit
// column mask.  bits 27-28 = number of loads, bit 31 = OVER (more loads than kept, or no usable filter result).
constexpr uint32_t G_OVER = 0x80000000u;
constexpr uint32_t G_NOFILTER = 0x40000000u;  // no usable filter result at all (NaN / overflow / forced exact)


// ---------------------------------------------------------------------------------------------------------
// Two-dimensional running minimum.  The codes a scan thread sees in one stage are laid out as a grid:
// grid row = one TMEM load of 16 consecutive codes (`it` = code / 16), grid column = position j inside the load.
// Per score the thread pays one FFMA (norm) and one FMNMX into the column minimum Cm[j]; per load it reduces
// the 16 scores to the load's minimum with 3-input minima, replaces its low 9 mantissa bits by `it` and
// inserts it into a sorted triple (+ fourth value).  No per-score index bookkeeping, no per-score sort.
//   * the best score is min_j Cm[j]; it sits at (argmin over loads, argmin over columns);
//   * every code scoring <= T lies in a load whose minimum is <= T AND in a column whose minimum is <= T,
//     so {loads <= T} x {columns <= T} is a superset of the candidates: it is re-scored exactly.
constexpr uint32_t IT_MASK = 0x1FFu;  // 32 chunks x 16 loads

// kX: the chunk holds codes above the stage's norm cap: their allowance rs * xc_k (xptr, shared memory; nrs = -rs) is
// subtracted (optimistic scores, see k0_bound in rvq_aux.cu); the na-proportional part X2_k is inside the stored norm
template <bool kX>
__device__ __forceinline__ void scan16_2d(const uint32_t (&v)[16], const float* __restrict__ nptr,
                                          const float* __restrict__ xptr, float na, float nrs, uint32_t it,
                                          float (&Cm)[16], float& m1, float& m2, float& m3, float& m4, float* dbg) {
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
        const float4 nn = *reinterpret_cast<const float4*>(nptr + j);  // shared memory, warp-uniform
        s[j + 0] = fmaf(na, nn.x, __uint_as_float(v[j + 0]));
        s[j + 1] = fmaf(na, nn.y, __uint_as_float(v[j + 1]));
        s[j + 2] = fmaf(na, nn.z, __uint_as_float(v[j + 2]));
        s[j + 3] = fmaf(na, nn.w, __uint_as_float(v[j + 3]));
        if (kX) {
            const float4 xx = *reinterpret_cast<const float4*>(xptr + j);
            s[j + 0] = fmaf(nrs, xx.x, s[j + 0]);
            s[j + 1] = fmaf(nrs, xx.y, s[j + 1]);
            s[j + 2] = fmaf(nrs, xx.z, s[j + 2]);
            s[j + 3] = fmaf(nrs, xx.w, s[j + 3]);
        }
    }
    if (dbg) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dbg[j] = s[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) Cm[j] = fminf(Cm[j], s[j]);
    float r = fminf(fminf(s[0], s[1]), s[2]);
#pragma unroll
    for (int j = 3; j < 15; j += 2) r = fminf(fminf(r, s[j]), s[j + 1]);
    r = fminf(r, s[15]);
    const float rp = __uint_as_float((__float_as_uint(r) & ~IT_MASK) | it);
    const float t = fmaxf(m1, rp);
    m1 = fminf(m1, rp);
    const float u = fmaxf(m2, t);
    m2 = fminf(m2, t);
    const float w = fmaxf(m3, u);
    m3 = fminf(m3, u);
    m4 = fminf(m4, w);
}

// Allowance X of the code behind a frame's best (optimistic) score `vb` (k0_bound in rvq_aux.cu): the threshold is
// T = vb + delta + 2 X.  The best score sits in column `jmin` (column minima are exact) of a load whose tagged minimum
// lies within the tag's 2^-14 |vb| (x 1.016) of vb, i.e. of one of the three tracked loads unless a FOURTH load is that
// close too - then +inf comes back and the frame takes the exact scan of the whole stage.
// `tab` = the stage's byte table (shared or global memory), xu = rs U1 + na U2 of this frame: X_k <= tab[k] * xu.
__device__ __forceinline__ float best_allowance(float vb, int jmin, float m1, float m2, float m3, float m4, float xu,
                                                const uint8_t* tab, int kmax) {
    const float lim = vb + fabsf(vb) * 6.2e-5f;
    const int k1 = min((int)((__float_as_uint(m1) & IT_MASK) * 16u) + jmin, kmax);  // m1 <= lim always
    const int k2 = min((int)((__float_as_uint(m2) & IT_MASK) * 16u) + jmin, kmax);
    const int k3 = min((int)((__float_as_uint(m3) & IT_MASK) * 16u) + jmin, kmax);
    // three independent loads, selected afterwards: one load latency on the stage's critical path
    const uint32_t b1 = tab[k1], b2 = tab[k2], b3 = tab[k3];
    const uint32_t b = max(b1, max(m2 <= lim ? b2 : 0u, m3 <= lim ? b3 : 0u));
    float x = (float)b * xu;
    x = (m4 <= lim) ? __int_as_float(0x7f800000) : x;
    return (vb < BIG) ? x : 0.f;
}

// same for accumulators that already contain the norm term (rvq_encode_tr.cu folds it into the MMA)
__device__ __forceinline__ void scan16_2d_raw(const uint32_t (&v)[16], uint32_t it, float (&Cm)[16], float& m1, float& m2, float& m3, float& m4,
                                          float* dbg) {
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = __uint_as_float(v[j]);
    if (dbg) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dbg[j] = s[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) Cm[j] = fminf(Cm[j], s[j]);
    float r = fminf(fminf(s[0], s[1]), s[2]);
#pragma unroll
    for (int j = 3; j < 15; j += 2) r = fminf(fminf(r, s[j]), s[j + 1]);
    r = fminf(r, s[15]);
    const float rp = __uint_as_float((__float_as_uint(r) & ~IT_MASK) | it);
    const float t = fmaxf(m1, rp);
    m1 = fminf(m1, rp);
    const float u = fmaxf(m2, t);
    m2 = fminf(m2, t);
    const float w = fmaxf(m3, u);
    m3 = fminf(m3, u);
    m4 = fminf(m4, w);
}

// enumeration of {loads} x {columns} of both scan groups
struct CandSet {
    uint32_t r[2], c[2];
    int n[2], pc[2];
    __device__ __forceinline__ CandSet(uint32_t r0, uint32_t c0, uint32_t r1, uint32_t c1) {
        r[0] = r0;
        r[1] = r1;
        c[0] = c0;
        c[1] = c1;
        pc[0] = __popc(c0);
        pc[1] = __popc(c1);
        n[0] = (int)((r0 >> 27) & 3u) * pc[0];
        n[1] = (int)((r1 >> 27) & 3u) * pc[1];
    }
    __device__ __forceinline__ int total() const { return n[0] + n[1]; }
    // code number e of the set, clamped into [0, kmax] (out-of-range e, or an empty set, gives a harmless code)
    __device__ __forceinline__ int code(int e, int kmax) const {
        e = max(0, min(e, total() - 1));
        const int g = e >= n[0];
        e -= g ? n[0] : 0;
        const int pcg = max(pc[g], 1);
        int a = 0;  // at most three loads per group
        if (e >= pcg) {
            e -= pcg;
            a = 1;
        }
        if (e >= pcg) {
            e -= pcg;
            a = 2;
        }
        uint32_t m = c[g];
        for (int i = 0; i < e && i < 15; ++i) m &= m - 1;  // drop the e lowest set bits
        const uint32_t it = (r[g] >> (9 * a)) & IT_MASK;
        const int k = (int)(it * 16u) + ((__ffs(m) - 1) & 15);
        return max(0, min(k, kmax));
    }
};

// job = (tile slot, tile, stage); every role walks the same sequence.
struct JobIter {
    int n_local, nq, nslots, i, q, slot;  // i = local tile index
    __device__ __forceinline__ JobIter(int n_local_, int nq_, int nslots_)
        : n_local(n_local_), nq(nq_), nslots(nslots_), i(0), q(0), slot(0) {}
    __device__ __forceinline__ bool valid() const { return i < n_local; }
    __device__ __forceinline__ void next() {
        // one slot: tile after tile.  two slots: pair p = (2p, 2p+1): for q: (slot0, q), (slot1, q)
        if (nslots == 1) {
            if (++q == nq) {
                q = 0;
                ++i;
            }
            return;
        }
        const int pair0 = i & ~1;
        if (slot == 0 && pair0 + 1 < n_local) {
            slot = 1;
            i = pair0 + 1;
        } else {
            slot = 0;
            i = pair0;
            if (++q == nq) {
                q = 0;
                i = pair0 + 2;
            }
        }
    }
};


}  // namespace rvq
