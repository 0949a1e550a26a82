// rvq_aux.cu -- the HBM-bound helpers around the fused encode kernel:
//   K0  codebook operand preparation (fp16 scaled copy, scaled norms, per-stage metadata)
//   K3  EMA finalize (count/sum decay, Laplace smoothing, codebook refresh)
//   dequantize (code lookup summed over stages)
//   exact-scan encode (every code scored in fp32 on CUDA cores; verification / RVQ_ALGO_EXACT_SCAN)
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"
#include "exact.cuh"

namespace rvq {

// ------------------------------------------------------------------------------------------ K0
// K0_SPLIT blocks per stage: max |c| over the valid codes -> meta[q] = {2^b, 0 (cnmax, filled by k0_convert), cmax, Kv}.
// Every block folds its slice into meta[q][2] with an integer atomicMax (non-negative floats order like ints; NaNs are
// dropped by fmaxf before); the block that takes the last ticket of the stage (meta[q][5], zeroed by the caller
// together with meta[q][2]) writes the derived entries.  One block per stage took 119 us per update step on C3.
constexpr int K0_SPLIT = 16;
__global__ void k0_stage_max(const float* __restrict__ cb, const int* __restrict__ k_valid, int K, int d,
                             float* __restrict__ meta) {
    const int q = blockIdx.x;
    const int Kv = k_valid ? min(max(k_valid[q], 0), K) : K;
    const float* base = cb + (size_t)q * K * d;
    const size_t n = (size_t)Kv * d;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x)
        m = fmaxf(m, fabsf(base[i]));
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) {
            float* mq = meta + q * META_STRIDE;
            atomicMax(reinterpret_cast<int*>(mq + 2), __float_as_int(m));
            __threadfence();
            const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(mq + 5), 1u);
            if (ticket == gridDim.y - 1) {
                __threadfence();
                const float mall = __int_as_float(atomicMax(reinterpret_cast<int*>(mq + 2), 0));
                mq[0] = exp2i(code_scale_exp(mall));
                mq[1] = 0.f;
                mq[3] = (float)Kv;
                mq[4] = (float)gridDim.x;  // stages prepared: locates the norm slices behind the norms
                mq[5] = mq[6] = mq[7] = 0.f;
            }
        }
    }
}

// one warp per (stage, padded code row): fp16 operand row, scaled norm, running max of ||c||_2
// Norm slice of a code (rvq_encode_tr.cu folds the norm into the MMA as one extra K = 16 step): 16 fp16 columns
// {B1, B2, B3, PAD, 0...} with 2^11 B1 + 2 B2 + 2^-4 B3 = scaled norm exactly (three 11-bit pieces of the fp32
// value) and PAD = 65504 for padding codes.  Stored per 128-code chunk in the canonical no-swizzle K-major UMMA
// layout (8-row x 16-byte core matrices: row j, 16-byte K chunk kc at (j / 8) * 256 + kc * 128 + (j % 8) * 16).
__device__ __forceinline__ void write_norm_slice(uint8_t* slices, int q, int Kpad, int k, float n, bool pad, float xc) {
    uint8_t* chunk = slices + ((size_t)q * (Kpad / 128) + k / 128) * 4096;
    const int j = k % 128;
    uint8_t* p0 = chunk + (j / 8) * 256 + (j % 8) * 16;
    float h1 = 0.f, h2 = 0.f, h3 = 0.f;
    if (!pad) {
        h1 = __uint_as_float(__float_as_uint(n) & 0xFFFFE000u);
        const float rem = n - h1;
        h2 = __uint_as_float(__float_as_uint(rem) & 0xFFFFE000u);
        h3 = rem - h2;
    }
    const __half2 a = __floats2half2_rn(h1 * 4.8828125e-4f, h2 * 0.5f);
    const __half2 b = __floats2half2_rn(h3 * 16.f, pad ? 65504.f : 0.f);
    // fifth column: -(64 xc), rounded AWAY from zero (the operand row holds rs / 64 there, rounded up): the optimistic
    // allowance rs * xc of a code whose norm exceeds the stage's cap (k0_bound)
    const __half2 c = __halves2half2(__hneg(__float2half_ru(xc * 64.f)), __float2half_rn(0.f));
    uint4 v;
    v.x = *reinterpret_cast<const uint32_t*>(&a);
    v.y = *reinterpret_cast<const uint32_t*>(&b);
    v.z = *reinterpret_cast<const uint32_t*>(&c);
    v.w = 0u;
    *reinterpret_cast<uint4*>(p0) = v;
    *reinterpret_cast<uint4*>(p0 + 128) = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void k0_convert(const float* __restrict__ cb, int nq, int K, int Kpad, int d, __half* __restrict__ op,
                           float* __restrict__ norm, float* __restrict__ meta) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nq * Kpad) return;
    const int q = warp / Kpad, k = warp % Kpad;
    float* mq = meta + q * META_STRIDE;
    const float sb = mq[0];
    const int Kv = (int)mq[3];
    __half* orow = op + ((size_t)q * Kpad + k) * d;
    if (k >= Kv) {
        for (int i = lane; i < d; i += 32) orow[i] = __float2half_rn(0.f);
        if (lane == 0) norm[(size_t)q * Kpad + k] = PAD_NORM;
        return;
    }
    const float* crow = cb + ((size_t)q * K + k) * d;
    float nrm = 0.f;
    const float m2sb = -2.f * sb;
    for (int i = lane; i < d; i += 32) {
        const float c = crow[i];
        nrm = fmaf(c, c, nrm);
        orow[i] = __float2half_rn(c * m2sb);  // exact power-of-two scaling, one fp16 rounding
    }
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) {
        norm[(size_t)q * Kpad + k] = nrm * sb * sb;   // raw scaled norm; k0_bound finishes it
        // upper bound of ||c||_2 (fp32 summation slack) ; positive floats order like ints
        const float cn = sqrtf(nrm) * (1.f + 1e-5f);
        atomicMax(reinterpret_cast<int*>(mq + 1), __float_as_int(cn));
    }
}

// Per-stage norm cap and the per-code allowances of the codes above it (one block per stage).
//
// The filter's error bound grows with the norm of the code.  A trained codebook holds a few codes
// several times larger than the ones in use (dead codes that kept their initial scale): building the per-frame bound
// from the LARGEST norm made the threshold 2-6x looser than the live codes need (profiles/r2e_c3_shards_probe.log).
// Instead the stage gets a cap cs0 = min(max, 1.5 x the 25th percentile of the scaled norms); the frame's bound E0 is
// built from cs0, and a code with cs_k > cs0 ("large") has its own excess X_k >= E_k + E32_k - (E0 + E32_0) SUBTRACTED
// from its approximate score (an optimistic score v = s~ - X_k), which keeps the certificate rigorous:
//     v(winner) <= v(j*) + 2 (E0 + E32_0) + 2 X_{j*}        for j* = argmin v        (derivation in DESIGN.md)
// X_k = rs * xc_k + na * X2_k with  xc_k = c1 dcs_k,  X2_k = c2 (cs_k^2 - cs0^2),  dcs_k = cs_k - cs0,
// c1 = alpha + 2 beta + 2 gamma,  c2 = beta + gamma  (row_consts); rs = 2^a ||r|| and na = 2^(a-b) belong to the frame.
// X2_k is folded into the stored norm; rs * xc_k is one more rank-1 term of the extra MMA step (fifth norm-slice
// column x rs in the operand row; TMEM kernels) or an FFMA in the epilogue of the chunks that hold a large code
// (generic kernel: xc array, xflag per 256-code chunk).
// The cap is the stage's lower-quartile norm (cs_max itself when the norms are concentrated, max <= 1.25 x quartile:
// freshly initialised codebooks - no code is flagged and the generic kernel's epilogue stays on its short path).
// Measured on C3 (profiles/r2g_probe_*.log, encode kernel with statistics after 150 updates, 1 / 8 shards' state):
// cap 1.5 x p25 14.66 / 15.41 ms, cap p25 13.86 / 14.72 ms (16.13 / 19.07 ms with the stage maximum).  Once the
// allowances are per code, WHICH low quantile is the cap no longer matters: p10 / p25 / p50 give 12.89 / 12.89 / 12.92 ms
// on the final kernel (profiles/r3j_cap_quantile.log).
#ifndef RVQ_CAP_DIV
#define RVQ_CAP_DIV 4
#endif
#ifndef RVQ_CAP_F
#define RVQ_CAP_F 1.0f
#endif
#ifndef RVQ_CAP_FLAT
#define RVQ_CAP_FLAT 1.25f
#endif
__global__ void __launch_bounds__(1024) k0_bound(float* __restrict__ norm, uint8_t* __restrict__ slices,
                                                 float* __restrict__ xc, uint8_t* __restrict__ xb, int* __restrict__ xflag,
                                                 float* __restrict__ meta, int Kpad, int d) {
    const int q = blockIdx.x, t = threadIdx.x;
    float* mq = meta + q * META_STRIDE;
    const float sb = mq[0], cnmax = mq[1];
    const int Kv = (int)mq[3];
    __shared__ float samp[1024];
    __shared__ float s_p25;
    __shared__ int s_xcmax, s_x2max;
    if (t == 0) s_xcmax = s_x2max = 0;
    const int n = min(Kv, 1024);
    const int stride = n > 0 ? Kv / n : 1;
    if (t < n) samp[t] = sqrtf(norm[(size_t)q * Kpad + (size_t)t * stride]);
    if (t == 0) s_p25 = 0.f;
    __syncthreads();
    if (t < n) {
        const float mine = samp[t];
        int rank = 0;
        for (int u = 0; u < n; ++u) rank += (samp[u] < mine || (samp[u] == mine && u < t)) ? 1 : 0;
        if (rank == n / RVQ_CAP_DIV) s_p25 = mine;
    }
    __syncthreads();
    const float cs_max = cnmax * sb;
    const float cs0 = (n > 0 && cs_max > RVQ_CAP_FLAT * s_p25) ? fminf(cs_max, RVQ_CAP_F * s_p25 * (1.f + 1e-5f)) : cs_max;
    const float gamma = (float)(d / 8 + 4) * 5.9604645e-8f;
    const float c1 = (1.02f * 0.001953125f + 2.f * 3.0517578125e-5f + 2.f * gamma) * 1.01f;
    const float c2 = (3.0517578125e-5f + gamma) * 1.01f;
    // pass 1: stage maxima of the two allowance factors;  pass 2: everything that is stored
    auto factors = [&](int k, float nk, float& XC, float& X2) {
        const float csk = sqrtf(nk) * (1.f + 1e-5f);
        const bool large = k < Kv && csk > cs0;
        XC = large ? c1 * (csk - cs0) : 0.f;
        X2 = large ? c2 * (csk * csk - cs0 * cs0) : 0.f;
    };
    {
        float mxc = 0.f, mx2 = 0.f;
        for (int k = t; k < Kpad; k += blockDim.x) {
            float XC, X2;
            factors(k, norm[(size_t)q * Kpad + k], XC, X2);
            mxc = fmaxf(mxc, XC);
            mx2 = fmaxf(mx2, X2);
        }
        // (non-negative floats order like ints)
        if (mxc > 0.f) atomicMax(&s_xcmax, __float_as_int(mxc));
        if (mx2 > 0.f) atomicMax(&s_x2max, __float_as_int(mx2));
    }
    __syncthreads();
    // one byte per code: X_k <= xb_k (rs U1 + na U2) with U = stage maximum / 255 (what the kernels look up for the
    // code behind a frame's best score: shared-memory table in the TMEM kernel)
    const float U1 = __int_as_float(s_xcmax) * (1.f / 255.f) * (1.f + 2e-6f);
    const float U2 = __int_as_float(s_x2max) * (1.f / 255.f) * (1.f + 2e-6f);
    for (int k = t; k < Kpad; k += blockDim.x) {
        const bool pad = k >= Kv;
        const size_t i = (size_t)q * Kpad + k;
        const float nk = norm[i];
        float XC, X2;
        factors(k, nk, XC, X2);
        xc[i] = XC;
        int bq = 0;
        if (XC > 0.f) {
            bq = (int)ceilf(XC / U1);
            if (U2 > 0.f) bq = max(bq, (int)ceilf(X2 / U2));
            while (bq < 255 && ((float)bq * U1 < XC || (float)bq * U2 < X2)) ++bq;
            bq = min(max(bq, 1), 255);
            atomicOr(xflag + (size_t)q * (Kpad / CHUNK_N) + k / CHUNK_N, 1);  // (zeroed by the caller)
        }
        xb[i] = (uint8_t)bq;
        const float np = pad ? PAD_NORM : nk - X2;
        norm[i] = np;
        write_norm_slice(slices, q, Kpad, k, pad ? 0.f : np, pad, XC);
    }
    if (t == 0) {
        mq[6] = cnmax;      // the largest norm (diagnostics)
        mq[1] = cs0 / sb;   // what row_consts builds the frame's bound from
        mq[7] = U1;         // units of the byte table
        mq[5] = U2;
    }
}

// ------------------------------------------------------------------------------------------ K3
// one block per stage: EMA of the counts and n_tot = sum_k ema_count (written to scratch[q])
__global__ void k3_counts(float* __restrict__ ema_count, const float* __restrict__ cnt, const int* __restrict__ k_valid,
                          int K, float decay, float omd, float* __restrict__ ntot) {
    const int q = blockIdx.x;
    const int Kv = k_valid ? min(max(k_valid[q], 0), K) : K;
    float acc = 0.f;
    for (int k = threadIdx.x; k < Kv; k += blockDim.x) {
        const size_t i = (size_t)q * K + k;
        const float v = __fadd_rn(__fmul_rn(decay, ema_count[i]), __fmul_rn(omd, cnt[i]));
        ema_count[i] = v;
        acc += v;
    }
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) ntot[q] = acc;
    }
}

// one warp per (stage, code): EMA of the sums and the smoothed codebook refresh
__global__ void k3_codes(float* __restrict__ cb, const float* __restrict__ ema_count, float* __restrict__ ema_sum,
                         const float* __restrict__ sum, const int* __restrict__ k_valid, const float* __restrict__ ntot,
                         int nq, int K, int d, float decay, float omd, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nq * K) return;
    const int q = warp / K, k = warp % K;
    const int Kv = k_valid ? min(max(k_valid[q], 0), K) : K;
    if (k >= Kv) return;
    const float n = ntot[q];
    const float smoothed = __fmul_rn(__fdiv_rn(__fadd_rn(ema_count[(size_t)q * K + k], eps),
                                               __fadd_rn(n, __fmul_rn((float)Kv, eps))), n);
    const size_t base = ((size_t)q * K + k) * d;
    for (int i = lane * 4; i < d; i += 128) {
        float4 es = *reinterpret_cast<const float4*>(ema_sum + base + i);
        const float4 sv = *reinterpret_cast<const float4*>(sum + base + i);
        es.x = __fadd_rn(__fmul_rn(decay, es.x), __fmul_rn(omd, sv.x));
        es.y = __fadd_rn(__fmul_rn(decay, es.y), __fmul_rn(omd, sv.y));
        es.z = __fadd_rn(__fmul_rn(decay, es.z), __fmul_rn(omd, sv.z));
        es.w = __fadd_rn(__fmul_rn(decay, es.w), __fmul_rn(omd, sv.w));
        *reinterpret_cast<float4*>(ema_sum + base + i) = es;
        float4 cv;
        cv.x = __fdiv_rn(es.x, smoothed);
        cv.y = __fdiv_rn(es.y, smoothed);
        cv.z = __fdiv_rn(es.z, smoothed);
        cv.w = __fdiv_rn(es.w, smoothed);
        *reinterpret_cast<float4*>(cb + base + i) = cv;
    }
}

// ------------------------------------------------------------------------------------------ dequantize
struct RowAddr {
    long long L, sb, sl, sd;
    __device__ __forceinline__ long long row(long long n) const { return (n / L) * sb + (n % L) * sl; }
};

__global__ void dequantize_rows(const float* __restrict__ cb, const long long* __restrict__ idx, long long N,
                                RowAddr ad, int d, int q0, int nq, int K, float4 w0, const float* __restrict__ w,
                                int accumulate, float* __restrict__ out) {
    // one warp per frame when features are contiguous, else frames-fastest scalar mapping
    if (ad.sd == 1) {
        const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31;
        if (n >= N) return;
        float* o = out + ad.row(n);
        for (int i = lane * 4; i < d; i += 128) {
            float4 acc = accumulate ? *reinterpret_cast<const float4*>(o + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < nq; ++q) {
                long long k = idx[n * nq + q];
                k = k < 0 ? 0 : (k >= K ? K - 1 : k);
                const float4 c = *reinterpret_cast<const float4*>(cb + ((size_t)(q0 + q) * K + k) * d + i);
                const float wq = w ? w[q] : 1.f;
                acc.x = fmaf(wq, c.x, acc.x);
                acc.y = fmaf(wq, c.y, acc.y);
                acc.z = fmaf(wq, c.z, acc.z);
                acc.w = fmaf(wq, c.w, acc.w);
            }
            *reinterpret_cast<float4*>(o + i) = acc;
        }
    } else {
        // launched with ceil(N / 128) * 128 * d threads: the last group of 128 frames is padded, so the guard is on
        // the frame number only (a guard on t < N * d dropped features of the last group when N % 128 != 0)
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long chunk = 128;  // frames-fastest inside groups of 128 frames
        const long long g = t / (chunk * d), rem = t % (chunk * d);
        const int i = (int)(rem / chunk);
        const long long n = g * chunk + rem % chunk;
        if (n >= N) return;
        float* o = out + ad.row(n) + (long long)i * ad.sd;
        float acc = accumulate ? *o : 0.f;
        for (int q = 0; q < nq; ++q) {
            long long k = idx[n * nq + q];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            acc = fmaf(w ? w[q] : 1.f, cb[((size_t)(q0 + q) * K + k) * d + i], acc);
        }
        *o = acc;
    }
}

// ------------------------------------------------------------------------------------------ exact-scan encode
// One warp per frame; the residual lives in shared memory; every code of every stage is scored with
// exact_score8.  Slow (CUDA cores) but it is the semantic definition the tensor path must reproduce.
__global__ void __launch_bounds__(256) encode_exact_scan(const float* __restrict__ x, long long N, RowAddr ad, int d,
                                                         int nq, int K, const float* __restrict__ cb,
                                                         const float* __restrict__ meta, float* __restrict__ xq,
                                                         long long* __restrict__ idx, double* __restrict__ commit_sq,
                                                         float* __restrict__ stats_sum, float* __restrict__ stats_cnt) {
    extern __shared__ __align__(16) float smem_r[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (n >= N) return;
    float* r = smem_r + (size_t)warp * d;
    const float* xr = x + ad.row(n);
    for (int i = lane; i < d; i += 32) r[i] = xr[(long long)i * ad.sd];
    __syncwarp();
    for (int q = 0; q < nq; ++q) {
        const float* cbq = cb + (size_t)q * K * d;
        const int Kv = (int)meta[q * META_STRIDE + 3];
        ScoreIdx best = exact_scan_warp(r, cbq, d, 0, Kv, lane);
        int kw = best.k;
        if (kw < 0 || kw >= Kv) kw = 0;
        const float* cw = cbq + (size_t)kw * d;
        float sq = 0.f;
        for (int i = lane * 4; i < d; i += 128) {
            const float4 rv = *reinterpret_cast<const float4*>(r + i);
            const float4 cv = ldg_nc_v4(cw + i);
            if (stats_sum) red_add_v4(stats_sum + ((size_t)q * K + kw) * d + i, rv);
            float4 nr;
            nr.x = rv.x - cv.x;
            nr.y = rv.y - cv.y;
            nr.z = rv.z - cv.z;
            nr.w = rv.w - cv.w;
            *reinterpret_cast<float4*>(r + i) = nr;
            sq = fmaf(nr.x, nr.x, sq);
            sq = fmaf(nr.y, nr.y, sq);
            sq = fmaf(nr.z, nr.z, sq);
            sq = fmaf(nr.w, nr.w, sq);
        }
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) {
            idx[n * nq + q] = kw;
            atomicAdd(commit_sq + q, (double)sq);
            if (stats_cnt) atomicAdd(stats_cnt + (size_t)q * K + kw, 1.f);
        }
        __syncwarp();
    }
    float* xo = xq + ad.row(n);
    for (int i = lane; i < d; i += 32) xo[(long long)i * ad.sd] = xr[(long long)i * ad.sd] - r[i];
}

}  // namespace rvq

// ------------------------------------------------------------------------------------------ host wrappers
using namespace rvq;

extern "C" int rvq_prepared_bytes(int nq, int K, int d, size_t* op_bytes, size_t* norm_bytes, size_t* meta_bytes) {
    if (nq <= 0 || K <= 0 || d <= 0) {
        set_error("rvq_prepared_bytes: nq, K, d must be positive (got %d, %d, %d)", nq, K, d);
        return RVQ_ERR_ARG;
    }
    const size_t Kpad = round_up(K, CHUNK_N);
    if (op_bytes) *op_bytes = (size_t)nq * Kpad * d * sizeof(__half);
    // scaled norms [nq, Kpad] fp32, the fp16 norm slices [nq, Kpad / 128, 4096 bytes], the allowances of the codes
    // above each stage's norm cap (xc per code, a byte table of the whole allowance), per-chunk flags (NormLayout)
    if (norm_bytes) *norm_bytes = NormLayout::bytes(nq, (int)Kpad);
    if (meta_bytes) *meta_bytes = (size_t)nq * META_STRIDE * sizeof(float);
    return RVQ_OK;
}

int rvq_check_shape(const char* who, int nq, int K, int d) {
    if (nq <= 0 || K <= 0 || d <= 0 || d > MAX_D || d % KSLICE != 0) {
        set_error("%s: unsupported shape nq=%d K=%d d=%d (need nq,K >= 1 and d a multiple of %d, <= %d)", who, nq, K, d,
                  KSLICE, MAX_D);
        return RVQ_ERR_ARG;
    }
    return RVQ_OK;
}

extern "C" int rvq_prepare_codebooks(const float* cb, const int* k_valid, int nq, int K, int d, void* cb_op,
                                     float* cb_norm, float* cb_meta, void* stream) {
    if (int e = rvq_check_shape("rvq_prepare_codebooks", nq, K, d)) return e;
    if (!cb || !cb_op || !cb_norm || !cb_meta) {
        set_error("rvq_prepare_codebooks: null pointer");
        return RVQ_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Kpad = round_up(K, CHUNK_N);
    RVQ_CUDA(cudaMemsetAsync(cb_meta, 0, sizeof(float) * (size_t)nq * META_STRIDE, st));
    k0_stage_max<<<dim3((unsigned)nq, K0_SPLIT), 512, 0, st>>>(cb, k_valid, K, d, cb_meta);
    const long long warps = (long long)nq * Kpad;
    const int block = 256;
    const long long grid = (warps * 32 + block - 1) / block;
    k0_convert<<<(unsigned)grid, block, 0, st>>>(cb, nq, K, Kpad, d, static_cast<__half*>(cb_op), cb_norm, cb_meta);
    const NormLayout nl(cb_norm, nq, Kpad);
    RVQ_CUDA(cudaMemsetAsync(nl.xflag, 0, sizeof(int) * (size_t)nq * (Kpad / CHUNK_N), st));
    k0_bound<<<(unsigned)nq, 1024, 0, st>>>(cb_norm, nl.slices, nl.xc, nl.xb, nl.xflag, cb_meta, Kpad, d);
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}

// Stream-ordered scratch for the few floats K3 and the weighted lookup need.  The device's DEFAULT memory pool hands
// freed memory back to the driver at the next synchronisation (release threshold 0), after which the next
// cudaMallocAsync maps memory again - the suspected cause of an occasional 0.7-1 ms between the encode kernel and the
// end of K3 (profiles/r2n_bench_c2.json: collective.k0_k3_ms 0.70 against 0.02-0.03 in the other runs).  An own pool
// that keeps what it has (threshold = max) makes the allocation a pointer bump either way.
static cudaError_t scratch_alloc(float** out, size_t bytes, cudaStream_t st) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(out, bytes, st);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&pools[dev], &props);
            if (e != cudaSuccess) return e;
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = pools[dev];
    }
    return cudaMallocFromPoolAsync(reinterpret_cast<void**>(out), bytes, pool, st);
}

extern "C" int rvq_ema_finalize(float* cb, float* ema_count, float* ema_sum, const float* stats_sum,
                                const float* stats_cnt, const int* k_valid, int nq_use, int K, int d, float decay,
                                float eps, void* stream) {
    if (int e = rvq_check_shape("rvq_ema_finalize", nq_use, K, d)) return e;
    if (!cb || !ema_count || !ema_sum || !stats_sum || !stats_cnt) {
        set_error("rvq_ema_finalize: null pointer");
        return RVQ_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // n_tot scratch: a small stream-ordered allocation (nq floats)
    float* ntot = nullptr;
    RVQ_CUDA(scratch_alloc(&ntot, sizeof(float) * nq_use, st));
    const float omd = (float)(1.0 - (double)decay);
    k3_counts<<<nq_use, 1024, 0, st>>>(ema_count, stats_cnt, k_valid, K, decay, omd, ntot);
    const long long warps = (long long)nq_use * K;
    const int block = 256;
    k3_codes<<<(unsigned)((warps * 32 + block - 1) / block), block, 0, st>>>(cb, ema_count, ema_sum, stats_sum, k_valid,
                                                                             ntot, nq_use, K, d, decay, omd, eps);
    RVQ_CUDA(cudaGetLastError());
    RVQ_CUDA(cudaFreeAsync(ntot, st));
    return RVQ_OK;
}

extern "C" int rvq_ema_counts(float* ema_count, const float* stats_cnt, const int* k_valid, int nq_use, int K,
                              float decay, void* stream) {
    if (nq_use <= 0 || K <= 0 || !ema_count || !stats_cnt) {
        set_error("rvq_ema_counts: bad argument");
        return RVQ_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* ntot = nullptr;
    RVQ_CUDA(scratch_alloc(&ntot, sizeof(float) * nq_use, st));
    const float omd = (float)(1.0 - (double)decay);
    k3_counts<<<nq_use, 1024, 0, st>>>(ema_count, stats_cnt, k_valid, K, decay, omd, ntot);
    RVQ_CUDA(cudaGetLastError());
    RVQ_CUDA(cudaFreeAsync(ntot, st));
    return RVQ_OK;
}

extern "C" int rvq_dequantize(const float* cb, const long long* idx, long long N, long long L, long long stride_b,
                              long long stride_l, long long stride_d, int d, int q0, int nq_use, int K, const float* w,
                              int accumulate, float* out, void* stream) {
    if (int e = rvq_check_shape("rvq_dequantize", nq_use, K, d)) return e;
    if (!cb || !idx || !out || N < 0 || L <= 0 || q0 < 0) {
        set_error("rvq_dequantize: bad argument");
        return RVQ_ERR_ARG;
    }
    if (N == 0) return RVQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* wdev = nullptr;
    if (w) {
        RVQ_CUDA(scratch_alloc(&wdev, sizeof(float) * nq_use, st));
        RVQ_CUDA(cudaMemcpyAsync(wdev, w, sizeof(float) * nq_use, cudaMemcpyHostToDevice, st));
    }
    RowAddr ad{L, stride_b, stride_l, stride_d};
    const int block = 256;
    const long long threads = stride_d == 1 ? N * 32 : ((N + 127) / 128) * 128 * d;
    dequantize_rows<<<(unsigned)((threads + block - 1) / block), block, 0, st>>>(cb, idx, N, ad, d, q0, nq_use, K,
                                                                                 make_float4(0, 0, 0, 0), wdev,
                                                                                 accumulate, out);
    RVQ_CUDA(cudaGetLastError());
    if (wdev) RVQ_CUDA(cudaFreeAsync(wdev, st));
    return RVQ_OK;
}

// ------------------------------------------------------------------------------------------ wire format
// Codes of a frame packed LSB-first, `bits` bits per stage, frames byte-aligned: bytes_per_frame =
// ceil(nq * bits / 8) (bits per frame of the codec = nq * log2 K, utils.py:137-147 bitrate_calculator).
// One thread per frame; HBM-bound: reads 8 nq bytes, writes bytes_per_frame (pack) or the reverse (unpack).
__global__ void pack_indices(const long long* __restrict__ idx, long long N, int nq, int bits, int bpf,
                             uint8_t* __restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const long long* in = idx + n * nq;
    uint8_t* o = out + n * bpf;
    unsigned long long acc = 0;
    int have = 0, w = 0;
    const unsigned long long mask = (1ull << bits) - 1ull;
    for (int q = 0; q < nq; ++q) {
        acc |= ((unsigned long long)in[q] & mask) << have;
        have += bits;
        while (have >= 8) {
            o[w++] = (uint8_t)(acc & 0xFFu);
            acc >>= 8;
            have -= 8;
        }
    }
    if (have > 0) o[w] = (uint8_t)(acc & 0xFFu);
}

__global__ void unpack_indices(const uint8_t* __restrict__ in, long long N, int nq, int bits, int bpf,
                               long long* __restrict__ idx) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint8_t* src = in + n * bpf;
    long long* o = idx + n * nq;
    unsigned long long acc = 0;
    int have = 0, r = 0;
    const unsigned long long mask = (1ull << bits) - 1ull;
    for (int q = 0; q < nq; ++q) {
        while (have < bits) {
            acc |= (unsigned long long)src[r++] << have;
            have += 8;
        }
        o[q] = (long long)(acc & mask);
        acc >>= bits;
        have -= bits;
    }
}

static int check_wire(const char* who, long long N, int nq, int bits) {
    if (N < 0 || nq <= 0 || bits < 1 || bits > 32) {
        set_error("%s: bad argument (N=%lld nq=%d bits=%d; 1 <= bits <= 32)", who, N, nq, bits);
        return RVQ_ERR_ARG;
    }
    return RVQ_OK;
}

extern "C" int rvq_packed_bytes_per_frame(int nq, int bits) { return (nq <= 0 || bits < 1 || bits > 32) ? RVQ_ERR_ARG : (nq * bits + 7) / 8; }

extern "C" int rvq_pack_indices(const long long* idx, long long N, int nq, int bits, void* packed, void* stream) {
    if (int e = check_wire("rvq_pack_indices", N, nq, bits)) return e;
    if (N == 0) return RVQ_OK;
    if (!idx || !packed) {
        set_error("rvq_pack_indices: null pointer");
        return RVQ_ERR_ARG;
    }
    const int block = 256;
    pack_indices<<<(unsigned)((N + block - 1) / block), block, 0, static_cast<cudaStream_t>(stream)>>>(
        idx, N, nq, bits, (nq * bits + 7) / 8, static_cast<uint8_t*>(packed));
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}

extern "C" int rvq_unpack_indices(const void* packed, long long N, int nq, int bits, long long* idx, void* stream) {
    if (int e = check_wire("rvq_unpack_indices", N, nq, bits)) return e;
    if (N == 0) return RVQ_OK;
    if (!idx || !packed) {
        set_error("rvq_unpack_indices: null pointer");
        return RVQ_ERR_ARG;
    }
    const int block = 256;
    unpack_indices<<<(unsigned)((N + block - 1) / block), block, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(packed), N, nq, bits, (nq * bits + 7) / 8, idx);
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}

// called from rvq_encode (rvq_abi.cu)
int rvq_launch_exact_scan(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d,
                          int nq, int K, const float* cb, const float* meta, float* xq, long long* idx,
                          double* commit_sq, float* stats_sum, float* stats_cnt, cudaStream_t st) {
    RowAddr ad{L, sb, sl, sd};
    const int block = 256, rows = block / 32;
    const size_t smem = (size_t)rows * d * sizeof(float);
    encode_exact_scan<<<(unsigned)((N + rows - 1) / rows), block, smem, st>>>(x, N, ad, d, nq, K, cb, meta, xq, idx,
                                                                              commit_sq, stats_sum, stats_cnt);
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}
