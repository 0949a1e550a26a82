// rvq_codebook.cu -- codebook maintenance around the EMA update:
//   som_spread      self-organising-map neighbourhood: the statistics of a code also flow to its grid neighbours
//                   (use_som / som_kernel_type, /root/reference/networks/vae.py:220-221,250-251; grid shape
//                   /root/reference/networks/utils.py:244-245,257)
//   reseed_gather   replacement vectors for stale codes: the stage-q residual of one pseudo-randomly chosen frame
//   reseed_apply    codes whose EMA count fell below the cutoff take the replacement (vq_cutoff_freq,
//                   vae.py:213,249; get_stale_clusters / update_cutoff, training.py:435,454,461)
// All three are HBM/L2-bound passes over [nq, K, d(+1)] floats; they run between the statistics all-reduce and the
// next encode, once per update step.
#include "common.cuh"

namespace rvq {

constexpr int SOM_MAX_RADIUS = 4;
constexpr int SOM_MAX_STAGES = 64;

struct SomParams {
    float w[(2 * SOM_MAX_RADIUS + 1) * (2 * SOM_MAX_RADIUS + 1)];
    short h[SOM_MAX_STAGES], wd[SOM_MAX_STAGES];
    int radius;
};

// One warp per (stage, code): out[k] = sum over the (2r+1)^2 window, in row-major (dy, dx) order, of
// w[dy][dx] * in[neighbour], neighbours outside the grid skipped, zero weights skipped; separate fp32 multiply and
// add (no contraction) so that a CPU restatement with the same loop order is bit-identical.
__global__ void som_spread(const float* __restrict__ in_sum, const float* __restrict__ in_cnt,
                           float* __restrict__ out_sum, float* __restrict__ out_cnt, int nq, int K, int d,
                           const __grid_constant__ SomParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nq * K) return;
    const int q = warp / K, k = warp % K;
    const int H = p.h[q], W = p.wd[q], r = p.radius, side = 2 * r + 1;
    const size_t stage = (size_t)q * K;
    if (k >= H * W) {  // beyond the map (padding codes): statistics pass through
        for (int i = lane * 4; i < d; i += 128)
            *reinterpret_cast<float4*>(out_sum + (stage + k) * d + i) =
                *reinterpret_cast<const float4*>(in_sum + (stage + k) * d + i);
        if (lane == 0) out_cnt[stage + k] = in_cnt[stage + k];
        return;
    }
    const int y = k / W, x = k % W;
    float c = 0.f;
    for (int i0 = 0; i0 < d; i0 += 128) {
        const int i = i0 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int dy = -r; dy <= r; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -r; dx <= r; ++dx) {
                const int xx = x + dx;
                const float wgt = p.w[(dy + r) * side + (dx + r)];
                if (xx < 0 || xx >= W || wgt == 0.f) continue;
                const size_t j = stage + (size_t)yy * W + xx;
                if (i < d) {
                    const float4 v = *reinterpret_cast<const float4*>(in_sum + j * d + i);
                    acc.x = __fadd_rn(acc.x, __fmul_rn(wgt, v.x));
                    acc.y = __fadd_rn(acc.y, __fmul_rn(wgt, v.y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(wgt, v.z));
                    acc.w = __fadd_rn(acc.w, __fmul_rn(wgt, v.w));
                }
                if (i0 == 0 && lane == 0) c = __fadd_rn(c, __fmul_rn(wgt, in_cnt[j]));
            }
        }
        if (i < d) *reinterpret_cast<float4*>(out_sum + (stage + k) * d + i) = acc;
    }
    if (lane == 0) out_cnt[stage + k] = c;
}

struct RowAddrC {
    long long L, sb, sl, sd;
    __device__ __forceinline__ long long row(long long n) const { return (n / L) * sb + (n % L) * sl; }
};

// the frame whose stage-q residual re-seeds code (q, k): splitmix64 of seed + (q K + k + 1) * golden, mod frames
__host__ __device__ __forceinline__ unsigned long long reseed_frame(unsigned long long seed, int q, int K, int k,
                                                                    unsigned long long frames_total) {
    unsigned long long z = seed + (unsigned long long)((long long)q * K + k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z % frames_total;
}

// One warp per (stage, code).  Candidates are drawn from the GLOBAL frame range [0, frames_total); a rank holding
// frames [frame_offset, frame_offset + N) writes the vector, every other rank writes zeros, so a SUM all-reduce of
// `rep` leaves the owner's vector bit for bit on every replica.  Only codes that can end up below the cutoff are
// gathered: ema_count_new = decay * ema_count + (1 - decay) * cnt >= decay * ema_count.
__global__ void reseed_gather(const float* __restrict__ x, long long N, RowAddrC ad, int d, int nq, int K,
                              const float* __restrict__ cb, const long long* __restrict__ idx,
                              const float* __restrict__ ema_count, float decay, float cutoff,
                              unsigned long long seed, long long frame_offset, unsigned long long frames_total,
                              float* __restrict__ rep) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nq * K) return;
    const int q = warp / K, k = warp % K;
    float* out = rep + ((size_t)q * K + k) * d;
    const long long n = (long long)reseed_frame(seed, q, K, k, frames_total) - frame_offset;
    const bool may_be_stale = !ema_count || __fmul_rn(decay, ema_count[(size_t)q * K + k]) < cutoff;
    if (n < 0 || n >= N || !may_be_stale) {
        for (int i = lane; i < d; i += 32) out[i] = 0.f;
        return;
    }
    const float* xr = x + ad.row(n);
    const long long* code = idx + n * nq;
    for (int i = lane; i < d; i += 32) {
        float v = xr[(long long)i * ad.sd];
        for (int s = 0; s < q; ++s) v = __fsub_rn(v, cb[((size_t)s * K + (size_t)code[s]) * d + i]);
        out[i] = v;
    }
}

__global__ void reseed_apply(float* __restrict__ cb, float* __restrict__ ema_count, float* __restrict__ ema_sum,
                             const float* __restrict__ rep, const int* __restrict__ k_valid, int nq, int K, int d,
                             float cutoff, float reset_count, int* __restrict__ n_replaced) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nq * K) return;
    const int q = warp / K, k = warp % K;
    const int Kv = k_valid ? min(max(k_valid[q], 0), K) : K;
    if (k >= Kv) return;
    const size_t c = (size_t)q * K + k;
    if (!(ema_count[c] < cutoff)) return;
    for (int i = lane; i < d; i += 32) {
        const float v = rep[c * d + i];
        cb[c * d + i] = v;
        ema_sum[c * d + i] = __fmul_rn(v, reset_count);
    }
    __syncwarp();
    if (lane == 0) {
        ema_count[c] = reset_count;
        if (n_replaced) atomicAdd(n_replaced + q, 1);
    }
}

}  // namespace rvq

using namespace rvq;
int rvq_check_shape(const char* who, int nq, int K, int d);

extern "C" int rvq_som_spread(const float* stats_sum, const float* stats_cnt, float* out_sum, float* out_cnt,
                              const int* grid_hw, int nq_use, int K, int d, int radius, const float* weights,
                              void* stream) {
    if (int e = rvq_check_shape("rvq_som_spread", nq_use, K, d)) return e;
    if (!stats_sum || !stats_cnt || !out_sum || !out_cnt || !grid_hw || !weights || stats_sum == out_sum ||
        stats_cnt == out_cnt) {
        set_error("rvq_som_spread: null pointer or in-place call (the stencil needs separate output buffers)");
        return RVQ_ERR_ARG;
    }
    if (radius < 0 || radius > SOM_MAX_RADIUS || nq_use > SOM_MAX_STAGES) {
        set_error("rvq_som_spread: radius must be in [0, %d] and nq <= %d (got %d, %d)", SOM_MAX_RADIUS,
                  SOM_MAX_STAGES, radius, nq_use);
        return RVQ_ERR_ARG;
    }
    SomParams p{};
    p.radius = radius;
    const int side = 2 * radius + 1;
    for (int i = 0; i < side * side; ++i) p.w[i] = weights[i];
    for (int q = 0; q < nq_use; ++q) {
        const int h = grid_hw[2 * q], w = grid_hw[2 * q + 1];
        if (h <= 0 || w <= 0 || (long long)h * w > K || h > 32767 || w > 32767) {
            set_error("rvq_som_spread: stage %d grid %d x %d does not fit K=%d", q, h, w, K);
            return RVQ_ERR_ARG;
        }
        p.h[q] = (short)h;
        p.wd[q] = (short)w;
    }
    const long long warps = (long long)nq_use * K;
    const int block = 256;
    som_spread<<<(unsigned)((warps * 32 + block - 1) / block), block, 0, static_cast<cudaStream_t>(stream)>>>(
        stats_sum, stats_cnt, out_sum, out_cnt, nq_use, K, d, p);
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}

extern "C" unsigned long long rvq_reseed_frame(unsigned long long seed, int q, int K, int k,
                                               unsigned long long frames_total) {
    return frames_total ? reseed_frame(seed, q, K, k, frames_total) : 0ull;
}

extern "C" int rvq_reseed_gather(const float* x, long long N, long long L, long long stride_b, long long stride_l,
                                 long long stride_d, int d, int nq_use, int K, const float* cb, const long long* idx,
                                 const float* ema_count, float decay, float cutoff, unsigned long long seed,
                                 long long frame_offset, long long frames_total, float* rep, void* stream) {
    if (int e = rvq_check_shape("rvq_reseed_gather", nq_use, K, d)) return e;
    if (!rep || !cb || frames_total <= 0 || frame_offset < 0 || N < 0 || (N > 0 && (!x || !idx || L <= 0))) {
        set_error("rvq_reseed_gather: bad argument");
        return RVQ_ERR_ARG;
    }
    RowAddrC ad{L > 0 ? L : 1, stride_b, stride_l, stride_d};
    const long long warps = (long long)nq_use * K;
    const int block = 256;
    reseed_gather<<<(unsigned)((warps * 32 + block - 1) / block), block, 0, static_cast<cudaStream_t>(stream)>>>(
        x, N, ad, d, nq_use, K, cb, idx, ema_count, decay, cutoff, seed, frame_offset,
        (unsigned long long)frames_total, rep);
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}

extern "C" int rvq_reseed_apply(float* cb, float* ema_count, float* ema_sum, const float* rep, const int* k_valid,
                                int nq_use, int K, int d, float cutoff, float reset_count, int* n_replaced,
                                void* stream) {
    if (int e = rvq_check_shape("rvq_reseed_apply", nq_use, K, d)) return e;
    if (!cb || !ema_count || !ema_sum || !rep) {
        set_error("rvq_reseed_apply: null pointer");
        return RVQ_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_replaced) RVQ_CUDA(cudaMemsetAsync(n_replaced, 0, sizeof(int) * nq_use, st));
    const long long warps = (long long)nq_use * K;
    const int block = 256;
    reseed_apply<<<(unsigned)((warps * 32 + block - 1) / block), block, 0, st>>>(cb, ema_count, ema_sum, rep, k_valid,
                                                                                nq_use, K, d, cutoff, reset_count,
                                                                                n_replaced);
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}
