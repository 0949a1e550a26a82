// rvq_backward.cu -- backward of the quantizer call: what autograd needs from
// `x_quantized, index, commit_loss = quantizer(x, ...)` (/root/reference/networks/vae.py:315-318) when the loss is
// taken on the decoder output plus the commit loss (/root/reference/networks/training.py:336,344-346):
//   straight-through:  d x_quantized / d x = I
//   commit loss        sum_q mean((r_q - sg z_q)^2)  ->  d/dx       = (2 / (N d)) sum_q r_{q+1}
//   codebook loss      sum_q mean((sg r_q - z_q)^2)  ->  d/dC_q[k]  = -(2 / (N d)) sum_{n: idx[n,q] = k} r_{q+1}[n]
//                      ("base" quantizer class: codebooks are parameters, config/training.yml:21)
// One pass over x and idx re-walks the residual chain (r_{q+1} = r_q - C_q[idx], fp32, stage order: the values
// the encode kernel saw) - HBM-bound: reads 4d (+4d for g_out) bytes per frame and 8 nq of indices, writes 4d;
// the codebook gradient is nq * d fp32 reductions per frame into an L2-resident [nq, K, d] buffer.
#include "common.cuh"

namespace rvq {

struct RowAddrB {
    long long L, sb, sl, sd;
    __device__ __forceinline__ long long row(long long n) const { return (n / L) * sb + (n % L) * sl; }
};

// features contiguous (stride_d == 1): one warp per frame, float4 per lane
__global__ void backward_rows(const float* __restrict__ x, long long N, RowAddrB ad, int d, int nq, int K,
                              const float* __restrict__ cb, const long long* __restrict__ idx,
                              const float* __restrict__ g_out, const float* __restrict__ g_commit, float scale,
                              float w_commit, float w_codebook, float* __restrict__ gx, float* __restrict__ gcb) {
    const long long n = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float coef = g_commit ? __fmul_rn(*g_commit, scale) : 0.f;
    const float cx = __fmul_rn(coef, w_commit), cc = -__fmul_rn(coef, w_codebook);
    const long long off = ad.row(n);
    for (int i = lane * 4; i < d; i += 128) {
        float4 r = *reinterpret_cast<const float4*>(x + off + i);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < nq; ++q) {
            long long k = idx[n * nq + q];
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);
            const size_t c = ((size_t)q * K + (size_t)k) * d + i;
            const float4 cv = *reinterpret_cast<const float4*>(cb + c);
            r.x = __fsub_rn(r.x, cv.x);
            r.y = __fsub_rn(r.y, cv.y);
            r.z = __fsub_rn(r.z, cv.z);
            r.w = __fsub_rn(r.w, cv.w);
            acc.x = __fadd_rn(acc.x, r.x);
            acc.y = __fadd_rn(acc.y, r.y);
            acc.z = __fadd_rn(acc.z, r.z);
            acc.w = __fadd_rn(acc.w, r.w);
            if (gcb) {
                float* g = gcb + c;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(__fmul_rn(r.x, cc)),
                             "f"(__fmul_rn(r.y, cc)), "f"(__fmul_rn(r.z, cc)), "f"(__fmul_rn(r.w, cc))
                             : "memory");
            }
        }
        if (gx) {
            float4 g = make_float4(__fmul_rn(acc.x, cx), __fmul_rn(acc.y, cx), __fmul_rn(acc.z, cx), __fmul_rn(acc.w, cx));
            if (g_out) {
                const float4 go = *reinterpret_cast<const float4*>(g_out + off + i);
                g.x = __fadd_rn(go.x, g.x);
                g.y = __fadd_rn(go.y, g.y);
                g.z = __fadd_rn(go.z, g.z);
                g.w = __fadd_rn(go.w, g.w);
            }
            *reinterpret_cast<float4*>(gx + off + i) = g;
        }
    }
}

// frames fastest (the reference's (B, L, d) view of (B, d, L) storage): thread = (frame, feature), consecutive
// threads = consecutive frames of a 128-frame group
__global__ void backward_cols(const float* __restrict__ x, long long N, RowAddrB ad, int d, int nq, int K,
                              const float* __restrict__ cb, const long long* __restrict__ idx,
                              const float* __restrict__ g_out, const float* __restrict__ g_commit, float scale,
                              float w_commit, float w_codebook, float* __restrict__ gx, float* __restrict__ gcb) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long chunk = 128;
    const long long g = t / (chunk * d), rem = t % (chunk * d);
    const int i = (int)(rem / chunk);
    const long long n = g * chunk + rem % chunk;
    if (n >= N) return;
    const float coef = g_commit ? __fmul_rn(*g_commit, scale) : 0.f;
    const float cx = __fmul_rn(coef, w_commit), cc = -__fmul_rn(coef, w_codebook);
    const long long off = ad.row(n) + (long long)i * ad.sd;
    float r = x[off], acc = 0.f;
    for (int q = 0; q < nq; ++q) {
        long long k = idx[n * nq + q];
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        const size_t c = ((size_t)q * K + (size_t)k) * d + i;
        r = __fsub_rn(r, cb[c]);
        acc = __fadd_rn(acc, r);
        if (gcb) atomicAdd(gcb + c, __fmul_rn(r, cc));
    }
    if (gx) {
        float gv = __fmul_rn(acc, cx);
        if (g_out) gv = __fadd_rn(g_out[off], gv);
        gx[off] = gv;
    }
}

}  // namespace rvq

using namespace rvq;
int rvq_check_shape(const char* who, int nq, int K, int d);

extern "C" int rvq_backward(const float* x, long long N, long long L, long long stride_b, long long stride_l,
                            long long stride_d, int d, int nq_use, int K, const float* cb, const long long* idx,
                            const float* g_out, const float* g_commit, float w_commit, float w_codebook, float* gx,
                            float* gcb, void* stream) {
    if (int e = rvq_check_shape("rvq_backward", nq_use, K, d)) return e;
    if (N == 0) return RVQ_OK;
    if (!x || !cb || !idx || N < 0 || L <= 0 || (N % L) != 0 || (!gx && !gcb)) {
        set_error("rvq_backward: bad argument");
        return RVQ_ERR_ARG;
    }
    if (stride_d == 1 && ((stride_l % 4) != 0 || (stride_b % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) ||
                          (reinterpret_cast<uintptr_t>(gx) & 15) || (reinterpret_cast<uintptr_t>(g_out) & 15))) {
        set_error("rvq_backward: feature-contiguous frames must be 16-byte aligned (stride_l, stride_b multiples of 4)");
        return RVQ_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RowAddrB ad{L, stride_b, stride_l, stride_d};
    const float scale = (float)(2.0 / ((double)N * (double)d));
    const int block = 256;
    if (stride_d == 1) {
        const long long threads = N * 32;
        backward_rows<<<(unsigned)((threads + block - 1) / block), block, 0, st>>>(
            x, N, ad, d, nq_use, K, cb, idx, g_out, g_commit, scale, w_commit, w_codebook, gx, gcb);
    } else {
        const long long threads = ((N + 127) / 128) * 128 * d;
        backward_cols<<<(unsigned)((threads + block - 1) / block), block, 0, st>>>(
            x, N, ad, d, nq_use, K, cb, idx, g_out, g_commit, scale, w_commit, w_codebook, gx, gcb);
    }
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}
