/*
 * rvq_sm100a.h -- C ABI of librvq_sm100a.so: the B200 (sm_100a) residual-vector-quantization
 * hot path that replaces the body of `som_quantizer.ResidualQuantizer.forward`
 * (third-party quantization-maps package imported at /root/reference/networks/vae.py:6 and
 * called at /root/reference/networks/vae.py:315-318).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked HOST;
 *   - every entry enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising the host (except where noted);
 *   - returns 0 on success, a negative rvq_status on failure; the message of the last failure
 *     on the calling thread is returned by rvq_last_error(); nothing throws across the ABI;
 *   - there is no CPU path: a device that is not compute capability 10.x is refused
 *     (RVQ_ERR_ARCH).
 *
 * Shapes: nq stages, K codes per stage, d features, N frames.  Kpad = K rounded up to 256.
 */
#ifndef RVQ_SM100A_H
#define RVQ_SM100A_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVQ_ABI_VERSION 7

typedef enum {
    RVQ_OK = 0,
    RVQ_ERR_ARG = -1,      /* bad shape / null pointer / unsupported size            */
    RVQ_ERR_ARCH = -2,     /* current device is not sm_100                           */
    RVQ_ERR_CUDA = -3,     /* CUDA runtime / driver error (launch, attribute, ...)   */
    RVQ_ERR_WORKSPACE = -4 /* workspace too small                                    */
} rvq_status;

/* encode flags */
#define RVQ_ALGO_TENSOR 0      /* tcgen05 fp16 filter + exact fp32 re-rank (product path)       */
#define RVQ_ALGO_EXACT_SCAN 1  /* exact fp32 CUDA-core scan of every code (verification, tiny N) */
#define RVQ_FLAG_ALGO_MASK 0xF
/* kernel choice of RVQ_ALGO_TENSOR (bits 4-7).  All kernels share one exact scorer and return bit-identical
 * codes; AUTO is the product path (the fastest measured kernel per d), the others let tests cross-check it. */
#define RVQ_KERNEL_AUTO 0x00     /* TMEM kernel for d = 64 / 128, generic kernel otherwise                          */
#define RVQ_KERNEL_GENERIC 0x10  /* rvq_encode_tc.cu: residual in an L2-resident scratch, any d                     */
#define RVQ_KERNEL_FRAME 0x20    /* rvq_encode_fr.cu: thread = frame, residual in tensor memory, d = 64 / 128 / 256 */
#define RVQ_KERNEL_TMEM 0x30     /* rvq_encode_tr.cu: residual in tensor memory, scan and update warps, d = 64 / 128 */
#define RVQ_FLAG_KERNEL_MASK 0xF0
/* CTAs per thread-block cluster (bits 8-10): 0 = the kernel's default (1: independent CTAs, the fastest measured),
 * else 1, 2 or 4.  RVQ_KERNEL_TMEM: 2 = the two CTAs drive their tensor cores as ONE tcgen05.mma.cta_group::2
 * instruction stream (B split between their shared memories), 4 = one codebook stream multicast to four CTAs;
 * RVQ_KERNEL_FRAME / GENERIC: 2, 4 = multicast.  Results are bit-identical for every value. */
#define RVQ_FLAG_CLUSTER_SHIFT 8
#define RVQ_FLAG_CLUSTER_MASK 0x700
/* event counters of the launch in the LAST 256 bytes of ws (uint64[32], zeroed by the call): test / profiling aid */
#define RVQ_FLAG_COUNTERS 0x1000

int rvq_version(void);
const char* rvq_last_error(void);
/* 1 if `device` (ordinal) can run the kernels (compute capability 10.x), 0 if not, <0 on error. HOST. */
int rvq_device_supported(int device);

/* Sizes (HOST pointers out) of the derived codebook operands written by rvq_prepare_codebooks:
 *   op_bytes   : fp16 [nq, Kpad, d]  = -2 * 2^b_q * C_q            (UMMA B operand, K-major rows)
 *   norm_bytes : fp32 [nq, Kpad]     = 2^(2 b_q) * ||c||^2 (minus the folded allowance term of the codes above the
 *                stage's norm cap), padding codes = 2^100; followed by the same norms as fp16 UMMA operand slices
 *                [nq, Kpad/128, 4096 bytes] (three exact 11-bit pieces per code + the allowance column), the first
 *                allowance factor fp32 [nq, Kpad], the allowance byte table [nq, Kpad] and one int flag per
 *                256-code chunk (opaque to the caller: written by rvq_prepare_codebooks, read by rvq_encode;
 *                DESIGN.md section 3 "Per-code bound")
 *   meta_bytes : fp32 [nq, 8]        = {2^b_q, norm cap / 2^b_q, max |c|, K_valid, nq prepared, U2, max_k ||c_k||_2, U1} */
int rvq_prepared_bytes(int nq, int K, int d, size_t* op_bytes, size_t* norm_bytes, size_t* meta_bytes);

/* K0. Derive the tensor-core operands from the fp32 master codebooks cb[nq, K, d].
 * k_valid (DEVICE int[nq], nullable): number of real codes per stage (<= K); codes beyond it
 * can never be selected.  Must be re-run after every codebook change
 * (replaces nothing in the reference: the reference uses the fp32 codebook directly). */
int rvq_prepare_codebooks(const float* cb, const int* k_valid, int nq, int K, int d,
                          void* cb_op, float* cb_norm, float* cb_meta, void* stream);

/* Scratch needed by rvq_encode for N frames. HOST pointer out. */
int rvq_workspace_bytes(int nq, int K, int d, long long N, size_t* out);

/* K1 (+K2 fused). Replaces the per-stage loop of ResidualQuantizer.forward
 * (distance -> argmin -> gather -> residual subtract -> [EMA statistics]) called at
 * /root/reference/networks/vae.py:315-318.
 *
 *   x        fp32, N = (N / L) * L frames addressed as x[(n / L) * stride_b + (n % L) * stride_l + i * stride_d]
 *            (the reference passes a (B, L, d) VIEW of a (B, d, L) tensor: stride_l = 1, stride_d = L;
 *            contiguous (N, d): L = N, stride_b = 0, stride_l = d, stride_d = 1)
 *   cb       fp32 master codebooks [nq_total, K, d]; stages 0..nq_use-1 are used
 *   cb_op / cb_norm / cb_meta   outputs of rvq_prepare_codebooks for the same cb
 *   xq       fp32 out, same addressing as x: sum of the selected code vectors (= x - final residual)
 *   idx      int64 out [N, nq_use] (what torch.nn.functional.one_hot needs, utils.py:253)
 *   commit_sq  fp64 out [nq_use]: sum over frames and features of (r_q - z_q)^2 per stage
 *              (commit loss of stage q = commit_sq[q] / (N*d)); zeroed by the call
 *   stats_sum  fp32 [nq_total, K, d] nullable, stats_cnt fp32 [nq_total, K] nullable: EMA statistics,
 *              ACCUMULATED into (caller zeroes): cnt[q,k] += #{n: idx=k}, sum[q,k,:] += r_q[n,:]
 *   ws / ws_bytes   scratch of at least rvq_workspace_bytes
 *   flags    RVQ_ALGO_* | RVQ_KERNEL_* | cluster size << RVQ_FLAG_CLUSTER_SHIFT | RVQ_FLAG_COUNTERS      */
int rvq_encode(const float* x, long long N, long long L, long long stride_b, long long stride_l,
               long long stride_d, int d, int nq_use, int K,
               const float* cb, const void* cb_op, const float* cb_norm, const float* cb_meta,
               float* xq, long long* idx, double* commit_sq,
               float* stats_sum, float* stats_cnt,
               void* ws, size_t ws_bytes, int flags, void* stream);

/* K3. EMA refresh of one call's statistics (after the cross-GPU all-reduce of stats_*):
 *   ema_count = decay*ema_count + (1-decay)*cnt;  ema_sum = decay*ema_sum + (1-decay)*sum;
 *   cb[k] = ema_sum[k] / ((ema_count[k]+eps)/(n+K_valid*eps)*n),  n = sum_k ema_count[k]
 * over stages 0..nq_use-1.  k_valid as in rvq_prepare_codebooks. */
int rvq_ema_finalize(float* cb, float* ema_count, float* ema_sum,
                     const float* stats_sum, const float* stats_cnt, const int* k_valid,
                     int nq_use, int K, int d, float decay, float eps, void* stream);

/* The count half of K3 alone: ema_count = decay*ema_count + (1-decay)*cnt over stages 0..nq_use-1.  For
 * quantizer_class "base" (/root/reference/config/training.yml:21), whose codebooks are trained by gradient: the
 * usage statistics still feed get_stale_clusters() / the stale-code re-seeding (training.py:435,461). */
int rvq_ema_counts(float* ema_count, const float* stats_cnt, const int* k_valid, int nq_use, int K, float decay,
                   void* stream);

/* Code lookup (ResidualQuantizer.quantizers[i].dequantize, /root/reference/networks/vae.py:333,
 * summed over stages as CausalVQAE.sample does at vae.py:329-334):
 *   out[n,:] (+)= sum_q w[q] * cb[q0+q, idx[n,q], :]   (w nullable = all ones; HOST float[nq_use])
 * out addressed like x in rvq_encode.  accumulate != 0 adds into out. */
int rvq_dequantize(const float* cb, const long long* idx, long long N, long long L,
                   long long stride_b, long long stride_l, long long stride_d,
                   int d, int q0, int nq_use, int K, const float* w, int accumulate,
                   float* out, void* stream);

/* Wire format of the codes (section 8f of SURVEY.md; bits per frame = nq * log2 K as in
 * /root/reference/networks/utils.py:137-147 bitrate_calculator): the nq codes of a frame packed LSB-first,
 * `bits` bits each (1..32), frames byte-aligned: rvq_packed_bytes_per_frame = ceil(nq * bits / 8).
 *   idx    int64 [N, nq] (what rvq_encode writes / rvq_dequantize reads)
 *   packed uint8 [N, bytes_per_frame]                                                          */
int rvq_packed_bytes_per_frame(int nq, int bits); /* HOST; < 0 on bad arguments */
int rvq_pack_indices(const long long* idx, long long N, int nq, int bits, void* packed, void* stream);
int rvq_unpack_indices(const void* packed, long long N, int nq, int bits, long long* idx, void* stream);

/* Self-organising-map neighbourhood of the EMA statistics (use_som / som_kernel_type of the reference's
 * constructor, /root/reference/networks/vae.py:220-221,250-251; the map of stage q is a height x width grid with
 * height * width = number of codes, /root/reference/networks/utils.py:244-245,257).  Runs after the cross-GPU
 * all-reduce and before rvq_ema_finalize:
 *   out[q, (y, x)] = sum_{dy, dx in [-radius, radius]} weights[dy + radius][dx + radius] * in[q, (y + dy, x + dx)]
 * neighbours outside the grid are skipped (no wrap-around), codes beyond height * width pass through.  The terms
 * are added in row-major (dy, dx) order with separate fp32 multiply and add; zero weights are skipped.
 *   grid_hw HOST int[2 * nq_use] = {height, width} per stage;  weights HOST float[(2 radius + 1)^2];
 *   radius <= 4, nq_use <= 64;  out_* must not alias stats_*.                                           */
int rvq_som_spread(const float* stats_sum, const float* stats_cnt, float* out_sum, float* out_cnt,
                   const int* grid_hw, int nq_use, int K, int d, int radius, const float* weights, void* stream);

/* Stale-code re-seeding (vq_cutoff_freq of the reference's constructor, vae.py:213,249; get_stale_clusters /
 * update_cutoff, /root/reference/networks/training.py:435,454,461).
 * rvq_reseed_frame (HOST, pure): the global frame whose stage-q residual re-seeds code (q, k):
 *   z = seed + (q K + k + 1) * 0x9E3779B97F4A7C15;  z = (z ^ z >> 30) * 0xBF58476D1CE4E5B9;
 *   z = (z ^ z >> 27) * 0x94D049BB133111EB;  z ^= z >> 31;  frame = z mod frames_total       (64-bit wrap-around)
 * rvq_reseed_gather: rep[q, k, :] = x[n] - sum_{s < q} cb[s, idx[n, s], :] (fp32, stage order: the residual the
 *   encode kernel saw) for n = frame - frame_offset when 0 <= n < N, zeros otherwise - so that a SUM all-reduce of
 *   rep over ranks holding disjoint frame ranges leaves the owner's vector on every replica.  Must run BEFORE
 *   rvq_ema_finalize (it needs the codebooks idx was computed with).  ema_count (nullable): codes with
 *   decay * ema_count >= cutoff cannot become stale in this step and get zeros without touching x.
 * rvq_reseed_apply (after rvq_ema_finalize): every valid code with ema_count < cutoff takes cb = rep,
 *   ema_sum = rep * reset_count, ema_count = reset_count;  n_replaced DEVICE int[nq_use] (nullable) = how many.  */
unsigned long long rvq_reseed_frame(unsigned long long seed, int q, int K, int k, unsigned long long frames_total);
int rvq_reseed_gather(const float* x, long long N, long long L, long long stride_b, long long stride_l,
                      long long stride_d, int d, int nq_use, int K, const float* cb, const long long* idx,
                      const float* ema_count, float decay, float cutoff, unsigned long long seed,
                      long long frame_offset, long long frames_total, float* rep, void* stream);
int rvq_reseed_apply(float* cb, float* ema_count, float* ema_sum, const float* rep, const int* k_valid,
                     int nq_use, int K, int d, float cutoff, float reset_count, int* n_replaced, void* stream);

/* Backward of the quantizer call for autograd (the loss is taken on the decoder output plus the commit loss,
 * /root/reference/networks/training.py:336,344-346; codebooks are parameters for quantizer_class "base",
 * /root/reference/config/training.yml:21).  With coef = *g_commit * 2 / (N d) and r_{q+1} the residual after stage q
 * (re-walked from x, idx and cb in fp32 stage order):
 *   gx[n, :]            = g_out[n, :] + coef * w_commit * sum_q r_{q+1}[n, :]      (straight-through + commit loss)
 *   gcb[q, idx[n,q], :] += -coef * w_codebook * r_{q+1}[n, :]                       (codebook loss; caller zeroes)
 * x, g_out (nullable = zeros), gx (nullable) share the frame addressing of rvq_encode; g_commit is a DEVICE scalar
 * (nullable = 0: no host synchronisation); gcb [nq_use.., K, d] nullable.                                     */
int rvq_backward(const float* x, long long N, long long L, long long stride_b, long long stride_l,
                 long long stride_d, int d, int nq_use, int K, const float* cb, const long long* idx,
                 const float* g_out, const float* g_commit, float w_commit, float w_codebook, float* gx, float* gcb,
                 void* stream);

/* Bring-up / test hook: run ONE stage of the tensor-core filter for the first 128 frames of x
 * (contiguous [128, d]) and write the approximate scaled scores fp32 [128, Kpad] and the per-row
 * scale 2^a [128].  Not used by the product path. */
int rvq_debug_stage_scores(const float* x, int d, int K, int stage,
                           const void* cb_op, const float* cb_norm, const float* cb_meta,
                           float* scores, float* row_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RVQ_SM100A_H */
