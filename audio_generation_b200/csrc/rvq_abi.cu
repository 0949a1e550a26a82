// rvq_abi.cu -- extern "C" entry points of librvq_sm100a.so (declared in include/rvq_sm100a.h):
// argument checking, device gate, error state, dispatch to the kernels.  No torch types cross this file.
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace rvq {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return RVQ_ERR_CUDA;
}
}  // namespace rvq
using namespace rvq;

int rvq_check_shape(const char* who, int nq, int K, int d);
int rvq_tc_workspace_bytes(int d, int num_sms, size_t* out);
int rvq_launch_tc(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d, int nq,
                  int K, int q_begin, const float* cb, const void* cb_op, int nq_total, const float* cb_norm,
                  const float* cb_meta, float* xq, long long* idx, double* commit_sq, float* stats_sum,
                  float* stats_cnt, void* ws, size_t ws_bytes, float* dbg_scores, float* dbg_rowscale, int cluster,
                  unsigned long long* prof, cudaStream_t st);
bool rvq_fr_supported(int d);
int rvq_launch_fr(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d, int nq,
                  int K, int q_begin, const float* cb, const void* cb_op, int nq_total, const float* cb_norm,
                  const float* cb_meta, float* xq, long long* idx, double* commit_sq, float* stats_sum,
                  float* stats_cnt, int cluster, unsigned long long* prof, cudaStream_t st);
bool rvq_tr_supported(int d);
int rvq_launch_tr(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d, int nq,
                  int K, int q_begin, const float* cb, const void* cb_op, int nq_total, const float* cb_norm,
                  const float* cb_meta, float* xq, long long* idx, double* commit_sq, float* stats_sum,
                  float* stats_cnt, int cluster, unsigned long long* prof, cudaStream_t st);
int rvq_launch_exact_scan(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d,
                          int nq, int K, const float* cb, const float* meta, float* xq, long long* idx,
                          double* commit_sq, float* stats_sum, float* stats_cnt, cudaStream_t st);

extern "C" int rvq_version(void) { return RVQ_ABI_VERSION; }
extern "C" const char* rvq_last_error(void) { return g_err; }

extern "C" int rvq_device_supported(int device) {
    int major = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    return major == 10 ? 1 : 0;
}

static int require_sm100(const char* who) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    const int ok = rvq_device_supported(dev);
    if (ok < 0) return ok;
    if (!ok) {
        set_error("%s: device %d is not compute capability 10.x (sm_100a kernels only; there is no fallback)", who, dev);
        return RVQ_ERR_ARCH;
    }
    return RVQ_OK;
}

extern "C" int rvq_workspace_bytes(int nq, int K, int d, long long N, size_t* out) {
    if (int e = rvq_check_shape("rvq_workspace_bytes", nq, K, d)) return e;
    if (!out || N < 0) {
        set_error("rvq_workspace_bytes: bad argument");
        return RVQ_ERR_ARG;
    }
    // sized for the largest B200 SM count so that the query needs no device
    return rvq_tc_workspace_bytes(d, 160, out);
}

extern "C" int rvq_encode(const float* x, long long N, long long L, long long stride_b, long long stride_l,
                          long long stride_d, int d, int nq_use, int K, const float* cb, const void* cb_op,
                          const float* cb_norm, const float* cb_meta, float* xq, long long* idx, double* commit_sq,
                          float* stats_sum, float* stats_cnt, void* ws, size_t ws_bytes, int flags, void* stream) {
    if (int e = rvq_check_shape("rvq_encode", nq_use, K, d)) return e;
    if (N == 0) {
        if (!commit_sq) return RVQ_OK;
        RVQ_CUDA(cudaMemsetAsync(commit_sq, 0, sizeof(double) * nq_use, static_cast<cudaStream_t>(stream)));
        return RVQ_OK;
    }
    if (N < 0 || L <= 0 || (N % L) != 0) {
        set_error("rvq_encode: bad frame addressing N=%lld L=%lld", N, L);
        return RVQ_ERR_ARG;
    }
    if (!x || !cb || !cb_norm || !cb_meta || !xq || !idx || !commit_sq) {
        set_error("rvq_encode: null pointer");
        return RVQ_ERR_ARG;
    }
    if ((stats_sum == nullptr) != (stats_cnt == nullptr)) {
        set_error("rvq_encode: stats_sum and stats_cnt must both be given or both be null");
        return RVQ_ERR_ARG;
    }
    if (stride_d == 1 && ((stride_l % 4) != 0 || (stride_b % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) ||
                          (reinterpret_cast<uintptr_t>(xq) & 15))) {
        set_error("rvq_encode: feature-contiguous frames must be 16-byte aligned (stride_l, stride_b multiples of 4)");
        return RVQ_ERR_ARG;
    }
    if (int e = require_sm100("rvq_encode")) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RVQ_CUDA(cudaMemsetAsync(commit_sq, 0, sizeof(double) * nq_use, st));
    const int algo = flags & RVQ_FLAG_ALGO_MASK;
    if (algo == RVQ_ALGO_EXACT_SCAN)
        return rvq_launch_exact_scan(x, N, L, stride_b, stride_l, stride_d, d, nq_use, K, cb, cb_meta, xq, idx,
                                     commit_sq, stats_sum, stats_cnt, st);
    if (algo != RVQ_ALGO_TENSOR) {
        set_error("rvq_encode: unknown algo %d", algo);
        return RVQ_ERR_ARG;
    }
    if (!cb_op) {
        set_error("rvq_encode: cb_op is null");
        return RVQ_ERR_ARG;
    }
    const int kernel = flags & RVQ_FLAG_KERNEL_MASK;
    const int cluster = (flags & RVQ_FLAG_CLUSTER_MASK) >> RVQ_FLAG_CLUSTER_SHIFT;
    if (cluster != 0 && cluster != 1 && cluster != 2 && cluster != 4) {
        set_error("rvq_encode: cluster size must be 1, 2 or 4 (got %d)", cluster);
        return RVQ_ERR_ARG;
    }
    unsigned long long* prof = nullptr;
    if (flags & RVQ_FLAG_COUNTERS) {
        if (!ws || ws_bytes < 256) {
            set_error("rvq_encode: RVQ_FLAG_COUNTERS needs a workspace of at least 256 bytes");
            return RVQ_ERR_WORKSPACE;
        }
        // 32 counters live in the LAST 256 bytes of the workspace
        prof = reinterpret_cast<unsigned long long*>(reinterpret_cast<uintptr_t>(ws) + ((ws_bytes - 256) & ~(size_t)7));
        RVQ_CUDA(cudaMemsetAsync(prof, 0, 256, st));
    }
    switch (kernel) {
        case RVQ_KERNEL_AUTO:
            // the fastest measured kernel per feature dimension (profiles/): d = 64 / 128 the
            // TMEM-resident kernel with separate scan and update warps, larger d the generic kernel
            if (rvq_tr_supported(d))
                return rvq_launch_tr(x, N, L, stride_b, stride_l, stride_d, d, nq_use, K, 0, cb, cb_op, nq_use, cb_norm,
                                     cb_meta, xq, idx, commit_sq, stats_sum, stats_cnt, cluster, prof, st);
            break;
        case RVQ_KERNEL_TMEM:
            if (!rvq_tr_supported(d)) {
                set_error("rvq_encode: RVQ_KERNEL_TMEM supports d = 64, 128 (got %d)", d);
                return RVQ_ERR_ARG;
            }
            return rvq_launch_tr(x, N, L, stride_b, stride_l, stride_d, d, nq_use, K, 0, cb, cb_op, nq_use, cb_norm,
                                 cb_meta, xq, idx, commit_sq, stats_sum, stats_cnt, cluster, prof, st);
        case RVQ_KERNEL_FRAME:
            if (!rvq_fr_supported(d)) {
                set_error("rvq_encode: RVQ_KERNEL_FRAME supports d = 64, 128, 256 (got %d)", d);
                return RVQ_ERR_ARG;
            }
            return rvq_launch_fr(x, N, L, stride_b, stride_l, stride_d, d, nq_use, K, 0, cb, cb_op, nq_use, cb_norm,
                                 cb_meta, xq, idx, commit_sq, stats_sum, stats_cnt, cluster, prof, st);
        case RVQ_KERNEL_GENERIC:
            break;
        default:
            set_error("rvq_encode: unknown kernel selector 0x%x", kernel);
            return RVQ_ERR_ARG;
    }
    // the TMA descriptor spans stages [0, nq_use): later stages are never addressed
    return rvq_launch_tc(x, N, L, stride_b, stride_l, stride_d, d, nq_use, K, 0, cb, cb_op, nq_use, cb_norm, cb_meta,
                         xq, idx, commit_sq, stats_sum, stats_cnt, ws, ws_bytes, nullptr, nullptr, cluster, prof, st);
}

extern "C" int rvq_debug_stage_scores(const float* x, int d, int K, int stage, const void* cb_op, const float* cb_norm,
                                      const float* cb_meta, float* scores, float* row_scale, void* stream) {
    if (int e = rvq_check_shape("rvq_debug_stage_scores", 1, K, d)) return e;
    if (!x || !cb_op || !cb_norm || !cb_meta || !scores || !row_scale || stage < 0) {
        set_error("rvq_debug_stage_scores: bad argument");
        return RVQ_ERR_ARG;
    }
    if (int e = require_sm100("rvq_debug_stage_scores")) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // scratch: fp32 master codebook is not needed for the filter itself, but the kernel finishes the stage
    // (gather + update), so the caller-visible outputs go to throw-away buffers
    const int Kpad = round_up(K, CHUNK_N);
    (void)Kpad;
    float* tmp = nullptr;
    const size_t n_xq = (size_t)TILE_M * d;
    size_t ws_bytes = 0;
    rvq_tc_workspace_bytes(d, 1, &ws_bytes);
    const size_t bytes = n_xq * 4 + TILE_M * 8 + 64 + ws_bytes + (size_t)(stage + 1) * K * d * 4;
    RVQ_CUDA(cudaMallocAsync(&tmp, bytes, st));
    RVQ_CUDA(cudaMemsetAsync(tmp, 0, bytes, st));
    uint8_t* b = reinterpret_cast<uint8_t*>(tmp);
    float* xq = reinterpret_cast<float*>(b);
    long long* idx = reinterpret_cast<long long*>(b + n_xq * 4);
    double* csq = reinterpret_cast<double*>(b + n_xq * 4 + TILE_M * 8);
    void* ws = b + n_xq * 4 + TILE_M * 8 + 64;
    float* fake_cb = reinterpret_cast<float*>(b + n_xq * 4 + TILE_M * 8 + 64 + ws_bytes);  // zeros: gather is harmless
    int rc = rvq_launch_tc(x, TILE_M, TILE_M, 0, d, 1, d, 1, K, stage, fake_cb, cb_op, stage + 1, cb_norm, cb_meta, xq,
                           idx, csq, nullptr, nullptr, ws, ws_bytes, scores, row_scale, 1, nullptr, st);
    cudaFreeAsync(tmp, st);
    return rc;
}
