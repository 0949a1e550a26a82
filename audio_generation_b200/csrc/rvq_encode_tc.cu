// rvq_encode_tc.cu -- K1 (+K2 fused): the residual-vector-quantization stage loop as ONE persistent
// sm_100a kernel.  Replaces the per-stage {distance, argmin, gather, subtract, EMA statistics} loop of
// som_quantizer.ResidualQuantizer.forward (called at /root/reference/networks/vae.py:315-318).
//
// Per CTA (one per SM, persistent over tiles of up to 128 frames, two tiles in flight while shared memory allows),
// per stage q:
//   filter   scores~[128 x K] = (2^a r) . (-2 * 2^b C_q)^T on tcgen05 (fp16 operands, fp32 accumulate in
//            TMEM, 256 codes per accumulator buffer, two buffers), B streamed by TMA through an mbarrier ring;
//            the residual operand A is produced by the CTA itself in the UMMA SWIZZLE_128B K-major layout.
//   argmin   scan warps read the accumulators with tcgen05.ld, add 2^(a-b) * (2^(2b)||c||^2) (minus the allowance
//            of codes above the stage's norm cap) and keep a two-dimensional running minimum per frame: the minimum
//            of each of the 16 score columns and the three smallest load minima (encode_common.cuh); nothing
//            N x K ever reaches HBM.
//   certify  every code whose approximate score is within the proven error bound of the best
//            is re-scored exactly in fp32 (exact.cuh), by candidate pairs; if a fourth load is in reach the frame's
//            columns in reach are scanned exactly.  The selected index is therefore the exact fp32 argmin.
//   update   r <- r - C_q[k] in fp32 (residual tile in an L2-resident per-CTA scratch), EMA statistics by
//            red.global.add.v4.f32, commit-loss partials, next stage's fp16 operand written back into the A tile;
//            software-pipelined passes of 32 frames, 8 lanes per frame.
//
// Warp roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-11 = two scan
// groups (accumulator buffers 0 / 1), warps 12-19 = update warps (both tile slots).  Small calls run partly filled
// tiles over more SMs (tile_rows / a_rows, rvq_launch_tc).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "exact.cuh"
#include "encode_common.cuh"

namespace rvq {

constexpr int SCAN_WARP0 = 4;
constexpr int SCAN_THREADS = 256;
constexpr int UPD_WARP0 = 12;
constexpr int UPD_THREADS = 256;
constexpr int NUM_THREADS = UPD_WARP0 * 32 + UPD_THREADS;  // 640
constexpr int MAX_STAGES_RING = 6;
constexpr int MAX_NQ = 64;
constexpr uint32_t B_STAGE_BYTES = CHUNK_N * KSLICE * 2;  // 32 KiB
constexpr uint32_t BAR_SCAN = 1;  // named barrier of the 256 scan threads
constexpr uint32_t BAR_UPD = 2;   // named barrier of the update threads
constexpr int ITEM_CAP = 256;     // re-rank work items per tile-stage (frames beyond take the exact scan)

struct EncParams {
    const float* x;
    long long N;
    RowAddrT ad;
    int d, nq, K, Kpad, q_begin;
    const float* cb;       // [*, K, d] fp32 master
    const float* cb_norm;  // [*, Kpad] scaled norms
    const float* cb_meta;  // [*, 8]
    float* xq;
    long long* idx;
    double* commit_sq;
    float* stats_sum;
    float* stats_cnt;
    float* r_scratch;  // per-CTA [128, d] fp32 when the residual tile does not fit in shared memory
    int num_tiles, nstage, nslots, r_pitch;
    int cluster;  // CTAs per cluster sharing the codebook stream (TMA multicast)
    int tile_rows;  // frames per tile (<= TILE_M): small calls spread over more SMs with partly filled tiles
    int a_rows;     // rows of every operand slice kept in shared memory (32 | 64 | 128, >= tile_rows)
    uint32_t off_B, off_misc;  // A tiles (one per slot) at offset 0
    float* dbg_scores;                // [128, Kpad] (bring-up hook) or null
    float* dbg_rowscale;              // [128] or null
    unsigned long long* prof;         // [16] cycle / event counters (RVQ_PROFILE=1) or null
};

struct __align__(16) Misc {
    uint64_t full[MAX_STAGES_RING], empty[MAX_STAGES_RING], tmem_full[2], tmem_empty[2], a_ready[2], scan_done[2];
    uint64_t norm_full[2];
    alignas(16) float norms[2][CHUNK_N];  // scaled ||c||^2 of the chunk in each accumulator buffer (bulk-copied)
    alignas(16) float xc[2][CHUNK_N];     // allowances XC_k of the chunk's codes above the norm cap (k0_bound)
    float grp_x[2][2][TILE_M];            // [job parity][scan group][frame]: allowance of the group's best code
    uint32_t tmem_base;
    int dirty_count[2];
    float row_na[2][TILE_M], row_delta[2][TILE_M], row_amax[2][TILE_M], row_rs[2][TILE_M];  // per tile slot
    float grp_best[2][2][TILE_M];            // [job parity][scan group][frame]: best score the group saw
    uint32_t g_rows[2][2][TILE_M];           // [slot][group][frame]: loads that may hold a candidate
    uint16_t g_cols[2][2][TILE_M];           // [slot][group][frame]: columns that may hold a candidate
    uint16_t dirty_cols[2][TILE_M];          // per exact-scan frame: columns to scan
    int win[2][TILE_M];                      // [slot][frame]: selected code, -1 = exact scan pending
    int dirty_rows[2][TILE_M];
    // exact re-rank work items: one PAIR of candidates of one frame each, so that a frame with four candidates is
    // scored by two 8-lane groups at once instead of in two dependent rounds (one buffer: jobs re-rank one at a time)
    int score_count[2];                 // [0] items, [1] frames with several candidates (counter only)
    int n_items;                        // items of the current job (published between two barriers)
    uint16_t item[ITEM_CAP];            // frame | pair << 8
    float item_s[ITEM_CAP];             // best exact score of the pair
    int item_k[ITEM_CAP];               // its code
    uint16_t row_item0[TILE_M];         // first item of the frame
    uint8_t row_nitem[TILE_M];          // number of items of the frame (0: not a re-rank frame)
    float dirty_s[8];
    int dirty_k[8];
    double commit_acc[MAX_NQ];
};

// ---------------------------------------------------------------------------------------------------------
// residual tile access: element (row, col) -> float*
struct RTile {
    float* base;
    int pitch;
    __device__ __forceinline__ float* at(int row, int col) const { return base + (size_t)row * pitch + col; }
};

// per-row constants of the next stage from the new residual's norm and the operand scale
__device__ __forceinline__ void write_row_consts(const EncParams& p, Misc* misc, int sl, int row, int d, float sq,
                                                 bool force_exact, int a, int b, float sb, float cnmax) {
    float na, delta, rs;
    row_consts(d, sq, force_exact, a, b, sb, cnmax, na, delta, rs);
    misc->row_na[sl][row] = na;
    misc->row_delta[sl][row] = delta;
    misc->row_rs[sl][row] = rs;
    if (p.dbg_rowscale) p.dbg_rowscale[row] = exp2i(a);
}

__device__ __forceinline__ void store_a4(uint8_t* smem_a, int row, int c, float4 v, float sa, uint32_t asb) {
    const __half2 h01 = __floats2half2_rn(v.x * sa, v.y * sa);
    const __half2 h23 = __floats2half2_rn(v.z * sa, v.w * sa);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h01);
    pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(smem_a + a_tile_offset(row, c, asb)) = pk;
}

// NP float4 pieces per lane of one frame: r <- r - c (+ statistics, norms, next operand)
template <int NP>
__device__ __forceinline__ void apply_seg(uint8_t* smem_a, const RTile& rt, int row, const float* __restrict__ cw,
                                          float* __restrict__ ssum, int c0, bool write_a, float sa, float& sq,
                                          uint32_t asb) {
    float4 rv[NP], cv[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        rv[i] = *reinterpret_cast<float4*>(rt.at(row, c0 + i * 32));
        cv[i] = ldg_nc_v4(cw + c0 + i * 32);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int c = c0 + i * 32;
        if (ssum) red_add_v4(ssum + c, rv[i]);
        float4 nr;
        nr.x = rv[i].x - cv[i].x;
        nr.y = rv[i].y - cv[i].y;
        nr.z = rv[i].z - cv[i].z;
        nr.w = rv[i].w - cv[i].w;
        *reinterpret_cast<float4*>(rt.at(row, c)) = nr;
        sq = fmaf(nr.x, nr.x, sq);
        sq = fmaf(nr.y, nr.y, sq);
        sq = fmaf(nr.z, nr.z, sq);
        sq = fmaf(nr.w, nr.w, sq);
        if (write_a) store_a4(smem_a, row, c, nr, sa, asb);
    }
}

// The two halves of apply_seg<8> for the software-pipelined apply (d % 256 == 0): the 16 loads of a 256-feature segment,
// and everything that consumes them.  Inactive frames (exact scan pending) load harmlessly and store nothing; every
// element of the register arrays is assigned (conditionally initialised arrays are demoted to local memory).
__device__ __forceinline__ void seg8_load(const RTile& rt, int row, const float* __restrict__ cw, int c0,
                                          float4 (&rv)[8], float4 (&cv)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        rv[i] = *reinterpret_cast<float4*>(rt.at(row, c0 + i * 32));
        cv[i] = ldg_nc_v4(cw + c0 + i * 32);
    }
}
__device__ __forceinline__ void seg8_consume(uint8_t* smem_a, const RTile& rt, int row, float* __restrict__ ssum, int c0,
                                             bool active, bool write_a, float sa, float& sq, uint32_t asb,
                                             const float4 (&rv)[8], const float4 (&cv)[8]) {
    // packed fp32 pairs (Blackwell FFMA2 / FMUL2): half the floating-point instructions of the pass.  r - c as
    // fma(c, -1, r) is exact like the subtraction; the squared norm runs in two interleaved chains (even / odd
    // elements) that are added at the end - it feeds the commit loss (rtol 1e-5) and the error bound (which carries a
    // 2e-5 slack for the fp32 summation), never an index.
    const float2 m1 = make_float2(-1.f, -1.f), sa2 = make_float2(sa, sa);
    float2 sq2 = make_float2(sq, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i * 32;
        if (ssum) red_add_v4(ssum + c, rv[i]);
        const float2 lo = __ffma2_rn(make_float2(cv[i].x, cv[i].y), m1, make_float2(rv[i].x, rv[i].y));
        const float2 hi = __ffma2_rn(make_float2(cv[i].z, cv[i].w), m1, make_float2(rv[i].z, rv[i].w));
        if (active) *reinterpret_cast<float4*>(rt.at(row, c)) = make_float4(lo.x, lo.y, hi.x, hi.y);
        sq2 = __ffma2_rn(lo, lo, sq2);
        sq2 = __ffma2_rn(hi, hi, sq2);
        if (active && write_a) {
            const float2 slo = __fmul2_rn(lo, sa2), shi = __fmul2_rn(hi, sa2);
            const __half2 h01 = __floats2half2_rn(slo.x, slo.y);
            const __half2 h23 = __floats2half2_rn(shi.x, shi.y);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h01);
            pk.y = *reinterpret_cast<const uint32_t*>(&h23);
            *reinterpret_cast<uint2*>(smem_a + a_tile_offset(row, c, asb)) = pk;
        }
    }
    sq = sq2.x + sq2.y;
}

// An 8-lane group applies one stage to one frame in ONE pass over memory:
//   r <- r - c_win (fp32, residual tile in the L2-resident scratch), EMA statistics of the stage input,
//   squared norm of the new residual, and the fp16 operand row + row constants of the next stage.
// The operand scale of the next stage is chosen from the bound max|r'| <= ||r||_2 + max|c| (known before
// the pass), so the converted row is written in the same pass; the error bound uses the exact new norm.
// All 32 lanes of the warp must call this together (8-lane shuffles with a full mask); `active` gates effects.
//   next_q_abs < 0 : last stage (no operand for a next stage)
// per-job constants of apply_row (read from the stage metadata ONCE per job, not per row: the metadata loads were a
// second global round trip in front of every row's residual / code loads)
struct StageC {
    float sb, cnmax, cmax_q;  // 2^b and max ||c||_2 of the NEXT stage, max |c| of this stage
};
__device__ __forceinline__ StageC load_stage_consts(const EncParams& p, int q_abs, int next_q_abs) {
    StageC sc{1.f, 0.f, 0.f};
    if (next_q_abs >= 0) {
        const float* mq = p.cb_meta + (size_t)next_q_abs * META_STRIDE;
        sc.sb = mq[0];
        sc.cnmax = mq[1];
        sc.cmax_q = p.cb_meta[(size_t)q_abs * META_STRIDE + 2];
    }
    return sc;
}

__device__ __forceinline__ void apply_row(const EncParams& p, Misc* misc, uint8_t* smem_a, const RTile& rt, int sl,
                                          int row, bool active, bool row_valid, int kwin, int q_abs, int next_q_abs,
                                          int sub, float* sq_out, const StageC& sc) {
#ifdef RVQ_TC_APPLY_PROF  // build-time experiment switch (RVQ_NVCC_DEFS): splits the apply time of thread 0 into
    const long long tp0 = clock64();  // prologue / memory pass / reduction + row constants (counters 12-15)
#endif
    const int d = p.d;
    float sq = 0.f;
    float sa = 0.f;
    const float sb = sc.sb, cnmax = sc.cnmax;
    int a = 0, b = 0;
    bool force_exact = false;
    const bool write_a = next_q_abs >= 0;
    if (write_a) {
        b = ilog2f_floor(sb);
        const float bound = misc->row_amax[sl][row] + sc.cmax_q;
        a = pick_row_exp(bound, b, force_exact);
        sa = exp2i(a);
    }
    const uint32_t asb = (uint32_t)p.a_rows * 128u;
#ifdef RVQ_TC_APPLY_PROF
    const long long tp1 = clock64();
#endif
    if (active) {
        const float* cw = p.cb + ((size_t)q_abs * p.K + kwin) * d;
        float* ssum = (p.stats_sum && row_valid) ? p.stats_sum + ((size_t)q_abs * p.K + kwin) * d : nullptr;
        int c0 = sub * 4;
#pragma unroll 1
        // 256 features per step when possible: one L2 round trip covers eight pieces of the frame and of the code
        for (; c0 + 256 <= d + sub * 4; c0 += 256) apply_seg<8>(smem_a, rt, row, cw, ssum, c0, write_a, sa, sq, asb);
        if (d & 128) {
            apply_seg<4>(smem_a, rt, row, cw, ssum, c0, write_a, sa, sq, asb);
            c0 += 128;
        }
        if (d & 64) apply_seg<2>(smem_a, rt, row, cw, ssum, c0, write_a, sa, sq, asb);
    }
#ifdef RVQ_TC_APPLY_PROF
    const long long tp2 = clock64();
#endif
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (sq_out) *sq_out = sq;
    if (active && sub == 0) {
        // bound on max|r'| for the next stage's operand scale: ||r'||_2 (as the TMEM kernel does).  The true maximum
        // cost four FMNMX per 16 bytes on the ALU pipe the scan warps saturate; the scale only has to keep fp16 from
        // overflowing, its relative precision does not depend on it
        misc->row_amax[sl][row] = sqrtf(sq) * 1.00002f;
        if (write_a) {
            if (!isfinite(sq)) force_exact = true;
            write_row_consts(p, misc, sl, row, d, sq, force_exact, a, b, sb, cnmax);
        }
    }
#ifdef RVQ_TC_APPLY_PROF
    if (p.prof && threadIdx.x == UPD_WARP0 * 32) {
        const long long tp3 = clock64();
        atomicAdd(p.prof + 12, (unsigned long long)(tp1 - tp0));
        atomicAdd(p.prof + 13, (unsigned long long)(tp2 - tp1));
        atomicAdd(p.prof + 14, (unsigned long long)(tp3 - tp2));
        atomicAdd(p.prof + 15, 1ull);
    }
#endif
}

// Stage-0 initialisation of one frame by an 8-lane group: x -> residual scratch, exact max -> operand scale,
// fp16 operand row and row constants of the first stage (two passes: the scale needs the row maximum).
__device__ __forceinline__ void init_row(const EncParams& p, Misc* misc, uint8_t* smem_a, const RTile& rt, int sl,
                                         int row, const float* __restrict__ xr, bool row_valid, int sub,
                                         bool stream_x) {
    const int d = p.d;
    float sq = 0.f, amax = 0.f;
    const int np = d / 32;  // 16-byte pieces per lane
#pragma unroll 1
    for (int i0 = 0; i0 < np; i0 += 8) {
        // eight pieces per lane in flight (one memory round trip per 256 features; d = 512 took sixteen dependent ones)
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = sub * 4 + (i0 + i) * 32;
            // x streams through L2 once (evict-first) so that it does not push the residual scratch out
            v[i] = (row_valid && i0 + i < np) ? (stream_x ? __ldcs(reinterpret_cast<const float4*>(xr + c))
                                                           : *reinterpret_cast<const float4*>(xr + c))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i0 + i < np) {
                *reinterpret_cast<float4*>(rt.at(row, sub * 4 + (i0 + i) * 32)) = v[i];
                sq = fmaf(v[i].x, v[i].x, sq);
                sq = fmaf(v[i].y, v[i].y, sq);
                sq = fmaf(v[i].z, v[i].z, sq);
                sq = fmaf(v[i].w, v[i].w, sq);
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
            }
        }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    const float* mq = p.cb_meta + (size_t)p.q_begin * META_STRIDE;
    const float sb = mq[0], cnmax = mq[1];
    const int b = ilog2f_floor(sb);
    bool force_exact = !isfinite(sq);
    const int a = pick_row_exp(amax, b, force_exact);
    const float sa = exp2i(a);
#pragma unroll 1
    for (int i0 = 0; i0 < np; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)  // (written by this lane above; every element assigned: no local-memory demotion)
            v[i] = *reinterpret_cast<const float4*>(rt.at(row, sub * 4 + min(i0 + i, np - 1) * 32));
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i0 + i < np) store_a4(smem_a, row, sub * 4 + (i0 + i) * 32, v[i], sa, (uint32_t)p.a_rows * 128u);
    }
    if (sub == 0) {
        misc->row_amax[sl][row] = amax;
        write_row_consts(p, misc, sl, row, d, sq, force_exact, a, b, sb, cnmax);
    }
}

template <bool kDebug>
__global__ void __launch_bounds__(NUM_THREADS, 1)
rvq_encode_tc_kernel(const __grid_constant__ CUtensorMap tmap_b, const EncParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_b = smem + p.off_B;
    Misc* misc = reinterpret_cast<Misc*>(smem + p.off_misc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.d, nq = p.nq;
    const int n_ks = d / KSLICE;
    const int n_chunks = p.Kpad / CHUNK_N;
    const int nstage = p.nstage;
    const uint32_t a_slice_bytes = (uint32_t)p.a_rows * 128u;
    const uint32_t a_tile_bytes = (uint32_t)n_ks * a_slice_bytes;
    // every CTA of a cluster walks the same job sequence (tiles past the end are empty: all frames invalid)
    const int n_local = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int CL = p.cluster;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    const int nslots = p.nslots;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_init(&misc->full[i], 1);
            mbar_init(&misc->empty[i], (uint32_t)CL);  // one tcgen05.commit arrive per CTA of the cluster
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&misc->tmem_full[i], 1);
            mbar_init(&misc->tmem_empty[i], 4);  // one arrive per scan warp of the group
            mbar_init(&misc->norm_full[i], 1);
            mbar_init(&misc->a_ready[i], UPD_THREADS);
            mbar_init(&misc->scan_done[i], SCAN_THREADS);
            misc->dirty_count[i] = 0;
            misc->score_count[i] = 0;
        }
        for (int i = 0; i < MAX_NQ; ++i) misc->commit_acc[i] = 0.0;
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_b);
    if (warp == 2) tmem_alloc<512>(&misc->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peers' barriers are initialised before anything is multicast to them
    tc_fence_after_sync();
    const uint32_t tmem_base = misc->tmem_base;

    // Register budget: 640 threads x 96 = 61440 registers at launch.  setmaxnreg.inc can only take what
    // setmaxnreg.dec released inside this CTA, so the totals after rebalancing must not exceed the launch
    // allocation: 40*128 (control) + 88*256 (scan) + 128*256 (update) = 60416 <= 61440.
    // setmaxnreg is warpgroup-aligned: the four warps of a warpgroup must execute the SAME instruction, so the
    // control warpgroup releases its registers once, before its warps split into roles.
    if (warp < SCAN_WARP0) reg_dealloc<40>();
    if (warp == 0) {
        // =========================================================== TMA producer (codebook slices)
        if (elect_one()) {
            uint32_t st = 0, ph = 0;
            const uint32_t part_bytes = B_STAGE_BYTES / (uint32_t)CL;
            const int part_rows = CHUNK_N / CL;
            for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
                const int row0 = (p.q_begin + job.q) * p.Kpad + (int)crank * part_rows;
                for (int c = 0; c < n_chunks; ++c) {
                    for (int ks = 0; ks < n_ks; ++ks) {
                        mbar_wait(&misc->empty[st], ph ^ 1);  // every CTA of the cluster has consumed the slot
                        mbar_arrive_expect_tx(&misc->full[st], B_STAGE_BYTES);
                        uint8_t* dst = smem_b + (size_t)st * B_STAGE_BYTES + crank * part_bytes;
                        if (CL > 1)  // my 1/CL of the slice goes to every CTA of the cluster
                            tma_load_2d_mc(dst, &tmap_b, &misc->full[st], ks * KSLICE, row0 + c * CHUNK_N, cmask);
                        else
                            tma_load_2d(dst, &tmap_b, &misc->full[st], ks * KSLICE, row0 + c * CHUNK_N);
                        if (++st == (uint32_t)nstage) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =========================================================== MMA issuer
        // ONE elected thread runs the whole loop: a lone warp pays ~8 cycles of latency per instruction, so the
        // issue loop stays at a handful of instructions per MMA (no divisions, no per-iteration warp syncs,
        // descriptors advanced by additions; elect.sync keeps ptxas from wrapping every MMA in a waterfall loop).
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(0 /*fp16*/, TILE_M, CHUNK_N);
            const uint64_t bdesc0 = umma_desc_sw128(smem_u32(smem_b));
            const int nq_prep = (int)p.cb_meta[4];  // stages prepared: locates the arrays behind the norms
            uint32_t g = 0, aphase = 0, st = 0, ph = 0;
            for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
                const int sl = job.slot % nslots;
                mbar_wait(&misc->a_ready[sl], (aphase >> sl) & 1);
                aphase ^= 1u << sl;
                tc_fence_after_sync();
                const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem + (size_t)sl * a_tile_bytes));
                const float* nsrc = p.cb_norm + (size_t)(p.q_begin + job.q) * p.Kpad;
                const float* xsrc = NormLayout(p.cb_norm, nq_prep, p.Kpad).xc + (size_t)(p.q_begin + job.q) * p.Kpad;
                for (int c = 0; c < n_chunks; ++c, ++g) {
                    const uint32_t buf = g & 1, use = g >> 1;
                    mbar_wait(&misc->tmem_empty[buf], (use & 1) ^ 1);
                    tc_fence_after_sync();
                    // the scan group has released this buffer: its norm slice can be replaced as well
                    mbar_arrive_expect_tx(&misc->norm_full[buf], 2 * CHUNK_N * 4);
                    bulk_load_1d(misc->norms[buf], nsrc + c * CHUNK_N, CHUNK_N * 4, &misc->norm_full[buf]);
                    bulk_load_1d(misc->xc[buf], xsrc + c * CHUNK_N, CHUNK_N * 4, &misc->norm_full[buf]);
                    const uint32_t tmem_d = tmem_base + buf * CHUNK_N;
                    uint64_t adesc = adesc0;
                    for (int ks = 0; ks < n_ks; ++ks) {
                        mbar_wait(&misc->full[st], ph);
                        tc_fence_after_sync();
                        const uint64_t bdesc = bdesc0 + (uint64_t)(st * (B_STAGE_BYTES >> 4));
                        // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                        umma_f16_ss(tmem_d, adesc, bdesc, idesc, ks != 0);
                        umma_f16_ss(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
                        umma_f16_ss(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
                        umma_f16_ss(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
                        // frees the ring slot (in every CTA of the cluster) when these MMAs retire
                        if (CL > 1)
                            umma_commit_mc(&misc->empty[st], cmask);
                        else
                            umma_commit(&misc->empty[st]);
                        adesc += (uint64_t)(a_slice_bytes >> 4);
                        if (++st == (uint32_t)nstage) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                    umma_commit(&misc->tmem_full[buf]);
                }
            }
        }
        __syncwarp();
    } else if (warp < SCAN_WARP0) {
    } else if (warp < UPD_WARP0) {
        reg_dealloc<88>();
        // =========================================================== scan groups (argmin epilogue)
        const int e = threadIdx.x - SCAN_WARP0 * 32;  // 0..255
        const int grp = e >> 7;                       // scan group = accumulator buffer
        const int my_row = (warp & 3) * 32 + lane;    // TMEM lane owned by this thread
        uint32_t g = 0, aphase = 0, jpar = 0;
        long long t_scan = 0, t_wait = 0, t_full = 0;
        const NormLayout nl(p.cb_norm, (int)p.cb_meta[4], p.Kpad);
        for (JobIter job(n_local, nq, nslots); job.valid(); job.next(), jpar ^= 1u) {
            const int sl = job.slot % nslots;
            const int q_abs = p.q_begin + job.q;
            const int* xflag = nl.xflag + (size_t)q_abs * n_chunks;
            const float2 mq_x = make_float2(p.cb_meta[(size_t)q_abs * META_STRIDE + 7], p.cb_meta[(size_t)q_abs * META_STRIDE + 5]);
            long long t0 = clock64();
            mbar_wait(&misc->a_ready[sl], (aphase >> sl) & 1);  // row constants of this job are visible
            aphase ^= 1u << sl;
            const float na = misc->row_na[sl][my_row];
            const float delta = misc->row_delta[sl][my_row];
            const float rs = misc->row_rs[sl][my_row];
            float Cm[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) Cm[j] = BIG;
            float m1 = BIG, m2 = BIG, m3 = BIG, m4 = BIG;
            long long t1 = clock64();
            t_wait += t1 - t0;
            for (int c = 0; c < n_chunks; ++c, ++g) {
                if ((int)(g & 1) != grp) continue;
                const long long tw0 = clock64();
                mbar_wait(&misc->tmem_full[grp], (g >> 1) & 1);
                mbar_wait(&misc->norm_full[grp], (g >> 1) & 1);
                tc_fence_after_sync();
                t_full += clock64() - tw0;
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + grp * CHUNK_N;
                const float* nptr = misc->norms[grp];
                const float* xptr = misc->xc[grp];
                const bool has_large = __ldg(xflag + c) != 0;  // warp-uniform: the chunk holds a code above the norm cap
                float* dbg = nullptr;
                if (kDebug && p.dbg_scores && job.i == 0 && job.q == 0)
                    dbg = p.dbg_scores + (size_t)my_row * p.Kpad + c * CHUNK_N;
                // 16 columns per TMEM load, double buffered; the loop body (32 columns) stays small enough for
                // the instruction cache
                uint32_t va[16], vb[16];
                tmem_ld_32x16(taddr, va);
                uint32_t it = (uint32_t)c * (CHUNK_N / 16);
                if (!has_large) {
#pragma unroll 1
                    for (int cb = 0; cb < CHUNK_N; cb += 32, it += 2) {
                        tmem_ld_wait();
                        tmem_ld_32x16(taddr + cb + 16, vb);
                        scan16_2d<false>(va, nptr + cb, nullptr, na, 0.f, it, Cm, m1, m2, m3, m4, kDebug && dbg ? dbg + cb : nullptr);
                        tmem_ld_wait();
                        if (cb + 32 < CHUNK_N) tmem_ld_32x16(taddr + cb + 32, va);
                        scan16_2d<false>(vb, nptr + cb + 16, nullptr, na, 0.f, it + 1, Cm, m1, m2, m3, m4,
                                         kDebug && dbg ? dbg + cb + 16 : nullptr);
                    }
                } else {
#pragma unroll 1
                    for (int cb = 0; cb < CHUNK_N; cb += 32, it += 2) {
                        tmem_ld_wait();
                        tmem_ld_32x16(taddr + cb + 16, vb);
                        scan16_2d<true>(va, nptr + cb, xptr + cb, na, -rs, it, Cm, m1, m2, m3, m4, kDebug && dbg ? dbg + cb : nullptr);
                        tmem_ld_wait();
                        if (cb + 32 < CHUNK_N) tmem_ld_32x16(taddr + cb + 32, va);
                        scan16_2d<true>(vb, nptr + cb + 16, xptr + cb + 16, na, -rs, it + 1, Cm, m1, m2, m3, m4,
                                        kDebug && dbg ? dbg + cb + 16 : nullptr);
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&misc->tmem_empty[grp]);
            }
            // ---------------- stage end: exchange the groups' best scores, list this group's candidates
            float vb_ = fminf(fminf(Cm[0], Cm[1]), Cm[2]);
#pragma unroll
            for (int j = 3; j < 15; j += 2) vb_ = fminf(fminf(vb_, Cm[j]), Cm[j + 1]);
            vb_ = fminf(vb_, Cm[15]);
            misc->grp_best[jpar][grp][my_row] = vb_;
            // allowance of the code behind the group's best (optimistic) score: the threshold needs twice the
            // allowance of the OVERALL best code (zero unless that code is above the stage's norm cap)
            {
                int jmin = 15;
#pragma unroll
                for (int j = 14; j >= 0; --j) jmin = (Cm[j] == vb_) ? j : jmin;
                misc->grp_x[jpar][grp][my_row] =
                    best_allowance(vb_, jmin, m1, m2, m3, m4, fmaf(rs, mq_x.x, na * mq_x.y),
                                   nl.xb + (size_t)q_abs * p.Kpad, p.Kpad - 1);
            }
            named_bar_sync(BAR_SCAN, SCAN_THREADS);
            {
                const float ob = misc->grp_best[jpar][grp ^ 1][my_row];
                const float best = fminf(vb_, ob);
                const float xm = misc->grp_x[jpar][grp][my_row], xo = misc->grp_x[jpar][grp ^ 1][my_row];
                const float tol = fabsf(best) * 6.2e-5f;
                const float xbest = vb_ + tol < ob ? xm : (ob + tol < vb_ ? xo : fmaxf(xm, xo));
                // Certificate: a code can be the exact argmin only if its optimistic score is <= T.
                const float dl = delta + 2.f * xbest;
                const float T = best + dl;
                // load minima carry `it` in their low 9 mantissa bits: |packed - r| <= 2^-14 |r|, and every load
                // minimum r of interest lies in [best, T], so |r| <= |best| + dl
                const float T2 = T + (fabsf(best) + 2.f * dl) * 1.220703125e-4f;
                // NaN / overflow / forced exact (no usable filter result), or more than three loads in reach
                const bool nofilter = !(best < BIG) || !(T2 < BIG);
                const bool over = nofilter || (m4 <= T2);
                const uint32_t nr = (uint32_t)(m1 <= T2) + (uint32_t)(m2 <= T2) + (uint32_t)(m3 <= T2);
                uint32_t cols = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) cols |= (Cm[j] <= T) ? (1u << j) : 0u;
                if (nofilter) cols = 0xFFFFu;
                const uint32_t rows = (__float_as_uint(m1) & IT_MASK) | ((__float_as_uint(m2) & IT_MASK) << 9) |
                                      ((__float_as_uint(m3) & IT_MASK) << 18) | (nr << 27) | (over ? G_OVER : 0u);
                misc->g_rows[sl][grp][my_row] = rows;
                misc->g_cols[sl][grp][my_row] = (uint16_t)cols;
            }
            mbar_arrive(&misc->scan_done[sl]);
            t_scan += clock64() - t1;
        }
        if (p.prof && e == 0) {
            atomicAdd(p.prof + 0, (unsigned long long)t_scan);
            atomicAdd(p.prof + 1, (unsigned long long)t_wait);
            atomicAdd(p.prof + 11, (unsigned long long)t_full);
        }
    } else {
        reg_alloc<128>();
        // =========================================================== update warps
        const int u = threadIdx.x - UPD_WARP0 * 32;  // 0..UPD_THREADS-1
        const int sub = u & 7, slot16 = u >> 3;      // 8-lane group per frame
        const int uwarp = u >> 5;
        constexpr int ROWS_PER_PASS = UPD_THREADS / 8;
        constexpr int UPD_WARPS = UPD_THREADS / 32;
        const bool row_major = (p.ad.sd == 1);
        long long t_upd = 0, t_dirty = 0, t_wait = 0, t_score = 0, t_apply = 0;
        unsigned long long n_dirty_tot = 0, n_two_tot = 0, n_jobs = 0, n_score_pass = 0;
        double commit_local = 0.0;

        // rows the update passes walk (whole passes of 32 frames; the rest of the 128 lanes stays zero)
        const int tile_rows = p.tile_rows;
        const int rows_eff = (tile_rows + ROWS_PER_PASS - 1) & ~(ROWS_PER_PASS - 1);
        auto rtile = [&](int sl) {
            RTile rt;
            rt.base = p.r_scratch + ((size_t)blockIdx.x * 2 + sl) * TILE_M * p.r_pitch;
            rt.pitch = p.r_pitch;
            return rt;
        };
        // load tile `tile` into slot `sl`: residual <- x, fp16 operand + row constants of stage 0
        auto load_tile = [&](int sl, int tile) {
            const RTile rt = rtile(sl);
            uint8_t* a_tile = smem + (size_t)sl * a_tile_bytes;
            const long long n0 = (long long)tile * tile_rows;
            if (!row_major) {
                // The reference's (B, d, L) storage: frames are the fastest axis.  A warp takes 32 frames x 32 features:
                // lane = frame (coalesced 128-byte reads per feature), the 32 loads of a lane are independent (one
                // memory round trip per unit, as many round trips per tile as the row-major load needs), and they leave
                // as eight 16-byte stores into the lane's residual row.
                const int nrb = rows_eff / 32, units = nrb * (d / 32);
#pragma unroll 1
                for (int unit = uwarp; unit < units; unit += UPD_WARPS) {
                    const int row = (unit % nrb) * 32 + (u & 31), c0 = (unit / nrb) * 32;
                    const long long n = n0 + row;
                    const bool ok = row < tile_rows && n < p.N;
                    const float* xr = p.x + (ok ? p.ad.row(n) : 0) + (long long)c0 * p.ad.sd;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ok ? __ldcs(xr + (long long)j * p.ad.sd) : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(rt.at(row, c0 + j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                named_bar_sync(BAR_UPD, UPD_THREADS);
            }
#pragma unroll 1
            for (int row = slot16; row < p.a_rows; row += ROWS_PER_PASS) {  // (the operand tile holds a_rows rows)
                const long long n = n0 + row;
                const bool ok = row < tile_rows && n < p.N;
                const float* xr = row_major ? p.x + (ok ? p.ad.row(n) : 0) : rt.at(row, 0);
                init_row(p, misc, a_tile, rt, sl, row, xr, row_major ? ok : row < rows_eff, sub, row_major);
            }
            fence_proxy_async_smem();
            mbar_arrive(&misc->a_ready[sl]);
        };

        if (n_local > 0) load_tile(0, blockIdx.x);
        if (n_local > 1 && nslots > 1) load_tile(1, blockIdx.x + gridDim.x);
        uint32_t sphase = 0;
        for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
            const int sl = job.slot % nslots;
            const int q = job.q, q_abs = p.q_begin + q;
            const int tile = blockIdx.x + job.i * gridDim.x;
            const long long n0 = (long long)tile * tile_rows;
            auto frame_ok = [&](int row) { return row < tile_rows && n0 + row < p.N; };
            const RTile rt = rtile(sl);
            uint8_t* a_tile = smem + (size_t)sl * a_tile_bytes;
            long long t0 = clock64();
            mbar_wait(&misc->scan_done[sl], (sphase >> sl) & 1);
            sphase ^= 1u << sl;
            long long t1 = clock64();
            t_wait += t1 - t0;
            const int next_q_abs = (q + 1 < nq) ? q_abs + 1 : -1;
            const float* cbq = p.cb + (size_t)q_abs * p.K * d;
            const StageC sc = load_stage_consts(p, q_abs, next_q_abs);
            // ---------------- classify the frames: certified (one candidate), re-rank list, exact-scan list
            const long long ts0 = clock64();
            const int Kv_q = (int)p.cb_meta[(size_t)q_abs * META_STRIDE + 3];
            if (u < rows_eff) {
                const uint32_t r0 = misc->g_rows[sl][0][u], r1 = misc->g_rows[sl][1][u];
                const uint32_t c0 = misc->g_cols[sl][0][u], c1 = misc->g_cols[sl][1][u];
                const int n0 = (int)((r0 >> 27) & 3u) * __popc(c0), n1 = (int)((r1 >> 27) & 3u) * __popc(c1);
                int w = -1, nitem = 0;
                bool dirty = ((r0 | r1) & G_OVER) || n0 + n1 == 0;
                if (!dirty && n0 + n1 == 1) {
                    w = n0 ? (int)((r0 & IT_MASK) * 16u) + __ffs(c0) - 1 : (int)((r1 & IT_MASK) * 16u) + __ffs(c1) - 1;
                    if (w >= Kv_q) w = 0;  // cannot happen (padding codes score 2^100); keeps the gather in bounds
                } else if (!dirty) {
                    const int npairs = (n0 + n1 + 1) >> 1;
                    const int pos = npairs <= 16 ? atomicAdd(&misc->score_count[0], npairs) : ITEM_CAP;
                    if (pos + npairs <= ITEM_CAP) {
                        for (int t = 0; t < npairs; ++t) misc->item[pos + t] = (uint16_t)(u | (t << 8));
                        misc->row_item0[u] = (uint16_t)pos;
                        nitem = npairs;
                        w = 0;  // replaced by the re-rank below
                        atomicAdd(&misc->score_count[1], 1);
                    } else {
                        dirty = true;  // more than 32 candidates (or the item list is full): exact scan of its columns
                    }
                }
                if (dirty) {
                    const int pos = atomicAdd(&misc->dirty_count[sl], 1);
                    misc->dirty_rows[sl][pos] = u;
                    misc->dirty_cols[sl][pos] = (uint16_t)((c0 | c1) ? (c0 | c1) : 0xFFFFu);
                }
                misc->row_nitem[u] = (uint8_t)nitem;
                misc->win[sl][u] = w;
            }
            named_bar_sync(BAR_UPD, UPD_THREADS);
            // the counters are taken and cleared between two barriers: the next job's classification (which no barrier
            // separates from the end of this job) finds them zero, and nobody reads them while they change
            if (u == 0) {
                misc->n_items = min(misc->score_count[0], ITEM_CAP);
                n_two_tot += misc->score_count[1];
                misc->score_count[0] = misc->score_count[1] = 0;
            }
            named_bar_sync(BAR_UPD, UPD_THREADS);
            // ---------------- exact re-rank: one pair of candidates per 8-lane group and round
            const int n_score = misc->n_items;
            for (int base = 0; base < n_score; base += ROWS_PER_PASS) {
                const int i = base + slot16;
                const bool sc = i < n_score;
                const uint32_t it = misc->item[sc ? i : 0];
                const int row = (int)(it & 0x7fu), pair = (int)(it >> 8) & 15;  // (masked: entries past a full list are stale)
                const CandSet cs(misc->g_rows[sl][0][row], misc->g_cols[sl][0][row], misc->g_rows[sl][1][row],
                                 misc->g_cols[sl][1][row]);
                const int c1 = cs.code(2 * pair, Kv_q - 1), c2 = cs.code(2 * pair + 1, Kv_q - 1);
                const float* cc[2] = {cbq + (size_t)c1 * d, cbq + (size_t)c2 * d};
                float sv[2];
                exact_score8_n<2>(rt.at(row, 0), cc, d, sub, sv);
                float bs = sv[0];
                int kwin = c1;
                if (better(sv[1], c2, bs, kwin)) {
                    bs = sv[1];
                    kwin = c2;
                }
                if (sc && sub == 0) {
                    misc->item_s[i] = bs;
                    misc->item_k[i] = kwin;
                }
                ++n_score_pass;
            }
            if (n_score > 0) {
                named_bar_sync(BAR_UPD, UPD_THREADS);
                if (u < rows_eff && misc->row_nitem[u] > 0) {
                    const int i0 = misc->row_item0[u], ni = misc->row_nitem[u];
                    float bs = misc->item_s[i0];
                    int kwin = misc->item_k[i0];
                    for (int t = 1; t < ni; ++t)
                        if (better(misc->item_s[i0 + t], misc->item_k[i0 + t], bs, kwin)) {
                            bs = misc->item_s[i0 + t];
                            kwin = misc->item_k[i0 + t];
                        }
                    misc->win[sl][u] = kwin;
                }
                named_bar_sync(BAR_UPD, UPD_THREADS);  // winners visible to the applying groups
            }
            const long long ts1 = clock64();
            // ---------------- gather, residual update, statistics, next operand
            auto post_row = [&](int row, bool active, int kwin, float sq) {
                const long long n = n0 + row;
                if (active && sub == 0 && frame_ok(row)) {
                    __stcs(p.idx + n * nq + q, (long long)kwin);
                    commit_local += (double)sq;
                    if (p.stats_cnt) atomicAdd(p.stats_cnt + (size_t)q_abs * p.K + kwin, 1.f);
                }
            };
            if ((d & 255) == 0) {
                // Software-pipelined passes: the loads of the NEXT frame's first segment are issued before the reduction,
                // row constants and index / statistics writes of the current frame, so that one memory round trip per
                // pass overlaps ~1.3 k cycles of dependent work instead of following it (profiles/r2z_apply_split.log:
                // loads 1.2 k, the rest 2.5 k cycles of a pass).  Same arithmetic in the same order as apply_row.
                const bool write_a = next_q_abs >= 0;
                const uint32_t asb = (uint32_t)p.a_rows * 128u;
                const int b_exp = write_a ? ilog2f_floor(sc.sb) : 0;
                float4 rv[8], cv[8];
                int row = slot16;
                bool active = misc->win[sl][row] >= 0;
                int kwin = active ? misc->win[sl][row] : 0;
                const float* cw = p.cb + ((size_t)q_abs * p.K + kwin) * d;
                seg8_load(rt, row, cw, sub * 4, rv, cv);
#pragma unroll 1
                for (; row < rows_eff; row += ROWS_PER_PASS) {
                    // operand scale of this frame's next stage (from a bound known before the pass)
                    bool force_exact = false;
                    int a_exp = 0;
                    float sa = 0.f;
                    if (write_a) {
                        a_exp = pick_row_exp(misc->row_amax[sl][row] + sc.cmax_q, b_exp, force_exact);
                        sa = exp2i(a_exp);
                    }
                    const bool ok = frame_ok(row);
                    float* ssum = (active && p.stats_sum && ok) ? p.stats_sum + ((size_t)q_abs * p.K + kwin) * d : nullptr;
                    float sq = 0.f;
                    seg8_consume(a_tile, rt, row, ssum, sub * 4, active, write_a, sa, sq, asb, rv, cv);
                    for (int c0 = sub * 4 + 256; c0 < d; c0 += 256) {
                        seg8_load(rt, row, cw, c0, rv, cv);
                        seg8_consume(a_tile, rt, row, ssum, c0, active, write_a, sa, sq, asb, rv, cv);
                    }
                    // next frame: its loads fly while this frame is finished
                    const int nrow = row + ROWS_PER_PASS;
                    const bool cur_active = active;
                    const int cur_kwin = kwin;
                    if (nrow < rows_eff) {
                        active = misc->win[sl][nrow] >= 0;
                        kwin = active ? misc->win[sl][nrow] : 0;
                        cw = p.cb + ((size_t)q_abs * p.K + kwin) * d;
                        seg8_load(rt, nrow, cw, sub * 4, rv, cv);
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    if (cur_active && sub == 0) {
                        misc->row_amax[sl][row] = sqrtf(sq) * 1.00002f;
                        if (write_a) {
                            if (!isfinite(sq)) force_exact = true;
                            write_row_consts(p, misc, sl, row, d, sq, force_exact, a_exp, b_exp, sc.sb, sc.cnmax);
                        }
                    }
                    post_row(row, cur_active, cur_kwin, sq);
                }
            } else {
#pragma unroll 1
                for (int row = slot16; row < rows_eff; row += ROWS_PER_PASS) {
                    const bool active = misc->win[sl][row] >= 0;
                    const int kwin = active ? misc->win[sl][row] : 0;
                    float sq;
                    apply_row(p, misc, a_tile, rt, sl, row, active, frame_ok(row), kwin, q_abs, next_q_abs, sub, &sq, sc);
                    post_row(row, active, kwin, sq);
                }
            }
            const long long ts2 = clock64();
            t_score += ts1 - ts0;
            t_apply += ts2 - ts1;
            // (one shared-memory double atomic per warp: they are CAS loops, and 32 lanes on one address were 2 % of the
            // kernel's stall samples)
            commit_local += __shfl_xor_sync(0xffffffffu, commit_local, 8);
            commit_local += __shfl_xor_sync(0xffffffffu, commit_local, 16);
            if ((u & 31) == 0 && commit_local != 0.0) atomicAdd(&misc->commit_acc[q], commit_local);
            commit_local = 0.0;
            long long t2 = clock64();
            // ---------------- frames the filter could not certify: exact scan of every code
            const int n_dirty = misc->dirty_count[sl];
            if (n_dirty > 0) {
                const int Kv = (int)p.cb_meta[(size_t)q_abs * META_STRIDE + 3];
#pragma unroll 1
                for (int i = 0; i < n_dirty; ++i) {
                    const int row = misc->dirty_rows[sl][i];
                    // only the columns whose minimum is in reach can hold a candidate; each update warp scans
                    // an equal share of the stage's loads
                    const uint32_t cols = misc->dirty_cols[sl][i];
                    const int n_it = (Kv + 15) / 16, per_w = (n_it + UPD_WARPS - 1) / UPD_WARPS;
                    const int it0 = min(n_it, uwarp * per_w), it1 = min(n_it, it0 + per_w);
                    const ScoreIdx b = exact_scan_cols(rt.at(row, 0), cbq, d, it0, it1, cols, Kv, lane);
                    if (lane == 0) {
                        misc->dirty_s[uwarp] = b.s;
                        misc->dirty_k[uwarp] = b.k;
                    }
                    named_bar_sync(BAR_UPD, UPD_THREADS);
                    if (uwarp == (i % UPD_WARPS)) {
                        float bs = misc->dirty_s[0];
                        int bk = misc->dirty_k[0];
                        for (int w = 1; w < UPD_WARPS; ++w)
                            if (better(misc->dirty_s[w], misc->dirty_k[w], bs, bk)) {
                                bs = misc->dirty_s[w];
                                bk = misc->dirty_k[w];
                            }
                        if (bk < 0 || bk >= Kv) bk = 0;
                        const long long n = n0 + row;
                        float sq;
                        apply_row(p, misc, a_tile, rt, sl, row, lane < 8, frame_ok(row), bk, q_abs, next_q_abs, sub, &sq, sc);
                        if (lane == 0 && frame_ok(row)) {
                            __stcs(p.idx + n * nq + q, (long long)bk);
                            atomicAdd(&misc->commit_acc[q], (double)sq);
                            if (p.stats_cnt) atomicAdd(p.stats_cnt + (size_t)q_abs * p.K + bk, 1.f);
                        }
                    }
                    named_bar_sync(BAR_UPD, UPD_THREADS);
                }
                if (u == 0) misc->dirty_count[sl] = 0;
            }
            long long t3 = clock64();
            if (next_q_abs >= 0) {
                fence_proxy_async_smem();
                mbar_arrive(&misc->a_ready[sl]);
            } else {
                // ---------------- last stage: xq = x - final residual, then the slot takes its next tile
                named_bar_sync(BAR_UPD, UPD_THREADS);  // dirty rows were finished by other warps
                if (row_major) {
#pragma unroll 1
                    for (int row = slot16; row < rows_eff; row += ROWS_PER_PASS) {
                        const long long n = n0 + row;
                        if (frame_ok(row)) {
                            const long long off = p.ad.row(n);
                            // four pieces of x and of the residual per lane in flight, then the stores (a load
                            // behind every store would be a dependent round trip: xq may alias as far as nvcc knows)
#pragma unroll 1
                            for (int c0 = sub * 4; c0 < d; c0 += 128) {
                                float4 xv[4], rv[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int c = min(c0 + i * 32, d - 32 + sub * 4);  // (every element assigned)
                                    xv[i] = __ldcs(reinterpret_cast<const float4*>(p.x + off + c));
                                    rv[i] = *reinterpret_cast<const float4*>(rt.at(row, c));
                                }
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    if (c0 + i * 32 < d) {
                                        float4 o;
                                        o.x = xv[i].x - rv[i].x;
                                        o.y = xv[i].y - rv[i].y;
                                        o.z = xv[i].z - rv[i].z;
                                        o.w = xv[i].w - rv[i].w;
                                        __stcs(reinterpret_cast<float4*>(p.xq + off + c0 + i * 32), o);
                                    }
                                }
                            }
                        }
                    }
                } else {
                    // frames-fastest storage: 32 frames x 16 features per unit (16 loads of x and four 16-byte loads of
                    // the residual per lane in flight, then the stores)
                    const int nrb = rows_eff / 32, units = nrb * (d / 16);
#pragma unroll 1
                    for (int unit = uwarp; unit < units; unit += UPD_WARPS) {
                        const int row = (unit % nrb) * 32 + (u & 31), c0 = (unit / nrb) * 16;
                        if (frame_ok(row)) {
                            const long long off = p.ad.row(n0 + row) + (long long)c0 * p.ad.sd;
                            float4 r4[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) r4[j] = *reinterpret_cast<const float4*>(rt.at(row, c0 + 4 * j));
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = __ldcs(p.x + off + (long long)j * p.ad.sd);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                __stcs(p.xq + off + (long long)(4 * j + 0) * p.ad.sd, v[4 * j + 0] - r4[j].x);
                                __stcs(p.xq + off + (long long)(4 * j + 1) * p.ad.sd, v[4 * j + 1] - r4[j].y);
                                __stcs(p.xq + off + (long long)(4 * j + 2) * p.ad.sd, v[4 * j + 2] - r4[j].z);
                                __stcs(p.xq + off + (long long)(4 * j + 3) * p.ad.sd, v[4 * j + 3] - r4[j].w);
                            }
                        }
                    }
                }
                named_bar_sync(BAR_UPD, UPD_THREADS);  // residual slot is rewritten by the next tile
                const int next_i = job.i + nslots;
                if (next_i < n_local) load_tile(sl, blockIdx.x + next_i * gridDim.x);
            }
            t_upd += t2 - t1 + (clock64() - t3);
            t_dirty += t3 - t2;
            n_dirty_tot += n_dirty;
            ++n_jobs;
        }
        if (u < nq) {
            const double v = misc->commit_acc[u];
            if (v != 0.0) atomicAdd(p.commit_sq + u, v);
        }
        if (p.prof) {
            if (u == 0) atomicAdd(p.prof + 6, n_two_tot);
            if (u == 0) {
                atomicAdd(p.prof + 2, (unsigned long long)t_upd);
                atomicAdd(p.prof + 3, (unsigned long long)t_dirty);
                atomicAdd(p.prof + 4, n_dirty_tot);
                atomicAdd(p.prof + 5, n_jobs);
                atomicAdd(p.prof + 7, (unsigned long long)t_wait);
                atomicAdd(p.prof + 8, (unsigned long long)t_score);
                atomicAdd(p.prof + 9, (unsigned long long)t_apply);
                atomicAdd(p.prof + 10, n_score_pass);
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace rvq

// ------------------------------------------------------------------------------------------ host side
using namespace rvq;

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

struct SmemPlan {
    uint32_t off_B, off_misc, total;
    int nstage, nslots;
};

SmemPlan plan_smem(int d, int smem_max, int a_rows) {
    SmemPlan s{};
    const uint32_t a_bytes = (uint32_t)(d / KSLICE) * (uint32_t)a_rows * 128u;
    const uint32_t misc_bytes = (uint32_t)((sizeof(Misc) + 1023) / 1024 * 1024);
    // two tiles in flight (ping-pong between scan and update warps) if >= 2 ring stages still fit: the epilogue,
    // not the MMA, is the long pole, so a shallow codebook ring costs less than serialising scan and update
    s.nslots = (2 * a_bytes + misc_bytes + 2 * B_STAGE_BYTES + 1024 <= (uint32_t)smem_max) ? 2 : 1;
    s.off_B = (uint32_t)s.nslots * a_bytes;
    const uint32_t fixed = s.off_B + misc_bytes + 1024;
    int ns = ((uint32_t)smem_max > fixed) ? (int)(((uint32_t)smem_max - fixed) / B_STAGE_BYTES) : 0;
    if (ns > MAX_STAGES_RING) ns = MAX_STAGES_RING;
    s.nstage = ns;
    s.off_misc = s.off_B + (uint32_t)ns * B_STAGE_BYTES;
    s.total = s.off_misc + misc_bytes + 1024;
    return s;
}
}  // namespace

int rvq_tc_workspace_bytes(int d, int num_sms, size_t* out) {
    // per-CTA residual scratch (used when the tile does not fit in shared memory)
    *out = (size_t)num_sms * 2 * TILE_M * d * sizeof(float) + 256 + 256;
    return RVQ_OK;
}

int rvq_launch_tc(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d, int nq,
                  int K, int q_begin, const float* cb, const void* cb_op, int nq_total, const float* cb_norm,
                  const float* cb_meta, float* xq, long long* idx, double* commit_sq, float* stats_sum,
                  float* stats_cnt, void* ws, size_t ws_bytes, float* dbg_scores, float* dbg_rowscale, int cluster,
                  unsigned long long* prof, cudaStream_t st) {
    if (K > 32 * CHUNK_N) {
        set_error("rvq_encode: at most %d codes per stage are supported (got %d)", 32 * CHUNK_N, K);
        return RVQ_ERR_ARG;
    }
    if (nq > MAX_NQ) {
        set_error("rvq_encode: at most %d stages are supported (got %d)", MAX_NQ, nq);
        return RVQ_ERR_ARG;
    }
    int dev = 0, num_sms = 0, smem_max = 0;
    RVQ_CUDA(cudaGetDevice(&dev));
    RVQ_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    RVQ_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int Kpad = round_up(K, CHUNK_N);
    // Tile size.  Small calls (the reference's training and inference shapes: 136 ... 4000 frames) would fill a handful
    // of SMs with whole 128-frame tiles and leave the rest idle, while a tile-stage costs the update warps one pass per
    // 32 frames: they get tiles of ceil(N / SMs) frames rounded up to a whole pass.  The MMA still runs M = 128, but the
    // operand tile keeps only a_rows = 32 | 64 | 128 rows per slice in shared memory (the lanes beyond read whatever
    // follows and are ignored), which leaves room for a deeper codebook ring (d = 512: 2 -> 4 stages).
    // (Measured and NOT done: 64-frame tiles in two slots for large calls at d = 512, where 128-row tiles leave room for
    // one slot only.  The update warps serve both slots and their cost per job is mostly fixed latency - re-rank 14 k
    // cycles for 10 frames as for 21 - so the model default 8 x 1024 x 512 went from 4.94 to 6.00 ms,
    // profiles/r2u_phase.log.)
    int tile_rows = TILE_M;
    if (!dbg_scores) {
        if (N < (long long)num_sms * TILE_M) {
            const long long per_sm = (N + num_sms - 1) / num_sms;
            tile_rows = (int)((per_sm + 31) / 32 * 32);
            if (tile_rows > TILE_M) tile_rows = TILE_M;
        }
    }
    const int a_rows = tile_rows <= 32 ? 32 : (tile_rows <= 64 ? 64 : TILE_M);
    const SmemPlan sp = plan_smem(d, smem_max, a_rows);
    if (sp.nstage < 2) {
        set_error("rvq_encode: d=%d leaves no room for the codebook ring in %d bytes of shared memory", d, smem_max);
        return RVQ_ERR_ARG;
    }
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) {
        set_error("rvq_encode: cuTensorMapEncodeTiled is not available from the driver");
        return RVQ_ERR_CUDA;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)nq_total * Kpad};
    const cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
    // Codebook multicast across a cluster is available here too (cluster = 2 | 4) but measured slower for
    // this kernel (C4 shape: 5.5 -> 3.1 M frames/s: a slow exact scan in one CTA stalls its whole cluster).
    const int CL = dbg_scores ? 1 : ((cluster == 2 || cluster == 4) ? cluster : 1);
    const cuuint32_t box[2] = {(cuuint32_t)KSLICE, (cuuint32_t)(CHUNK_N / CL)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(cb_op), gdim, gstride, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("rvq_encode: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return RVQ_ERR_CUDA;
    }
    const int num_tiles = (int)((N + tile_rows - 1) / tile_rows);
    // persistent grid of whole clusters: as many as the tiles need, at most one CTA per SM
    const int want_clusters = (num_tiles + CL - 1) / CL, max_clusters = num_sms / CL;
    const int grid = (want_clusters < max_clusters ? want_clusters : max_clusters) * CL;
    EncParams p{};
    p.cluster = CL;
    p.tile_rows = tile_rows;
    p.a_rows = a_rows;
    p.x = x;
    p.N = N;
    p.ad = RowAddrT{L, sb, sl, sd};
    p.d = d;
    p.nq = nq;
    p.K = K;
    p.Kpad = Kpad;
    p.q_begin = q_begin;
    p.cb = cb;
    p.cb_norm = cb_norm;
    p.cb_meta = cb_meta;
    p.xq = xq;
    p.idx = idx;
    p.commit_sq = commit_sq;
    p.stats_sum = stats_sum;
    p.stats_cnt = stats_cnt;
    p.num_tiles = num_tiles;
    p.nstage = sp.nstage;
    p.nslots = sp.nslots;
    p.r_pitch = d;
    p.off_B = sp.off_B;
    p.off_misc = sp.off_misc;
    p.dbg_scores = dbg_scores;
    p.dbg_rowscale = dbg_rowscale;
    p.prof = prof;  // 32 counters (RVQ_FLAG_COUNTERS) or null
    {
        const size_t need = (size_t)grid * 2 * TILE_M * d * sizeof(float);
        uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
        if (!ws || base + need > reinterpret_cast<uintptr_t>(ws) + ws_bytes) {
            set_error("rvq_encode: workspace too small (%zu bytes given, %zu needed)", ws_bytes, need + 256);
            return RVQ_ERR_WORKSPACE;
        }
        p.r_scratch = reinterpret_cast<float*>(base);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = sp.total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (dbg_scores) {
        RVQ_CUDA(cudaFuncSetAttribute(rvq_encode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sp.total));
        RVQ_CUDA(cudaLaunchKernelEx(&cfg, rvq_encode_tc_kernel<true>, tmap, p));
    } else {
        RVQ_CUDA(cudaFuncSetAttribute(rvq_encode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sp.total));
        RVQ_CUDA(cudaLaunchKernelEx(&cfg, rvq_encode_tc_kernel<false>, tmap, p));
    }
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}
