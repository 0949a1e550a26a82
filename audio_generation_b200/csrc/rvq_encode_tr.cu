// rvq_encode_tr.cu -- K1 (+K2 fused) for d <= 128: the residual-vector-quantization stage loop as ONE persistent
// sm_100a kernel whose fp32 residual tile lives in TENSOR MEMORY for all stages.  Replaces the per-stage
// {distance, argmin, gather, subtract, EMA statistics} loop of som_quantizer.ResidualQuantizer.forward
// (called at /root/reference/networks/vae.py:315-318).
//
// Why a second kernel: the generic kernel (rvq_encode_tc.cu) keeps the residual in an L2-resident scratch and
// streams the codebook once per tile, which costs ~450 KB of L2->SM traffic per 128-frame tile-stage at
// d = 128 -- more than the ~43 B/cycle an SM gets from L2 allows at the target rate.  Here
//   * TMEM columns [0, 256)   = two 128-column fp32 accumulators (128 codes per MMA, double buffered),
//   * TMEM columns [256, 512) = the fp32 residual of the two 128-frame tiles in flight (d columns each);
//     an update thread owns ONE frame (TMEM lane) and reads / rewrites its residual with tcgen05.ld / .st,
//   * the selected fp32 code vectors, the frames x and the outputs xq move between global and shared memory
//     as per-frame bulk copies (cp.async.bulk) through one padded staging buffer, so no LSU wavefront is spent
//     on uncoalesced global accesses,
//   * EMA sums leave as bulk reductions (cp.reduce.async.bulk .add.f32) of the staged residual rows.
// What remains per tile-stage is the fp16 codebook stream (2 K d bytes) and the gathered code rows (512 d bytes).
//
// Warp roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 / 8-11 = scan groups 0 / 1 (accumulator buffers 0 / 1), warps 12-15 / 16-19 = update groups of
// tile slots 0 / 1 (thread = frame).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "exact.cuh"
#include "encode_common.cuh"

namespace rvq {
namespace tr {

#ifndef RVQ_TR_CH
#define RVQ_TR_CH 128
#endif
constexpr int CH = RVQ_TR_CH;  // codes per MMA / accumulator buffer (64 or 128)
constexpr uint32_t NSLICE_BYTES = CH * 32;  // norm slice of one chunk
constexpr int SCAN_WARP0 = 4;
constexpr int SCAN_THREADS = 256;
constexpr int UPD_WARP0 = 12;
constexpr int GRP_THREADS = 128;  // one update group = 4 warps = the 128 TMEM lanes
constexpr int NUM_THREADS = UPD_WARP0 * 32 + 2 * GRP_THREADS;  // 640
constexpr int MAX_RING = 8;
constexpr int MAX_NQ = 64;
constexpr uint32_t B_STAGE_BYTES = CH * KSLICE * 2;  // 16 KiB
constexpr uint32_t BAR_SCAN = 1;                     // named barrier of the 256 scan threads
constexpr uint32_t BAR_GRP0 = 2;                     // + slot: named barrier of one update group
constexpr uint32_t TMEM_RES_COL = 256;               // first residual column
constexpr int RS_ROWS = 16;  // re-rank entries (residual row + four candidates) per round: 8 lanes per entry

struct Params {
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
    int num_tiles, nstage, nslots, pitch;  // pitch: floats per staging row (d + 4)
    int cluster;                           // CTAs per cluster sharing the codebook stream (TMA multicast)
    uint32_t off_stg, off_B, off_misc;  // A tiles (one per slot) at offset 0
    uint32_t off_tab;                   // two allowance byte tables of Kpad bytes each (0: looked up in global memory)
    unsigned long long* prof;              // [16] cycle / event counters (RVQ_PROFILE=1) or null
};

struct __align__(16) Misc {
    uint64_t full[MAX_RING], empty[MAX_RING], tmem_full[2], tmem_empty[2], a_ready[2], scan_done[2];
    uint64_t stg_full[2], stg_free;
    uint64_t rc_ready[2];  // paired mode: this CTA's row constants of a job are written (a_ready lives in the leader)
    // norm term as one extra K = 16 MMA step, no-swizzle K-major operands (8-row x 16-byte core matrices):
    // (B side: the chunk's norm slices ride in the ring stage of its last slice)
    alignas(128) uint8_t a_extra[2][4096];  // A: per tile slot, row = {2^(a-b+11), 2^(a-b+1), 2^(a-b-4), 2^14, 0...}
    uint32_t tmem_base;
    float row_na[2][TILE_M], row_delta[2][TILE_M], row_amax[2][TILE_M], row_rs[2][TILE_M];  // per tile slot
    float grp_best[2][2][TILE_M];   // [job parity][scan group][frame]: best score the group saw
    float grp_x[2][2][TILE_M];      // [job parity][scan group][frame]: allowance of that score's code (k0_bound)
    uint16_t wbest[2][2][TILE_M];   // [slot][scan group][frame]: the code that scored it (approximate argmin)
    float vbest[2][2][TILE_M];      // [slot][scan group][frame]: that score
    // candidate sets, double buffered by stage parity: the verification of stage q may still read them while the
    // scan of stage q + 1 writes
    uint32_t g_rows[2][2][2][TILE_M];  // [slot][stage parity][group][frame]: loads that may hold a candidate
    uint16_t g_cols[2][2][2][TILE_M];  // [slot][stage parity][group][frame]: columns that may hold a candidate
    int win[2][TILE_M];             // [slot][frame]: selected code
    int n_special[2], n_dirty[2];
    uint32_t pairs[2][RS_ROWS * 4];      // [slot][entry * 4 + t]: code scored
    float pair_score[2][RS_ROWS * 4];    // its exact score
    uint16_t special_rows[2][4 * TILE_M];  // [slot][entry]: frame | block << 8 | 0x8000 if it needs the exact scan
    float red_s[2][4];
    int red_k[2][4];
    double commit_acc[MAX_NQ];
    float2 stage_u[MAX_NQ];  // units of the allowance byte table per stage (meta[7], meta[5]; k0_bound)
};


// kStats: EMA statistics requested (compile-time so that each instantiation carries one apply path only: the
// update threads are register-bound)
// kPair: the two CTAs of a cluster drive their tensor cores as ONE cta_group::2 MMA (M = 256: 128 frames per CTA,
// N = 128 codes split 64 / 64 between the CTAs' shared memory).  A 128 x 128 x 16 MMA whose operands both sit in
// one CTA's shared memory reads 8 KiB per 64 cycles = the SM's whole 128 B/clk of shared-memory bandwidth, on top of
// which the TMA writes the very same B bytes: 108 KiB per 128-code chunk, 864 cycles at best for 576 cycles of tensor
// work.  Paired, every SM reads its A (4 KiB) and only its half of B (2 KiB) per MMA and receives only that half
// from the TMA: 72 KiB per chunk = the 576 cycles.  The leader (even CTA) issues; the peer's frames take part through
// cluster-scope mbarrier arrivals on the leader's barriers and multicast tcgen05.commit.
template <bool kStats, bool kPair>
__global__ void __launch_bounds__(NUM_THREADS, 1)
rvq_encode_tr_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_n,
                     const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_b = smem + p.off_B;
    float* staging = reinterpret_cast<float*>(smem + p.off_stg);
    Misc* misc = reinterpret_cast<Misc*>(smem + p.off_misc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = p.d, nq = p.nq;
    const int n_ks = d / KSLICE;
    const int n_chunks = p.Kpad / CH;
    const int nstage = p.nstage;
    const uint32_t a_tile_bytes = (uint32_t)n_ks * A_SLICE_BYTES;
    // every CTA of a cluster walks the same job sequence (tiles past the end are empty: all frames invalid)
    const int n_local = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int CL = p.cluster;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    const int nslots = p.nslots;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_init(&misc->full[i], 1);
            // one tcgen05.commit arrive per CTA of the cluster (paired: ONE multicast commit of the leader)
            mbar_init(&misc->empty[i], kPair ? 1u : (uint32_t)CL);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&misc->tmem_full[i], 1);
            mbar_init(&misc->tmem_empty[i], kPair ? 8 : 4);  // one arrive per scan warp of the group (of both CTAs)
            mbar_init(&misc->a_ready[i], kPair ? 2 * GRP_THREADS : GRP_THREADS);
            mbar_init(&misc->scan_done[i], SCAN_THREADS);
            mbar_init(&misc->stg_full[i], GRP_THREADS);
            mbar_init(&misc->rc_ready[i], GRP_THREADS);
            misc->n_special[i] = 0;
            misc->n_dirty[i] = 0;
        }
        mbar_init(&misc->stg_free, GRP_THREADS);
        for (int i = 0; i < MAX_NQ; ++i) misc->commit_acc[i] = 0.0;
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 2 * 4096 / 16; i += NUM_THREADS)
        reinterpret_cast<uint4*>(&misc->a_extra[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < nq; i += NUM_THREADS) {
        const float* mq = p.cb_meta + (size_t)(p.q_begin + i) * META_STRIDE;
        misc->stage_u[i] = make_float2(mq[7], mq[5]);
    }
    // allowance byte table of the first stage (the scan threads keep the next stage's table one step ahead)
    uint8_t* xtab = smem + p.off_tab;
    if (p.off_tab) {
        const NormLayout nl0(p.cb_norm, (int)p.cb_meta[4], p.Kpad);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(nl0.xb + (size_t)p.q_begin * p.Kpad);
        for (int i = threadIdx.x; i < p.Kpad / 4; i += NUM_THREADS) reinterpret_cast<uint32_t*>(xtab)[i] = __ldg(src + i);
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_b);
        if constexpr (kPair) tma_prefetch_desc(&tmap_n);
    }
    if (warp == 2) {
        if constexpr (kPair)
            tmem_alloc_pair<512>(&misc->tmem_base);
        else
            tmem_alloc<512>(&misc->tmem_base);
    }
#ifdef RVQ_DEBUG_HANG
    if (threadIdx.x == 0 && blockIdx.x < 2)
        printf("block %d barriers: full 0x%x empty 0x%x tmem_full 0x%x tmem_empty 0x%x a_ready 0x%x scan_done 0x%x "
               "stg_full 0x%x stg_free 0x%x\n", (int)blockIdx.x, smem_u32(misc->full), smem_u32(misc->empty),
               smem_u32(misc->tmem_full), smem_u32(misc->tmem_empty), smem_u32(misc->a_ready),
               smem_u32(misc->scan_done), smem_u32(misc->stg_full), smem_u32(&misc->stg_free));
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peers' barriers are initialised before anything is multicast to them
    tc_fence_after_sync();
    const uint32_t tmem_base = misc->tmem_base;

    // Register budget: 640 threads x 96 = 61440 registers at launch; setmaxnreg.inc can only take what
    // setmaxnreg.dec released inside this CTA: 40*128 (control) + 88*256 (scan) + 128*256 (update) = 60416.
    // setmaxnreg is warpgroup-aligned: one instruction per warpgroup, before its warps split into roles.
    if (warp < SCAN_WARP0) {
        reg_dealloc<40>();
        if (warp == 0 && kPair) {
            // ======================================================= TMA producer, paired: MY half of every chunk
            // (64 codes x d features + their norm slices) into MY shared memory, bytes counted on the LEADER's barrier
            if (elect_one()) {
                uint32_t st = 0, ph = 0;
                const uint32_t stage_bytes = (uint32_t)n_ks * (B_STAGE_BYTES / 2) + NSLICE_BYTES / 2;
                const int nq_prep = (int)p.cb_meta[4];  // stages prepared: the norm slices follow their norms
                const int nrow_base = nq_prep * p.Kpad / 64;  // norm slices as rows of 256 bytes of the cb_norm buffer
                for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
                    const int q_abs = p.q_begin + job.q;
                    const int row0 = q_abs * p.Kpad + (int)crank * (CH / 2);
                    const int nrow0 = nrow_base + q_abs * n_chunks * (int)(NSLICE_BYTES / 256) + (int)crank * (int)(NSLICE_BYTES / 512);
                    for (int c = 0; c < n_chunks; ++c) {
                        mbar_wait(&misc->empty[st], ph ^ 1);  // the pair's MMAs on this slot have retired
                        if (crank == 0) mbar_arrive_expect_tx(&misc->full[st], 2u * stage_bytes);
                        uint8_t* dst = smem_b + (size_t)st * stage_bytes;
                        for (int ks = 0; ks < n_ks; ++ks)
                            tma_load_2d_pair(dst + (size_t)ks * (B_STAGE_BYTES / 2), &tmap_b, &misc->full[st], ks * KSLICE,
                                             row0 + c * CH);
                        tma_load_2d_pair(dst + (size_t)n_ks * (B_STAGE_BYTES / 2), &tmap_n, &misc->full[st], 0,
                                         nrow0 + c * (int)(NSLICE_BYTES / 256));
                        if (++st == (uint32_t)nstage) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1 && kPair) {
            // ======================================================= MMA issuer, paired: the leader CTA's elected thread
            if (crank == 0 && elect_one()) {
                const uint32_t idesc = umma_idesc_f16(0 /*fp16*/, 2 * TILE_M, CH);
                const uint32_t stage_bytes = (uint32_t)n_ks * (B_STAGE_BYTES / 2) + NSLICE_BYTES / 2;
                const uint32_t smem_b_u32 = smem_u32(smem_b);
                uint32_t g = 0, aphase = 0, st = 0, ph = 0;
                long long t_aready = 0;
                for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
                    const int sl = job.slot % nslots;
                    const long long tw = clock64();
                    mbar_wait_cluster(&misc->a_ready[sl], (aphase >> sl) & 1);  // both CTAs' update groups
                    t_aready += clock64() - tw;
                    aphase ^= 1u << sl;
                    tc_fence_after_sync();
                    const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem + (size_t)sl * a_tile_bytes));
                    const uint64_t adesc_x = umma_desc_nosw(smem_u32(misc->a_extra[sl]), 128, 256);
                    for (int c = 0; c < n_chunks; ++c, ++g) {
                        const uint32_t buf = g & 1, use = g >> 1;
                        mbar_wait_cluster(&misc->tmem_empty[buf], (use & 1) ^ 1);  // both CTAs' scan groups
                        mbar_wait(&misc->full[st], ph);
                        tc_fence_after_sync();
                        const uint32_t tmem_d = tmem_base + buf * CH;
                        const uint32_t sbase = smem_b_u32 + st * stage_bytes;
                        uint64_t adesc = adesc0;
                        for (int ks = 0; ks < n_ks; ++ks) {
                            const uint64_t bdesc = umma_desc_sw128(sbase + (uint32_t)ks * (B_STAGE_BYTES / 2));
                            umma_f16_ss_pair(tmem_d, adesc, bdesc, idesc, ks != 0);
                            umma_f16_ss_pair(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
                            umma_f16_ss_pair(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
                            umma_f16_ss_pair(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
                            adesc += (uint64_t)(A_SLICE_BYTES >> 4);
                        }
                        // the norm term: + A_extra . B_extra^T, my 64 codes' slices behind my half of the chunk
                        umma_f16_ss_pair(tmem_d, adesc_x,
                                         umma_desc_nosw(sbase + (uint32_t)n_ks * (B_STAGE_BYTES / 2), 128, 256), idesc, 1);
                        umma_commit_pair_mc(&misc->empty[st], 3);        // both CTAs' ring slots
                        umma_commit_pair_mc(&misc->tmem_full[buf], 3);   // both CTAs' scan groups
                        if (++st == (uint32_t)nstage) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
                if (p.prof) atomicAdd(p.prof + 18, (unsigned long long)t_aready);
            }
            __syncwarp();
        }
        if constexpr (!kPair) {
        // ring stage = one 64-feature slice of a 128-code chunk (16 KiB) + room for the chunk's norm slices (4 KiB),
        // which travel with the chunk's LAST slice: the MMA thread waits on ONE barrier per slice and issues nothing
        // but MMAs and commits (round 1 had it issue the norm-slice copy and wait on a third barrier per chunk: ~150
        // single-thread instructions, ~1.2 k cycles, per 576 cycles of tensor work)
        const uint32_t stage_bytes = B_STAGE_BYTES + NSLICE_BYTES;
        if (warp == 0) {
            // ======================================================= TMA producer (codebook slices + norm slices)
            if (elect_one()) {
                uint32_t st = 0, ph = 0;
                const uint32_t part_bytes = B_STAGE_BYTES / (uint32_t)CL;
                const int part_rows = CH / CL;
                const int nq_prep = (int)p.cb_meta[4];  // stages prepared: the norm slices follow their norms
                const uint8_t* nbase = reinterpret_cast<const uint8_t*>(p.cb_norm + (size_t)nq_prep * p.Kpad);
                for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
                    const int row0 = (p.q_begin + job.q) * p.Kpad + (int)crank * part_rows;
                    const uint8_t* nsrc = nbase + (size_t)(p.q_begin + job.q) * n_chunks * NSLICE_BYTES;
                    for (int c = 0; c < n_chunks; ++c) {
                        for (int ks = 0; ks < n_ks; ++ks) {
                            const bool last = ks == n_ks - 1;
                            mbar_wait(&misc->empty[st], ph ^ 1);  // every CTA of the cluster has consumed the slot
                            mbar_arrive_expect_tx(&misc->full[st], B_STAGE_BYTES + (last ? NSLICE_BYTES : 0u));
                            uint8_t* sbase = smem_b + (size_t)st * stage_bytes;
                            uint8_t* dst = sbase + crank * part_bytes;
                            if (CL > 1)  // my 1/CL of the slice goes to every CTA of the cluster
                                tma_load_2d_mc(dst, &tmap_b, &misc->full[st], ks * KSLICE, row0 + c * CH, cmask);
                            else
                                tma_load_2d(dst, &tmap_b, &misc->full[st], ks * KSLICE, row0 + c * CH);
                            if (last)
                                bulk_load_1d(sbase + B_STAGE_BYTES, nsrc + (size_t)c * NSLICE_BYTES, NSLICE_BYTES,
                                             &misc->full[st]);
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
            // ======================================================= MMA issuer
            // ONE elected thread runs the whole loop: every instruction of a lone warp costs ~8 cycles of latency,
            // and a 128 x 128 x 16 MMA lasts 64 cycles, so the issue loop must stay at a handful of instructions
            // per MMA (no divisions, no per-iteration warp syncs, descriptors advanced by additions).
            if (elect_one()) {
                const uint32_t idesc = umma_idesc_f16(0 /*fp16*/, TILE_M, CH);
                const uint32_t smem_b_u32 = smem_u32(smem_b);
                uint32_t g = 0, aphase = 0, st = 0, ph = 0;
                long long t_aready = 0;
                for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
                    const int sl = job.slot % nslots;
                    const long long tw = clock64();
                    mbar_wait(&misc->a_ready[sl], (aphase >> sl) & 1);
                    t_aready += clock64() - tw;
                    aphase ^= 1u << sl;
                    tc_fence_after_sync();
                    const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem + (size_t)sl * a_tile_bytes));
                    const uint64_t adesc_x = umma_desc_nosw(smem_u32(misc->a_extra[sl]), 128, 256);
                    for (int c = 0; c < n_chunks; ++c, ++g) {
                        const uint32_t buf = g & 1, use = g >> 1;
                        mbar_wait(&misc->tmem_empty[buf], (use & 1) ^ 1);  // the scan group has released this buffer
                        const uint32_t tmem_d = tmem_base + buf * CH;
                        uint64_t adesc = adesc0;
                        for (int ks = 0; ks < n_ks; ++ks) {
                            mbar_wait(&misc->full[st], ph);
                            tc_fence_after_sync();
                            const uint32_t sbase = smem_b_u32 + st * stage_bytes;
                            const uint64_t bdesc = umma_desc_sw128(sbase);
                            // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                            umma_f16_ss(tmem_d, adesc, bdesc, idesc, ks != 0);
                            umma_f16_ss(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
                            umma_f16_ss(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
                            umma_f16_ss(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
                            if (ks == n_ks - 1)  // the norm term: + A_extra . B_extra^T (write_norm_slice, rvq_aux.cu)
                                umma_f16_ss(tmem_d, adesc_x, umma_desc_nosw(sbase + B_STAGE_BYTES, 128, 256), idesc, 1);
                            // frees the ring slot (in every CTA of the cluster) when these MMAs retire
                            if (CL > 1)
                                umma_commit_mc(&misc->empty[st], cmask);
                            else
                                umma_commit(&misc->empty[st]);
                            if (ks == n_ks - 1) umma_commit(&misc->tmem_full[buf]);
                            adesc += (uint64_t)(A_SLICE_BYTES >> 4);
                            if (++st == (uint32_t)nstage) {
                                st = 0;
                                ph ^= 1u;
                            }
                        }
                    }
                }
                if (p.prof) atomicAdd(p.prof + 18, (unsigned long long)t_aready);
            }
            __syncwarp();
        }
        }  // !kPair
    } else if (warp < UPD_WARP0) {
        reg_dealloc<88>();
        // =========================================================== scan groups (argmin epilogue)
        const int e = threadIdx.x - SCAN_WARP0 * 32;  // 0..255
        const int grp = e >> 7;                       // scan group = accumulator buffer
        const int my_row = (warp & 3) * 32 + lane;    // TMEM lane owned by this thread
        uint32_t g = 0, aphase = 0, jpar = 0, step = 0;
        long long t_scan = 0, t_wait = 0, t_full = 0;
        const NormLayout nl(p.cb_norm, (int)p.cb_meta[4], p.Kpad);
        const int tab_words = p.Kpad / 4;
        for (JobIter job(n_local, nq, nslots); job.valid(); job.next(), jpar ^= 1u) {
            const int sl = job.slot % nslots;
            long long t0 = clock64();
            // Both slots run the same stage in consecutive jobs ("step").  The first job of a step copies the NEXT
            // step's byte table into the other buffer: every reader of that buffer's old content (step - 1) is past
            // the scan barrier that ended step - 1, and the scan barrier of this job orders the copy before the first
            // lookup of the next step.  The word is loaded here and stored after the scan.
            const bool tab_copy = p.off_tab != 0 && job.slot == 0;
            step += job.slot == 0 ? 1u : 0u;
            const uint32_t* tab_src = reinterpret_cast<const uint32_t*>(
                nl.xb + (size_t)(p.q_begin + (job.q + 1 == nq ? 0 : job.q + 1)) * p.Kpad);
            uint32_t tab_w = 0;
            if (tab_copy && e < tab_words) tab_w = __ldg(tab_src + e);
            mbar_wait(kPair ? &misc->rc_ready[sl] : &misc->a_ready[sl], (aphase >> sl) & 1);  // row constants visible
            aphase ^= 1u << sl;
            const float delta = misc->row_delta[sl][my_row];
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
                tc_fence_after_sync();
                t_full += clock64() - tw0;
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + grp * CH;
                uint32_t va[16], vb[16];
                tmem_ld_32x16(taddr, va);
                uint32_t it = (uint32_t)c * (CH / 16);
#pragma unroll 1
                for (int cb = 0; cb < CH; cb += 32, it += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x16(taddr + cb + 16, vb);
                    scan16_2d_raw(va, it, Cm, m1, m2, m3, m4, nullptr);
                    tmem_ld_wait();
                    if (cb + 32 < CH) tmem_ld_32x16(taddr + cb + 32, va);
                    scan16_2d_raw(vb, it + 1, Cm, m1, m2, m3, m4, nullptr);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kPair)
                        mbar_arrive_cluster(&misc->tmem_empty[grp], 0);  // the leader's barrier counts both CTAs
                    else
                        mbar_arrive(&misc->tmem_empty[grp]);
                }
            }
            // ---------------- stage end: exchange the groups' best scores, list this group's candidates
            float vb_ = fminf(fminf(Cm[0], Cm[1]), Cm[2]);
#pragma unroll
            for (int j = 3; j < 15; j += 2) vb_ = fminf(fminf(vb_, Cm[j]), Cm[j + 1]);
            vb_ = fminf(vb_, Cm[15]);
            misc->grp_best[jpar][grp][my_row] = vb_;
            float xm;
            {
                // the code behind the group's best score: the load of the smallest load minimum, the first column
                // that attains the column minimum
                int jmin = 15;
#pragma unroll
                for (int j = 14; j >= 0; --j) jmin = (Cm[j] == vb_) ? j : jmin;
                const uint32_t kb = (__float_as_uint(m1) & IT_MASK) * 16u + (uint32_t)jmin;
                misc->wbest[sl][grp][my_row] = (uint16_t)kb;
                misc->vbest[sl][grp][my_row] = vb_;
                // allowance of the code behind that score (zero unless it is above the stage's norm cap, k0_bound)
                const float2 u = misc->stage_u[job.q];
                const float xu = fmaf(misc->row_rs[sl][my_row], u.x, misc->row_na[sl][my_row] * u.y);
                // (two call sites: the shared-memory one compiles to LDS instead of generic loads)
                xm = p.off_tab ? best_allowance(vb_, jmin, m1, m2, m3, m4, xu, xtab + ((step & 1u) ^ 1u) * (uint32_t)p.Kpad,
                                                p.Kpad - 1)
                               : best_allowance(vb_, jmin, m1, m2, m3, m4, xu,
                                                nl.xb + (size_t)(p.q_begin + job.q) * p.Kpad, p.Kpad - 1);
                misc->grp_x[jpar][grp][my_row] = xm;
                if (tab_copy) {
                    uint32_t* dst = reinterpret_cast<uint32_t*>(xtab + (step & 1u) * (uint32_t)p.Kpad);
                    if (e < tab_words) dst[e] = tab_w;
                    for (int w = e + 2 * GRP_THREADS; w < tab_words; w += 2 * GRP_THREADS) dst[w] = __ldg(tab_src + w);
                }
            }
            named_bar_sync(BAR_SCAN, SCAN_THREADS);
            {
                const float ob = misc->grp_best[jpar][grp ^ 1][my_row];
                const float best = fminf(vb_, ob);
                const float xo = misc->grp_x[jpar][grp ^ 1][my_row];
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
                                      ((__float_as_uint(m3) & IT_MASK) << 18) | (nr << 27) | (over ? G_OVER : 0u) |
                                      (nofilter ? G_NOFILTER : 0u);
                misc->g_rows[sl][job.q & 1][grp][my_row] = rows;
                misc->g_cols[sl][job.q & 1][grp][my_row] = (uint16_t)cols;
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
        // =========================================================== update groups (thread = frame)
        const int s = (warp - UPD_WARP0) >> 2;    // tile slot served by this group
        const int row = (warp & 3) * 32 + lane;   // frame of the tile = TMEM lane (UPD_WARP0 % 4 == 0)
        const int gw = warp & 3;                  // warp inside the group
        if (s < nslots) {
            const uint32_t t_r = tmem_base + ((uint32_t)(gw * 32) << 16) + TMEM_RES_COL + (uint32_t)(s * d);
            float* stg_row = staging + (size_t)row * p.pitch;
            float* stg_warp = staging + (size_t)(gw * 32) * p.pitch;  // first staging row of this warp's frames
            uint8_t* a_tile = smem + (size_t)s * a_tile_bytes;
            // A-tile address pieces of my frame: 16-byte chunk j of a 128-byte swizzle row sits at (j ^ (row & 7)) << 4
            uint8_t* a_row = a_tile + (uint32_t)row * 128u;
            const uint32_t rx = ((uint32_t)row & 7u) << 4;
            const bool row_major = (p.ad.sd == 1);
            const uint32_t bar_grp = BAR_GRP0 + (uint32_t)s;
            const uint32_t row_bytes = (uint32_t)d * 4u;
            uint32_t stg_par = 0, sphase = 0;
            long long t_upd = 0, t_wait = 0, t_rank = 0, t_gather = 0, t_apply = 0, t_tail = 0, t_acq = 0;
            long long t_k1 = 0, t_k2 = 0, t_k3 = 0;
            unsigned long long n_dirty_tot = 0, n_multi_tot = 0, n_jobs = 0;

            // the staging buffer is handed from critical section to critical section in one global order:
            // prologue of slot 0, prologue of slot 1, then the jobs in JobIter order
            auto acquire = [&](int ticket) {
                if (ticket > 0) mbar_wait(&misc->stg_free, (uint32_t)(ticket - 1) & 1u);
            };
            auto release = [&]() { mbar_arrive(&misc->stg_free); };
            auto wait_staging = [&]() {
                mbar_wait(&misc->stg_full[s], stg_par);
                stg_par ^= 1u;
            };
            // 32 consecutive features (c0 .. c0+31, inside one 64-feature slice) of my frame -> fp16 operand
            auto store_a = [&](int c0, const uint32_t (&v)[32], float sa) {
                uint8_t* base = a_row + (uint32_t)(c0 >> 6) * A_SLICE_BYTES;
                const uint32_t j0 = ((uint32_t)c0 >> 3) & 7u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 pk;
                    __half2 h;
                    h = __floats2half2_rn(__uint_as_float(v[8 * j + 0]) * sa, __uint_as_float(v[8 * j + 1]) * sa);
                    pk.x = *reinterpret_cast<const uint32_t*>(&h);
                    h = __floats2half2_rn(__uint_as_float(v[8 * j + 2]) * sa, __uint_as_float(v[8 * j + 3]) * sa);
                    pk.y = *reinterpret_cast<const uint32_t*>(&h);
                    h = __floats2half2_rn(__uint_as_float(v[8 * j + 4]) * sa, __uint_as_float(v[8 * j + 5]) * sa);
                    pk.z = *reinterpret_cast<const uint32_t*>(&h);
                    h = __floats2half2_rn(__uint_as_float(v[8 * j + 6]) * sa, __uint_as_float(v[8 * j + 7]) * sa);
                    pk.w = *reinterpret_cast<const uint32_t*>(&h);
                    *reinterpret_cast<uint4*>(base + ((((j0 + (uint32_t)j) << 4)) ^ rx)) = pk;
                }
            };
            // operand row of the norm term for a frame whose operand exponents are a (row) and b (codes)
            constexpr int NORM_WINDOW_LO = -10;  // below: 2^(a-b-4) leaves the fp16 normal range -> exact scan
            auto store_a_extra = [&](int a_, int b_) {
                const int e = max(NORM_WINDOW_LO, min(ROW_OVER_CODE_MAX, a_ - b_));
                const __half2 h01 = __floats2half2_rn(exp2i(e + 11), exp2i(e + 1));
                const __half2 h23 = __floats2half2_rn(exp2i(e - 4), 16384.f);
                uint4 v;
                v.x = *reinterpret_cast<const uint32_t*>(&h01);
                v.y = *reinterpret_cast<const uint32_t*>(&h23);
                v.z = v.w = 0u;   // fifth column (rs / 64): store_a_rs, once the new residual's norm is known
                *reinterpret_cast<uint4*>(misc->a_extra[s] + (row >> 3) * 256 + (row & 7) * 16) = v;
            };
            // fifth operand column: rs / 64 rounded up, x (-64 xc_k) of the norm slice = the allowance rs * xc_k of the
            // codes above the stage's norm cap (k0_bound in rvq_aux.cu)
            auto store_a_rs = [&](float rs) {
                const __half2 h45 = __halves2half2(__float2half_ru(rs * 0.015625f), __float2half_rn(0.f));
                *reinterpret_cast<uint32_t*>(misc->a_extra[s] + (row >> 3) * 256 + (row & 7) * 16 + 8) =
                    *reinterpret_cast<const uint32_t*>(&h45);
            };
            // Coalesced asynchronous gather of one d-float row per frame of this warp into the staging buffer:
            // lane r's row number (index into `table`, rows of d floats; < 0 = none) is broadcast and the 32 lanes
            // copy 16 bytes each (d <= 128: one instruction per row).  Every thread then arrives on stg_full[s]
            // when ITS copies have landed (count = 128 arrivals).
            // `off_mine`: element offset of my frame inside `table` (< 0 = no frame)
            auto gather_rows = [&](const float* table, long long off_mine) {
                __syncwarp();  // every lane has finished with the staging rows of this warp
                const int lo = (int)(off_mine & 0xffffffffll), hi = (int)(off_mine >> 32);
                const float* src_lane = table + lane * 4;
                float* dst_lane = stg_warp + lane * 4;
                const bool lane_on = lane * 4 < d;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const int rlo = __shfl_sync(0xffffffffu, lo, r), rhi = __shfl_sync(0xffffffffu, hi, r);
                    const long long ro = ((long long)rhi << 32) | (unsigned int)rlo;
                    if (ro >= 0 && lane_on) cp_async16(dst_lane + (size_t)r * p.pitch, src_lane + ro);
                }
                cp_async_arrive_noinc(&misc->stg_full[s]);
            };
            auto gather_codes = [&](const float* cbq_, int w_mine) {
                __syncwarp();
                const float* src_lane = cbq_ + lane * 4;
                float* dst_lane = stg_warp + lane * 4;
                const bool lane_on = lane * 4 < d;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const int wr = __shfl_sync(0xffffffffu, w_mine, r);
                    if (lane_on) cp_async16(dst_lane + (size_t)r * p.pitch, src_lane + (size_t)wr * d);
                }
                cp_async_arrive_noinc(&misc->stg_full[s]);
            };

            // load tile `tile` into this slot (staging held): residual <- x, fp16 operand + row constants of stage 0
            auto load_tile = [&](int tile) {
                const long long n = (long long)tile * TILE_M + row;
                const bool valid = n < p.N;
                const long long off = valid ? p.ad.row(n) : 0;
                if (row_major) {
                    gather_rows(p.x, valid ? off : -1);
                    wait_staging();
                }
                float sq = 0.f, amax = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < d; c0 += 32) {
                    uint32_t v[32];
                    if (row_major) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid) t = *reinterpret_cast<const float4*>(stg_row + c0 + j);
                            v[j + 0] = __float_as_uint(t.x);
                            v[j + 1] = __float_as_uint(t.y);
                            v[j + 2] = __float_as_uint(t.z);
                            v[j + 3] = __float_as_uint(t.w);
                        }
                    } else {
                        // frames-fastest storage (the reference's (B, d, L) tensor): lanes = consecutive frames
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = valid ? __float_as_uint(__ldcs(p.x + off + (long long)(c0 + j) * p.ad.sd)) : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f = __uint_as_float(v[j]);
                        sq = fmaf(f, f, sq);
                        amax = fmaxf(amax, fabsf(f));
                    }
                    tmem_st_32x32(t_r + c0, v);
                }
                tmem_st_wait();
                const float* mq = p.cb_meta + (size_t)p.q_begin * META_STRIDE;
                const float sb = mq[0], cnmax = mq[1];
                const int b = ilog2f_floor(sb);
                bool force_exact = !isfinite(sq);
                const int a = pick_row_exp(amax, b, force_exact);
                if (a - b < NORM_WINDOW_LO) force_exact = true;
                const float sa = exp2i(a);
#pragma unroll 1
                for (int c0 = 0; c0 < d; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_r + c0, v);
                    tmem_ld_wait();
                    store_a(c0, v, sa);
                }
                store_a_extra(a, b);
                float na, delta, rs;
                row_consts(d, sq, force_exact, a, b, sb, cnmax, na, delta, rs);
                store_a_rs(rs);
                misc->row_amax[s][row] = amax;
                misc->row_na[s][row] = na;
                misc->row_delta[s][row] = delta;
                misc->row_rs[s][row] = rs;
                fence_proxy_async_smem();
            };

            int ticket = 0;
            // ---------------- prologue: first tile of each slot
            for (int ps = 0; ps < nslots && ps < n_local; ++ps, ++ticket) {
                if (ps != s) continue;
                acquire(ticket);
                load_tile(blockIdx.x + ps * gridDim.x);
                release();
                if constexpr (kPair) {
                    mbar_arrive(&misc->rc_ready[s]);            // my CTA's scan groups
                    mbar_arrive_cluster(&misc->a_ready[s], 0);  // the leader issues the pair's MMAs
                } else
                    mbar_arrive(&misc->a_ready[s]);
            }
            constexpr bool stats = kStats;
            // a job takes a turn at the staging buffer only if it uses it: EMA statistics (the bulk reduction reads
            // the residual rows from shared memory) or the last stage (x / xq rows, next tile)
            for (JobIter job(n_local, nq, nslots); job.valid(); ticket += (stats || job.q + 1 == nq) ? 1 : 0, job.next()) {
                if (job.slot % nslots != s) continue;
                const int q = job.q, q_abs = p.q_begin + q;
                const bool staged = stats || q + 1 == nq;
                const int tile = blockIdx.x + job.i * gridDim.x;
                const long long n = (long long)tile * TILE_M + row;
                const bool valid = n < p.N;
                const int next_q_abs = (q + 1 < nq) ? q_abs + 1 : -1;
                const float* cbq = p.cb + (size_t)q_abs * p.K * d;
                const int Kv = (int)p.cb_meta[(size_t)q_abs * META_STRIDE + 3];
                const long long tj0 = clock64();
                mbar_wait(&misc->scan_done[s], sphase);
                sphase ^= 1u;
                const long long tj1 = clock64();
                // ---------------- classify my frame: certified (one candidate), several candidates, exact scan.
                // A frame with nc candidates takes ceil(nc / 4) re-rank entries (four candidates per entry).
                const uint32_t r0 = misc->g_rows[s][q & 1][0][row], r1 = misc->g_rows[s][q & 1][1][row];
                const uint32_t c0m = misc->g_cols[s][q & 1][0][row], c1m = misc->g_cols[s][q & 1][1][row];
                const int n0 = (int)((r0 >> 27) & 3u) * __popc(c0m), n1 = (int)((r1 >> 27) & 3u) * __popc(c1m);
                const int nc = n0 + n1;
                int w = 0, mypos = -1, myk = 0;
                bool dirty = false;
                if (((r0 | r1) & G_OVER) || nc == 0 || nc > 16) {
                    dirty = true;
                    myk = 1;
                } else if (nc == 1) {
                    w = n0 ? (int)((r0 & IT_MASK) * 16u) + __ffs(c0m) - 1 : (int)((r1 & IT_MASK) * 16u) + __ffs(c1m) - 1;
                    w = max(0, min(w, Kv - 1));  // cannot bind (padding codes score 2^100); keeps the gather in bounds
                } else {
                    myk = (nc + 3) >> 2;
                }
                if (myk > 0) {
                    mypos = atomicAdd(&misc->n_special[s], myk);
                    for (int i = 0; i < myk; ++i)
                        misc->special_rows[s][mypos + i] = (uint16_t)(row | (i << 8) | (dirty ? 0x8000 : 0));
                    if (dirty) atomicAdd(&misc->n_dirty[s], 1);
                }
                // ---------------- exact scores, RS_ROWS entries per round: the frames expose their residual rows
                // in shared memory and every 8-lane group re-scores the four candidates of one entry
                // exposed rows live in this slot's A tile: its MMAs have retired, the next operand is written later
                float* rstage = reinterpret_cast<float*>(a_tile);
                float bs = __int_as_float(0x7f800000);
                int kwin = 0x7fffffff;
                int n_special = RS_ROWS, n_dirty = 0;  // read after the first barrier of round 0
                const long long tk1 = clock64();
                t_k1 += tk1 - tj1;
#pragma unroll 1
                for (int base = 0; base < n_special; base += RS_ROWS) {
                    // my first entry inside this round (if any) is where my row is exposed
                    const int lo = max(mypos, base), hi = min(mypos + myk, base + RS_ROWS);
                    const bool in_round = mypos >= 0 && lo < hi;
                    if (__any_sync(0xffffffffu, in_round)) {
                        float* dst = rstage + (size_t)(in_round ? lo - base : 0) * p.pitch;
#pragma unroll 1
                        for (int c0 = 0; c0 < d; c0 += 32) {
                            uint32_t v[32];
                            tmem_ld_32x32(t_r + c0, v);
                            tmem_ld_wait();
                            if (in_round) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            }
                        }
                    }
                    named_bar_sync(bar_grp, GRP_THREADS);  // entries listed (round 0) and rows exposed
                    if (base == 0) {
                        n_special = misc->n_special[s];
                        n_dirty = misc->n_dirty[s];
                    }
                    const long long tk2 = clock64();
                    const int n_round = min(RS_ROWS, n_special - base);
                    {
                        const int sub = lane & 7, slot16 = (gw * 32 + lane) >> 3;
                        const int ent = slot16 < n_round ? misc->special_rows[s][base + slot16] : 0x8000;
                        if (__any_sync(0xffffffffu, !(ent & 0x8000))) {
                            const int rr = ent & 0x7f, blk = (ent >> 8) & 0x7f;
                            const CandSet cs(misc->g_rows[s][q & 1][0][rr], misc->g_cols[s][q & 1][0][rr], misc->g_rows[s][q & 1][1][rr],
                                             misc->g_cols[s][q & 1][1][rr]);
                            // the frame's row sits at its first entry of this round
                            const int first = max(slot16 - blk, 0);
                            const float* rrow = rstage + (size_t)first * p.pitch;
                            int k[4];
                            const float* cc[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                k[t] = cs.code(4 * blk + t, Kv - 1);
                                cc[t] = cbq + (size_t)k[t] * d;
                            }
                            float sv[4];
                            exact_score8_n<4>(rrow, cc, d, sub, sv);
                            if (sub == 0 && !(ent & 0x8000)) {
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    misc->pair_score[s][slot16 * 4 + t] = sv[t];
                                    misc->pairs[s][slot16 * 4 + t] = (uint32_t)k[t];
                                }
                            }
                        }
                    }
                    named_bar_sync(bar_grp, GRP_THREADS);
                    if (in_round && !dirty) {
                        for (int e = lo; e < hi; ++e) {
                            const int blk = e - mypos;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float sc = misc->pair_score[s][(e - base) * 4 + t];
                                const int kk = (int)misc->pairs[s][(e - base) * 4 + t];
                                if (4 * blk + t < nc && better(sc, kk, bs, kwin)) {
                                    bs = sc;
                                    kwin = kk;
                                }
                            }
                        }
                    }
                    const long long tk3 = clock64();
                    if (base == 0) {
                        t_k2 += tk2 - tk1;
                        t_k3 += tk3 - tk2;
                    }
                    // frames the filter could not bound: exact scan of the columns in reach, all four warps
                    if (n_dirty > 0) {
#pragma unroll 1
                        for (int i = 0; i < n_round; ++i) {
                            const int ent = misc->special_rows[s][base + i];
                            if (!(ent & 0x8000)) continue;
                            const int rr = ent & 0x7f;
                            uint32_t cols = (uint32_t)misc->g_cols[s][q & 1][0][rr] | (uint32_t)misc->g_cols[s][q & 1][1][rr];
                            if (cols == 0) cols = 0xFFFFu;
                            const int n_it = (Kv + 15) / 16, per_w = (n_it + 3) / 4;
                            const int it0 = min(n_it, gw * per_w), it1 = min(n_it, it0 + per_w);
                            const ScoreIdx bsc =
                                exact_scan_cols(rstage + (size_t)i * p.pitch, cbq, d, it0, it1, cols, Kv, lane);
                            if (lane == 0) {
                                misc->red_s[s][gw] = bsc.s;
                                misc->red_k[s][gw] = bsc.k;
                            }
                            named_bar_sync(bar_grp, GRP_THREADS);
                            if (gw == 0 && lane == 0) {
                                float rs_ = misc->red_s[s][0];
                                int rk_ = misc->red_k[s][0];
                                for (int ww = 1; ww < 4; ++ww)
                                    if (better(misc->red_s[s][ww], misc->red_k[s][ww], rs_, rk_)) {
                                        rs_ = misc->red_s[s][ww];
                                        rk_ = misc->red_k[s][ww];
                                    }
                                if (rk_ < 0 || rk_ >= Kv) rk_ = 0;
                                misc->win[s][rr] = rk_;
                            }
                            named_bar_sync(bar_grp, GRP_THREADS);
                            ++n_dirty_tot;
                        }
                    }
                    if (base + RS_ROWS < n_special) named_bar_sync(bar_grp, GRP_THREADS);  // exposed rows are replaced
                }
                if (mypos >= 0) w = dirty ? misc->win[s][row] : min(kwin, Kv - 1);
                if (gw == 0 && lane == 0) {  // next use is after the next scan of this slot
                    misc->n_special[s] = 0;
                    misc->n_dirty[s] = 0;
                }
                if (p.prof) {
                    const int w_spec = misc->vbest[s][0][row] <= misc->vbest[s][1][row] ? misc->wbest[s][0][row]
                                                                                        : misc->wbest[s][1][row];
                    const unsigned miss = __ballot_sync(0xffffffffu, valid && w_spec != w);
                    if (lane == 0 && miss) atomicAdd(p.prof + 20, (unsigned long long)__popc(miss));
                }
                const long long tj2 = clock64();
                // the selected code vector: with statistics it goes through the staging buffer (whose row then
                // becomes the source of the bulk reduction); without, my thread reads its 4d bytes straight into
                // registers with 256-bit loads, two 32-feature pieces ahead of their use
                const float* crow = cbq + (size_t)w * d;
                uint32_t ca[kStats ? 1 : 16], cb_[kStats ? 1 : 16];
                if (staged) acquire(ticket);
                if constexpr (kStats) {
                    gather_codes(cbq, w);
                } else {
                    ldg_nc_16f(crow, ca);
                    ldg_nc_16f(crow + 16, cb_);
                }
                const long long tj3 = clock64();
                // ---------------- constants of the next stage's operand (scale chosen from a bound known now)
                const bool write_a = next_q_abs >= 0;
                float sb = 1.f, cnmax = 0.f, sa = 0.f;
                int a = 0, b = 0;
                bool force_exact = false;
                if (write_a) {
                    const float* mq = p.cb_meta + (size_t)next_q_abs * META_STRIDE;
                    sb = mq[0];
                    cnmax = mq[1];
                    b = ilog2f_floor(sb);
                    a = pick_row_exp(misc->row_amax[s][row] + p.cb_meta[(size_t)q_abs * META_STRIDE + 2], b, force_exact);
                    if (a - b < NORM_WINDOW_LO) force_exact = true;
                    sa = exp2i(a);
                    store_a_extra(a, b);
                }
                if (valid) {
                    p.idx[n * nq + q] = w;
                    if (p.stats_cnt) atomicAdd(p.stats_cnt + (size_t)q_abs * p.K + w, 1.f);
                }
                if constexpr (kStats) wait_staging();  // the selected code vectors have landed
                const long long tj4 = clock64();
                // ---------------- r <- r - c (fp32, tensor memory), next operand row, statistics row
                float sq = 0.f;
                auto apply32 = [&](uint32_t (&v)[32], int c0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 cv = *reinterpret_cast<const float4*>(stg_row + c0 + j);
                        // the staging row becomes the stage INPUT residual (source of the EMA bulk reduction)
                        if (stats) *reinterpret_cast<uint4*>(stg_row + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        const float n0f = __uint_as_float(v[j + 0]) - cv.x;
                        const float n1f = __uint_as_float(v[j + 1]) - cv.y;
                        const float n2f = __uint_as_float(v[j + 2]) - cv.z;
                        const float n3f = __uint_as_float(v[j + 3]) - cv.w;
                        sq = fmaf(n0f, n0f, sq);
                        sq = fmaf(n1f, n1f, sq);
                        sq = fmaf(n2f, n2f, sq);
                        sq = fmaf(n3f, n3f, sq);
                        v[j + 0] = __float_as_uint(n0f);
                        v[j + 1] = __float_as_uint(n1f);
                        v[j + 2] = __float_as_uint(n2f);
                        v[j + 3] = __float_as_uint(n3f);
                    }
                    tmem_st_32x32(t_r + c0, v);
                    if (write_a) store_a(c0, v, sa);
                };
                // 16 features of my frame with the code vector piece in registers
                auto apply16r = [&](uint32_t (&v)[16], const uint32_t (&c)[kStats ? 1 : 16], int c0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float nf = __uint_as_float(v[j]) - __uint_as_float(c[kStats ? 0 : j]);
                        sq = fmaf(nf, nf, sq);
                        v[j] = __float_as_uint(nf);
                    }
                    tmem_st_32x16(t_r + c0, v);
                    if (write_a) {
                        uint8_t* base = a_row + (uint32_t)(c0 >> 6) * A_SLICE_BYTES;
                        const uint32_t j0 = ((uint32_t)c0 >> 3) & 7u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint4 pk;
                            __half2 hh;
                            hh = __floats2half2_rn(__uint_as_float(v[8 * j + 0]) * sa, __uint_as_float(v[8 * j + 1]) * sa);
                            pk.x = *reinterpret_cast<const uint32_t*>(&hh);
                            hh = __floats2half2_rn(__uint_as_float(v[8 * j + 2]) * sa, __uint_as_float(v[8 * j + 3]) * sa);
                            pk.y = *reinterpret_cast<const uint32_t*>(&hh);
                            hh = __floats2half2_rn(__uint_as_float(v[8 * j + 4]) * sa, __uint_as_float(v[8 * j + 5]) * sa);
                            pk.z = *reinterpret_cast<const uint32_t*>(&hh);
                            hh = __floats2half2_rn(__uint_as_float(v[8 * j + 6]) * sa, __uint_as_float(v[8 * j + 7]) * sa);
                            pk.w = *reinterpret_cast<const uint32_t*>(&hh);
                            *reinterpret_cast<uint4*>(base + ((((j0 + (uint32_t)j) << 4)) ^ rx)) = pk;
                        }
                    }
                };
                if constexpr (kStats) {
                    // two 32-feature pieces in flight: the next TMEM load is issued before the current piece is used
                    uint32_t va[32], vb[32];
                    tmem_ld_32x32(t_r, va);
#pragma unroll 1
                    for (int c0 = 0; c0 < d; c0 += 64) {
                        tmem_ld_wait();
                        tmem_ld_32x32(t_r + c0 + 32, vb);  // d is a multiple of 64
                        apply32(va, c0);
                        tmem_ld_wait();
                        if (c0 + 64 < d) tmem_ld_32x32(t_r + c0 + 64, va);
                        apply32(vb, c0 + 32);
                    }
                } else {
                    // 16-feature pieces: TMEM load one piece ahead, code loads two pieces ahead
                    uint32_t va[16], vb[16];
                    tmem_ld_32x16(t_r, va);
#pragma unroll 1
                    for (int c0 = 0; c0 < d; c0 += 32) {
                        tmem_ld_wait();
                        tmem_ld_32x16(t_r + c0 + 16, vb);
                        apply16r(va, ca, c0);
                        if constexpr (!kStats) {
                            if (c0 + 32 < d) ldg_nc_16f(crow + c0 + 32, ca);
                        }
                        tmem_ld_wait();
                        if (c0 + 32 < d) tmem_ld_32x16(t_r + c0 + 32, va);
                        apply16r(vb, cb_, c0 + 16);
                        if constexpr (!kStats) {
                            if (c0 + 32 < d) ldg_nc_16f(crow + c0 + 48, cb_);
                        }
                    }
                }
                if (stats) {
                    fence_proxy_async_smem();
                    if (valid) {
                        bulk_reduce_add_f32(p.stats_sum + ((size_t)q_abs * p.K + w) * d, stg_row, row_bytes);
                        bulk_commit();
                    }
                }
                tmem_st_wait();
                misc->row_amax[s][row] = sqrtf(sq) * 1.00002f;  // ||r'||_2 >= max|r'|
                if (write_a) {
                    if (!isfinite(sq)) force_exact = true;
                    float na, delta, rs;
                    row_consts(d, sq, force_exact, a, b, sb, cnmax, na, delta, rs);
                    store_a_rs(rs);
                    misc->row_na[s][row] = na;
                    misc->row_delta[s][row] = delta;
                    misc->row_rs[s][row] = rs;
                }
                {
                    // commit-loss partial: sum over the valid frames of this warp
                    double cs = valid ? (double)sq : 0.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
                    if (lane == 0 && cs != 0.0) atomicAdd(&misc->commit_acc[q], cs);
                }
                if (stats && valid) bulk_wait_read0();  // the reduction has read my staging row
                const long long tj5 = clock64();
                if (write_a) {
                    fence_proxy_async_smem();
                    if (staged) release();
                    if constexpr (kPair) {
                    mbar_arrive(&misc->rc_ready[s]);            // my CTA's scan groups
                    mbar_arrive_cluster(&misc->a_ready[s], 0);  // the leader issues the pair's MMAs
                } else
                    mbar_arrive(&misc->a_ready[s]);
                } else {
                    // ---------------- last stage: xq = x - final residual, then the slot takes its next tile
                    const long long off = valid ? p.ad.row(n) : 0;
                    if (row_major) {
                        if (stats) fence_proxy_async_smem();
                        gather_rows(p.x, valid ? off : -1);
                        wait_staging();
                    }
#pragma unroll 1
                    for (int c0 = 0; c0 < d; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_r + c0, v);
                        tmem_ld_wait();
                        if (row_major) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 xv = *reinterpret_cast<const float4*>(stg_row + c0 + j);
                                xv.x -= __uint_as_float(v[j + 0]);
                                xv.y -= __uint_as_float(v[j + 1]);
                                xv.z -= __uint_as_float(v[j + 2]);
                                xv.w -= __uint_as_float(v[j + 3]);
                                *reinterpret_cast<float4*>(stg_row + c0 + j) = xv;
                            }
                        } else if (valid) {
                            // all 32 loads before the first store: x and xq may alias as far as the compiler knows,
                            // and a load behind every store made this 32 dependent memory round trips per piece
                            float xv[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) xv[j] = __ldcs(p.x + off + (long long)(c0 + j) * p.ad.sd);
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                __stcs(p.xq + off + (long long)(c0 + j) * p.ad.sd, xv[j] - __uint_as_float(v[j]));
                        }
                    }
                    if (row_major) {
                        // coalesced store of the warp's 32 output rows (512 bytes per instruction)
                        __syncwarp();
                        const unsigned long long dp = valid ? reinterpret_cast<unsigned long long>(p.xq + off) : 0ull;
#pragma unroll 4
                        for (int r = 0; r < 32; ++r) {
                            const unsigned long long dpr = __shfl_sync(0xffffffffu, dp, r);
                            if (dpr != 0ull && lane * 4 < d)
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(dpr) + lane * 4) =
                                    *reinterpret_cast<const float4*>(stg_warp + (size_t)r * p.pitch + lane * 4);
                        }
                        __syncwarp();
                    }
                    const int next_i = job.i + nslots;
                    if (next_i < n_local) {
                        load_tile(blockIdx.x + next_i * gridDim.x);
                        release();
                        if constexpr (kPair) {
                    mbar_arrive(&misc->rc_ready[s]);            // my CTA's scan groups
                    mbar_arrive_cluster(&misc->a_ready[s], 0);  // the leader issues the pair's MMAs
                } else
                    mbar_arrive(&misc->a_ready[s]);
                    } else {
                        release();
                    }
                }
                const long long tj6 = clock64();
                t_wait += tj1 - tj0;
                t_rank += tj2 - tj1;
                t_acq += tj3 - tj2;
                t_gather += tj4 - tj3;
                t_apply += tj5 - tj4;
                t_tail += tj6 - tj5;
                t_upd += tj6 - tj1;
                n_multi_tot += n_special;
                ++n_jobs;
            }
            bulk_wait0();  // my bulk reductions are complete before the kernel ends
            if (p.prof && gw == 0 && lane == 0) {
                atomicAdd(p.prof + 2, (unsigned long long)t_upd);
                atomicAdd(p.prof + 3, (unsigned long long)t_acq);
                atomicAdd(p.prof + 4, n_dirty_tot);
                atomicAdd(p.prof + 5, n_jobs);
                atomicAdd(p.prof + 6, n_multi_tot);
                atomicAdd(p.prof + 7, (unsigned long long)t_wait);
                atomicAdd(p.prof + 8, (unsigned long long)t_rank);
                atomicAdd(p.prof + 9, (unsigned long long)t_apply);
                atomicAdd(p.prof + 10, (unsigned long long)t_gather);
                atomicAdd(p.prof + 12, (unsigned long long)t_tail);
                atomicAdd(p.prof + 13, (unsigned long long)t_k1);
                atomicAdd(p.prof + 14, (unsigned long long)t_k2);
                atomicAdd(p.prof + 15, (unsigned long long)t_k3);
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
    if (threadIdx.x < nq) {
        const double v = misc->commit_acc[threadIdx.x];
        if (v != 0.0) atomicAdd(p.commit_sq + threadIdx.x, v);
    }
    if (warp == 2) {
        tc_fence_after_sync();
        if constexpr (kPair)
            tmem_dealloc_pair<512>(tmem_base);
        else
            tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace tr
}  // namespace rvq

// ------------------------------------------------------------------------------------------ host side
using namespace rvq;

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled_tr() {
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
}  // namespace

// d in {64, 128}: both tiles' residuals fit in TMEM columns [256, 512)
bool rvq_tr_supported(int d) { return d == 64 || d == 128; }

int rvq_launch_tr(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d, int nq,
                  int K, int q_begin, const float* cb, const void* cb_op, int nq_total, const float* cb_norm,
                  const float* cb_meta, float* xq, long long* idx, double* commit_sq, float* stats_sum,
                  float* stats_cnt, int cluster, unsigned long long* prof, cudaStream_t st) {
    if (K > 32 * CHUNK_N) {
        set_error("rvq_encode: at most %d codes per stage are supported (got %d)", 32 * CHUNK_N, K);
        return RVQ_ERR_ARG;
    }
    if (nq > tr::MAX_NQ) {
        set_error("rvq_encode: at most %d stages are supported (got %d)", tr::MAX_NQ, nq);
        return RVQ_ERR_ARG;
    }
    int dev = 0, num_sms = 0, smem_max = 0;
    RVQ_CUDA(cudaGetDevice(&dev));
    RVQ_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    RVQ_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int Kpad = round_up(K, CHUNK_N);
    tr::Params p{};
    p.nslots = 2;
    p.pitch = d + 4;
    const uint32_t a_bytes = (uint32_t)(d / KSLICE) * A_SLICE_BYTES;
    // cluster: 0 / 1 = independent CTAs (default: fastest measured, profiles/r2d_*); 2 = the two CTAs of a cluster as
    // ONE cta_group::2 MMA; 4 = independent tensor cores, one codebook stream multicast to four CTAs
    const bool pair = cluster == 2;
    const int CL = pair ? 2 : (cluster == 4 ? 4 : 1);
    p.cluster = CL;
    const int stg_rows = TILE_M;  // staging buffer of one tile
    const uint32_t stg_bytes = (uint32_t)((stg_rows * p.pitch * 4 + 1023) / 1024 * 1024);
    const uint32_t misc_bytes = (uint32_t)((sizeof(tr::Misc) + 1023) / 1024 * 1024);
    // ring stage: one 64-feature slice of a chunk; paired: MY half of a whole chunk (all slices + norm slices)
    const uint32_t stage_bytes = pair ? (uint32_t)(d / KSLICE) * (tr::B_STAGE_BYTES / 2) + tr::NSLICE_BYTES / 2
                                      : tr::B_STAGE_BYTES + tr::NSLICE_BYTES;
    // allowance byte tables of two stages (k0_bound): in shared memory while they are small, else read from global
    const uint32_t tab_bytes = 2u * (uint32_t)Kpad <= 4096u ? 2u * (uint32_t)Kpad : 0u;
    p.off_stg = (uint32_t)p.nslots * a_bytes;
    p.off_B = p.off_stg + stg_bytes;
    const uint32_t fixed = p.off_B + misc_bytes + tab_bytes + 1024;
    int ns = ((uint32_t)smem_max > fixed) ? (int)(((uint32_t)smem_max - fixed) / stage_bytes) : 0;
    if (ns > tr::MAX_RING) ns = tr::MAX_RING;
    if (ns < 2) {
        set_error("rvq_encode: d=%d leaves no room for the codebook ring in %d bytes of shared memory", d, smem_max);
        return RVQ_ERR_ARG;
    }
    p.nstage = ns;
    p.off_misc = (p.off_B + (uint32_t)ns * stage_bytes + 1023u) / 1024u * 1024u;
    p.off_tab = tab_bytes ? p.off_misc + misc_bytes : 0u;
    const uint32_t smem_total = p.off_misc + misc_bytes + tab_bytes + 1024;

    EncodeTiledFn encode = get_encode_tiled_tr();
    if (!encode) {
        set_error("rvq_encode: cuTensorMapEncodeTiled is not available from the driver");
        return RVQ_ERR_CUDA;
    }
    CUtensorMap tmap, tmap_n;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)nq_total * Kpad};
        const cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
        const cuuint32_t box[2] = {(cuuint32_t)KSLICE, (cuuint32_t)(tr::CH / (pair ? 2 : CL))};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(cb_op), gdim, gstride, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("rvq_encode: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
            return RVQ_ERR_CUDA;
        }
    }
    {
        // the cb_norm buffer as rows of 256 bytes: the paired producer fetches the norm slices of its 64 codes
        // (2 KiB = 8 rows) with the tensor-map TMA, the only bulk copy that may complete on the peer CTA's mbarrier.
        // The extent is a bound only: the buffer's real size (which depends on the number of prepared stages) is
        // known to the device (cb_meta[4]), and no coordinate ever leaves it.
        const cuuint64_t gdim[2] = {256, (cuuint64_t)1 << 24};
        const cuuint64_t gstride[1] = {256};
        const cuuint32_t box[2] = {256, (cuuint32_t)(tr::NSLICE_BYTES / 512)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode(&tmap_n, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<float*>(cb_norm), gdim, gstride, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("rvq_encode: cuTensorMapEncodeTiled (norm slices) failed with CUresult %d", (int)cr);
            return RVQ_ERR_CUDA;
        }
    }
    const int num_tiles = (int)((N + TILE_M - 1) / TILE_M);
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
    p.prof = prof;  // 32 counters (RVQ_FLAG_COUNTERS) or null
    auto kern = stats_sum ? (pair ? tr::rvq_encode_tr_kernel<true, true> : tr::rvq_encode_tr_kernel<true, false>)
                          : (pair ? tr::rvq_encode_tr_kernel<false, true> : tr::rvq_encode_tr_kernel<false, false>);
    RVQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(tr::NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent grid: as many co-resident clusters as the device takes, not more than the tiles need
    int max_clusters = num_sms / CL;
    cfg.gridDim = dim3((unsigned)(max_clusters * CL), 1, 1);
    {
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc > 0 && nc < max_clusters)
            max_clusters = nc;
        (void)cudaGetLastError();
    }
    const int want_clusters = (num_tiles + CL - 1) / CL;
    const int grid = (want_clusters < max_clusters ? want_clusters : max_clusters) * CL;
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    RVQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, tmap_n, p));
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}
