// rvq_encode_fr.cu -- K1 (+K2 fused) for d in {64, 128, 256}: the residual-vector-quantization stage loop as ONE
// persistent sm_100a kernel in which a FRAME IS OWNED BY ONE THREAD for its whole life in the SM.  Replaces the
// per-stage {distance, argmin, gather, subtract, EMA statistics} loop of som_quantizer.ResidualQuantizer.forward
// (called at /root/reference/networks/vae.py:315-318).
//
// Round-1's kernels split the epilogue into "scan" warps and "update" warps that handed every frame over through
// shared memory (candidate lists, named barriers, exposed residual rows): the hand-over chain, not any throughput
// limit, bounded them.  Here the thread that owns TMEM lane `row`
//   * scans the stage's scores of ITS frame straight out of the accumulators (two-dimensional running minimum),
//   * classifies the frame in registers (certified / several candidates / no usable bound),
//   * re-scores uncertified frames exactly with its warp only (rows exposed in the slot's idle A tile; no barrier
//     wider than a warp exists in the worker code),
//   * applies r <- r - c on its own TMEM lane, writes the next fp16 operand row and arrives on the slot's barrier.
// TMEM columns [0, 256) = two 128-code accumulators, [256, 512) = the fp32 residual(s): two 128-frame tile slots
// for d <= 128 (their stages alternate on the tensor pipe: one slot's MMAs run while the other slot updates),
// one slot for d = 256.  x, xq and the selected code vectors move between global memory and registers with 256-bit
// accesses of whole 32-byte sectors (thread = frame), or as coalesced 128-byte rows when the frames are stored
// feature-major (the reference's (B, d, L) tensor, vae.py:313).  No staging buffer, no scratch in global memory.
//
// Warp roles: warp 0 = TMA producer (codebook slices, cluster multicast), warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-7 / 8-11 = worker group of tile slot 0 / 1.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "exact.cuh"
#include "encode_common.cuh"

namespace rvq {
namespace fr {

constexpr int BCH = 128;                     // codes per ring stage (one TMA box column / norm slice)
constexpr uint32_t NSLICE_BYTES = BCH * 32;  // norm slice of one 128-code chunk
constexpr int CTRL_THREADS = 128;            // warps 0-3
constexpr int GRP_THREADS = 128;             // one worker group = 4 warps = the 128 TMEM lanes
constexpr int MAX_RING = 8;
constexpr int MAX_NQ = 64;
constexpr uint32_t B_SLICE_BYTES = BCH * KSLICE * 2;  // 16 KiB: [128 codes x 64 features] fp16, SWIZZLE_128B
constexpr int SCR_ROWS = 16;                         // exposed residual rows per warp (inside its own A-tile rows)
constexpr int NORM_WINDOW_LO = -10;  // below: 2^(a-b-4) leaves the fp16 normal range -> exact scan

#ifndef RVQ_FR_SLOTS
#define RVQ_FR_SLOTS(D) ((D) <= 128 ? 3 : 1)
#endif

template <int D>
struct Cfg {
    // tile slots in flight: TMEM holds 512 columns = two accumulator buffers of CH codes + NSLOTS x D residual
    // columns.  d <= 128: three slots with 64-code accumulators; d = 256: one slot with 128-code accumulators.
    static constexpr int NSLOTS = RVQ_FR_SLOTS(D);
    static constexpr int CH = (2 * 128 + NSLOTS * D <= 512) ? 128 : 64;  // codes per MMA / accumulator buffer
    static constexpr int SUBS = BCH / CH;                                // accumulator chunks per ring stage
    static constexpr uint32_t TMEM_RES_COL = 2 * CH;                     // first residual column
    static constexpr int NUM_THREADS = CTRL_THREADS + NSLOTS * GRP_THREADS;
    static constexpr int N_KS = D / KSLICE;
    static constexpr uint32_t A_TILE_BYTES = (uint32_t)N_KS * A_SLICE_BYTES;
    // one ring stage = up to two 64-feature slices of a 128-code chunk + room for the chunk's norm slice (which
    // travels with the chunk's LAST stage): one mbarrier wait of the MMA warp covers 8 (+1) MMAs
    static constexpr int SL_PER_ST = N_KS < 2 ? N_KS : 2;
    static constexpr int ST_PER_CHUNK = N_KS / SL_PER_ST;
    static constexpr uint32_t STAGE_BYTES = (uint32_t)SL_PER_ST * B_SLICE_BYTES + NSLICE_BYTES;
};

struct Params {
    const float* x;
    long long N;
    RowAddrT ad;
    int nq, K, Kpad, q_begin;
    const float* cb;       // [*, K, d] fp32 master
    const float* cb_norm;  // [*, Kpad] scaled norms, then the norm slices
    const float* cb_meta;  // [*, 8]
    float* xq;
    long long* idx;
    double* commit_sq;
    float* stats_sum;
    float* stats_cnt;
    int num_tiles, nstage;
    int cluster;  // CTAs per cluster sharing the codebook stream (TMA multicast)
    uint32_t off_B, off_misc;  // A tiles (one per slot) at offset 0
    unsigned long long* prof;  // [32] event counters or null
};

struct __align__(16) Misc {
    uint64_t full[MAX_RING], empty[MAX_RING];
    // accumulator hand-over, per (tile slot, buffer): each barrier is waited on by ONE party in strict phase order
    uint64_t tmem_full[3][2], tmem_empty[3][2], a_ready[3];
    // norm term as one extra K = 16 MMA step, no-swizzle K-major operands (8-row x 16-byte core matrices): the B
    // side (norm slices of a chunk) rides in the ring stage, the A side is written by the frame's owner:
    alignas(128) uint8_t a_extra[3][4096];  // per tile slot, row = {2^(a-b+11), 2^(a-b+1), 2^(a-b-4), 2^14, 0...}
    float4 stage_meta[MAX_NQ];              // per stage {2^b, max ||c||_2, max |c|, K_valid}
    uint32_t tmem_base;
    double commit_acc[MAX_NQ];
};

// job = (tile slot, tile, stage) for NS tile slots: groups of NS consecutive local tiles walk their stages together,
// for q: (slot 0, q), (slot 1, q), ...; every role walks the same sequence.
template <int NS>
struct JobIterN {
    int n_local, nq, i, q, slot;  // i = local tile index = group * NS + slot
    __device__ __forceinline__ JobIterN(int n_local_, int nq_) : n_local(n_local_), nq(nq_), i(0), q(0), slot(0) {}
    __device__ __forceinline__ bool valid() const { return i < n_local; }
    __device__ __forceinline__ void next() {
        const int g0 = i - slot;  // first tile of the group
        if (slot + 1 < NS && g0 + slot + 1 < n_local) {
            ++slot;
            ++i;
        } else {
            slot = 0;
            i = g0;
            if (++q == nq) {
                q = 0;
                i = g0 + NS;
            }
        }
    }
};

// Candidate set of one frame, factorised: up to three loads (`it`, 9 bits each in `rows`, count in bits 27-28) x a
// 16-bit column mask.  Same enumeration order as CandSet (encode_common.cuh) restricted to one scan group.
struct Cand1 {
    uint32_t rows, cols;
    int pc, n;
    __device__ __forceinline__ Cand1(uint32_t r, uint32_t c) : rows(r), cols(c) {
        pc = __popc(c);
        n = (int)((r >> 27) & 3u) * pc;
    }
    // code number e of the set, clamped into [0, kmax] (out-of-range e, or an empty set, gives a harmless code)
    __device__ __forceinline__ int code(int e, int kmax) const {
        e = max(0, min(e, n - 1));
        const int pcg = max(pc, 1);
        uint32_t sh = 0;  // at most three loads
        if (e >= pcg) {
            e -= pcg;
            sh = 9;
        }
        if (e >= pcg) {
            e -= pcg;
            sh = 18;
        }
        uint32_t m = cols;
        for (int i = 0; i < e && i < 15; ++i) m &= m - 1;  // drop the e lowest set bits
        const uint32_t it = (rows >> sh) & IT_MASK;
        const int k = (int)(it * 16u) + ((__ffs(m) - 1) & 15);
        return max(0, min(k, kmax));
    }
};

// 256-bit global accesses of 32 consecutive floats with cache hints
__device__ __forceinline__ void ldg_32f(const float* p, uint32_t (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[8 * i + 0]), "=r"(v[8 * i + 1]), "=r"(v[8 * i + 2]), "=r"(v[8 * i + 3]), "=r"(v[8 * i + 4]),
                       "=r"(v[8 * i + 5]), "=r"(v[8 * i + 6]), "=r"(v[8 * i + 7])
                     : "l"(p + 8 * i));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int D, bool kStats>
__global__ void __launch_bounds__(Cfg<D>::NUM_THREADS, 1)
rvq_encode_fr_kernel(const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    constexpr int NSLOTS = Cfg<D>::NSLOTS, CH = Cfg<D>::CH, SUBS = Cfg<D>::SUBS;
    constexpr uint32_t TMEM_RES_COL = Cfg<D>::TMEM_RES_COL;
    constexpr int n_ks = Cfg<D>::N_KS;
    constexpr uint32_t a_tile_bytes = Cfg<D>::A_TILE_BYTES;
    constexpr int SL_PER_ST = Cfg<D>::SL_PER_ST, ST_PER_CHUNK = Cfg<D>::ST_PER_CHUNK;
    constexpr uint32_t STAGE_BYTES = Cfg<D>::STAGE_BYTES;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_b = smem + p.off_B;
    Misc* misc = reinterpret_cast<Misc*>(smem + p.off_misc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = p.nq;
    const int n_chunks = p.Kpad / BCH;  // 128-code ring chunks per stage; even: Kpad is a multiple of 256
    // uses of EACH accumulator buffer per job (the buffers alternate over the accumulator chunks)
    const uint32_t uses_per_job = (uint32_t)(n_chunks * SUBS) >> 1;
    const int nstage = p.nstage;
    // every CTA of a cluster walks the same job sequence (tiles past the end are empty: all frames invalid)
    const int n_local = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int CL = p.cluster;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_init(&misc->full[i], 1);
            mbar_init(&misc->empty[i], (uint32_t)CL);  // one tcgen05.commit arrive per CTA of the cluster
        }
        for (int s = 0; s < 3; ++s) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(&misc->tmem_full[s][b], 1);
                mbar_init(&misc->tmem_empty[s][b], 4);  // one arrive per warp of the consuming group
            }
            mbar_init(&misc->a_ready[s], GRP_THREADS);
        }
        for (int i = 0; i < MAX_NQ; ++i) misc->commit_acc[i] = 0.0;
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 3 * 4096 / 16; i += Cfg<D>::NUM_THREADS)
        reinterpret_cast<uint4*>(&misc->a_extra[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < nq) {
        const float* mq = p.cb_meta + (size_t)(p.q_begin + threadIdx.x) * META_STRIDE;
        misc->stage_meta[threadIdx.x] = make_float4(mq[0], mq[1], mq[2], mq[3]);
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_b);
    if (warp == 2) tmem_alloc<512>(&misc->tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peers' barriers are initialised before anything is multicast to them
    tc_fence_after_sync();
    const uint32_t tmem_base = misc->tmem_base;

    if (warp < 4) {
        // Register budget: control warps give back what the workers take.  Two slots: 384 threads x 168 at launch,
        // 128 x 40 + 256 x 232 = 64512; three slots: 512 x 128 at launch, 128 x 40 + 384 x 152 = 63488 (<= 65536).
        // One slot (256 threads) launches with enough for everybody.
        if constexpr (NSLOTS >= 2) reg_dealloc<40>();
        if (warp == 0) {
            // ======================================================= TMA producer (codebook slices + norm slices)
            if (elect_one()) {
                uint32_t st = 0, ph = 0;
                const uint32_t part_bytes = B_SLICE_BYTES / (uint32_t)CL;
                const int part_rows = BCH / CL;
                const int nq_prep = (int)p.cb_meta[4];  // stages prepared: the norm slices follow their norms
                const uint8_t* nbase = reinterpret_cast<const uint8_t*>(p.cb_norm + (size_t)nq_prep * p.Kpad);
                for (JobIterN<NSLOTS> job(n_local, nq); job.valid(); job.next()) {
                    const int row0 = (p.q_begin + job.q) * p.Kpad + (int)crank * part_rows;
                    const uint8_t* nsrc = nbase + (size_t)(p.q_begin + job.q) * n_chunks * NSLICE_BYTES;
                    for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
                        for (int h = 0; h < ST_PER_CHUNK; ++h) {
                            const bool last = h == ST_PER_CHUNK - 1;
                            mbar_wait(&misc->empty[st], ph ^ 1);  // every CTA of the cluster has consumed the slot
                            mbar_arrive_expect_tx(&misc->full[st],
                                                  (uint32_t)SL_PER_ST * B_SLICE_BYTES + (last ? NSLICE_BYTES : 0u));
                            uint8_t* sbase = smem_b + (size_t)st * STAGE_BYTES;
#pragma unroll
                            for (int ks = 0; ks < SL_PER_ST; ++ks) {
                                uint8_t* dst = sbase + (uint32_t)ks * B_SLICE_BYTES + crank * part_bytes;
                                const int col = (h * SL_PER_ST + ks) * KSLICE;
                                if (CL > 1)  // my 1/CL of the slice goes to every CTA of the cluster
                                    tma_load_2d_mc(dst, &tmap_b, &misc->full[st], col, row0 + c * BCH, cmask);
                                else
                                    tma_load_2d(dst, &tmap_b, &misc->full[st], col, row0 + c * BCH);
                            }
                            if (last)
                                bulk_load_1d(sbase + (uint32_t)SL_PER_ST * B_SLICE_BYTES, nsrc + (size_t)c * NSLICE_BYTES,
                                             NSLICE_BYTES, &misc->full[st]);
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
            // The WHOLE warp walks the loop (uniform control flow: barrier phases, descriptors and addresses stay
            // in uniform registers); one elected lane issues the tcgen05 instructions.  A lone thread pays ~8-10
            // cycles per instruction, so everything per ring stage is unrolled onto compile-time descriptor offsets.
            const uint32_t idesc = umma_idesc_f16(0 /*fp16*/, TILE_M, CH);
            const uint32_t smem_b_u32 = smem_u32(smem_b);
            uint32_t st = 0, ph = 0, aphase = 0;
            uint32_t u0 = 0, u1 = 0, u2 = 0;  // per slot: uses of EACH accumulator buffer issued by finished jobs
            int prev_sl = -1;                  // slot of the previous job: the last reader of both buffers
#ifdef RVQ_FR_PROFILE
            long long m_ready = 0, m_empty = 0, m_full = 0, m_issue = 0;
            const long long m_t0 = clock64();
#define RVQ_MCLK(acc, stmt) { const long long t_ = clock64(); stmt; acc += clock64() - t_; }
#else
#define RVQ_MCLK(acc, stmt) { stmt; }
#endif
            for (JobIterN<NSLOTS> job(n_local, nq); job.valid(); job.next()) {
                const int sl = job.slot;
                RVQ_MCLK(m_ready, mbar_wait(&misc->a_ready[sl], (aphase >> sl) & 1));
                aphase ^= 1u << sl;
                tc_fence_after_sync();
                const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem + (size_t)sl * a_tile_bytes));
                const uint64_t adesc_x = umma_desc_nosw(smem_u32(misc->a_extra[sl]), 128, 256);
                const uint32_t ubase = sl == 0 ? u0 : (sl == 1 ? u1 : u2);
                const uint32_t uprev = prev_sl == 0 ? u0 : (prev_sl == 1 ? u1 : u2);
                for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
                    for (int h = 0; h < ST_PER_CHUNK; ++h) {
                        if (h == 0) {
                            // the last reader of each accumulator buffer this ring chunk writes has released it: my
                            // slot's previous use of it and (first use of the job) the previous job's last use
#pragma unroll
                            for (int sub = 0; sub < SUBS; ++sub) {
                                const uint32_t a = (uint32_t)(c * SUBS + sub);  // accumulator chunk of the job
                                const uint32_t buf = a & 1u;
                                RVQ_MCLK(m_empty, mbar_wait(&misc->tmem_empty[sl][buf], ((ubase + (a >> 1)) & 1u) ^ 1u));
                                if (NSLOTS > 1 && a < 2 && prev_sl >= 0 && prev_sl != sl)
                                    RVQ_MCLK(m_empty, mbar_wait(&misc->tmem_empty[prev_sl][buf], (uprev & 1u) ^ 1u));
                            }
                        }
                        RVQ_MCLK(m_full, mbar_wait(&misc->full[st], ph));
                        tc_fence_after_sync();
#ifdef RVQ_FR_PROFILE
                        const long long t_is = clock64();
#endif
                        if (elect_one()) {
                            const uint32_t sbase = smem_b_u32 + st * STAGE_BYTES;
                            const uint64_t bdesc = umma_desc_sw128(sbase);
#pragma unroll
                            for (int sub = 0; sub < SUBS; ++sub) {
                                const uint32_t buf = (uint32_t)(c * SUBS + sub) & 1u;
                                const uint32_t tmem_d = tmem_base + buf * CH;
                                // codes [sub * CH, (sub + 1) * CH) of the ring chunk: CH rows of 128 bytes further on
                                const int boff = sub * CH * 128 >> 4;
#pragma unroll
                                for (int ks = 0; ks < SL_PER_ST; ++ks) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)  // +32 bytes per K=16 step inside the swizzle row (>> 4)
                                        umma_f16_ss(tmem_d,
                                                    adesc0 + (uint64_t)((h * SL_PER_ST + ks) * (int)(A_SLICE_BYTES >> 4) + 2 * j),
                                                    bdesc + (uint64_t)(boff + ks * (int)(B_SLICE_BYTES >> 4) + 2 * j), idesc,
                                                    (h | ks | j) != 0);
                                }
                                if (h == ST_PER_CHUNK - 1) {
                                    // the norm term: + A_extra . B_extra^T (write_norm_slice, rvq_aux.cu); the slices
                                    // of codes [sub * CH, ...) start sub * CH / 8 core-matrix rows of 256 bytes in
                                    umma_f16_ss(tmem_d, adesc_x,
                                                umma_desc_nosw(sbase + (uint32_t)SL_PER_ST * B_SLICE_BYTES +
                                                                   (uint32_t)(sub * CH * 32),
                                                               128, 256),
                                                idesc, 1);
                                    if (sub + 1 < SUBS) umma_commit(&misc->tmem_full[sl][buf]);
                                }
                            }
                            // frees the ring slot (in every CTA of the cluster) when these MMAs retire
                            if (CL > 1)
                                umma_commit_mc(&misc->empty[st], cmask);
                            else
                                umma_commit(&misc->empty[st]);
                            if (h == ST_PER_CHUNK - 1)
                                umma_commit(&misc->tmem_full[sl][(uint32_t)(c * SUBS + SUBS - 1) & 1u]);
                        }
                        __syncwarp();
#ifdef RVQ_FR_PROFILE
                        m_issue += clock64() - t_is;
#endif
                        if (++st == (uint32_t)nstage) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
                if (sl == 0)
                    u0 += uses_per_job;
                else if (sl == 1)
                    u1 += uses_per_job;
                else
                    u2 += uses_per_job;
                prev_sl = sl;
            }
#ifdef RVQ_FR_PROFILE
            if (p.prof && lane == 0) {
                atomicAdd(p.prof + 10, (unsigned long long)(clock64() - m_t0));
                atomicAdd(p.prof + 11, (unsigned long long)m_ready);
                atomicAdd(p.prof + 12, (unsigned long long)m_empty);
                atomicAdd(p.prof + 13, (unsigned long long)m_full);
                atomicAdd(p.prof + 14, (unsigned long long)m_issue);
            }
#endif
        }
    } else {
        if constexpr (NSLOTS == 2) reg_alloc<232>();
        if constexpr (NSLOTS == 3) reg_alloc<152>();
        // =========================================================== worker groups (thread = frame)
        const int s = (warp - 4) >> 2;           // tile slot served by this group
        const int gw = warp & 3;                 // warp inside the group = TMEM lane quarter
        const int row = gw * 32 + lane;          // frame of the tile = TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(gw * 32) << 16);
        const uint32_t t_r = t_lane + TMEM_RES_COL + (uint32_t)(s * D);
        uint8_t* a_tile = smem + (size_t)s * a_tile_bytes;
        // A-tile address pieces of my frame: 16-byte chunk j of a 128-byte swizzle row sits at (j ^ (row & 7)) << 4
        uint8_t* a_row = a_tile + (uint32_t)row * 128u;
        const uint32_t rx = ((uint32_t)row & 7u) << 4;
        const bool row_major = (p.ad.sd == 1);
        // exposed residual rows of this warp live inside its OWN 32 operand rows (4 KiB per 64-feature slice)
        constexpr int RPP = 1024 / D;  // fp32 rows per 4 KiB piece
        auto scr_row = [&](int r) -> float* {
            return reinterpret_cast<float*>(a_tile + (uint32_t)(r / RPP) * A_SLICE_BYTES + (uint32_t)gw * 4096u +
                                            (uint32_t)(r % RPP) * (uint32_t)(D * 4));
        };
        // 16 consecutive features (c0 .. c0+15, inside one 64-feature slice) of my frame -> fp16 operand
        auto store_a16 = [&](int c0, const uint32_t (&v)[16], float sa) {
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
        };
        // operand row of the norm term for a frame whose operand exponents are a (row) and b (codes)
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
        // fifth operand column: rs / 64 rounded up, x (-64 xc_k) of the norm slice = the allowance rs * xc_k of the codes
        // above the stage's norm cap (k0_bound in rvq_aux.cu)
        auto store_a_rs = [&](float rs) {
            const __half2 h45 = __halves2half2(__float2half_ru(rs * 0.015625f), __float2half_rn(0.f));
            *reinterpret_cast<uint32_t*>(misc->a_extra[s] + (row >> 3) * 256 + (row & 7) * 16 + 8) =
                *reinterpret_cast<const uint32_t*>(&h45);
        };
        // the warp's residual rows -> its scratch rows: lanes with `mine` store their row at scratch row `rank`
        auto expose_rows = [&](bool mine, int rank) {
            float* dst = scr_row(mine ? rank : 0);
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_r + c0, v);
                tmem_ld_wait();
                if (mine) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            __syncwarp();
        };

        // per-frame state carried from the update of one stage to the scan of the next, in registers
        float delta = 0.f, amax_bound = 0.f, na_row = 1.f, rs_row = 0.f;
        const NormLayout nl(p.cb_norm, (int)p.cb_meta[4], p.Kpad);

        // load tile `tile` into this slot: residual <- x, fp16 operand + row constants of the first stage
        auto load_tile = [&](int tile) {
            const long long n = (long long)tile * TILE_M + row;
            const bool valid = n < p.N;
            const long long off = valid ? p.ad.row(n) : 0;
            float sq = 0.f, amax = 0.f;
            if (row_major) {
                // thread = frame: whole 32-byte sectors, two 32-feature pieces in flight
                uint32_t va[32], vb[32];
                const float* xr = p.x + off;
                if (valid) ldg_32f(xr, va);
#pragma unroll 1
                for (int c0 = 0; c0 < D; c0 += 64) {
                    if (valid) ldg_32f(xr + c0 + 32, vb);
                    if (!valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) va[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f = __uint_as_float(va[j]);
                        sq = fmaf(f, f, sq);
                        amax = fmaxf(amax, fabsf(f));
                    }
                    tmem_st_32x32(t_r + c0, va);
                    if (valid && c0 + 64 < D) ldg_32f(xr + c0 + 64, va);
                    if (!valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) vb[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f = __uint_as_float(vb[j]);
                        sq = fmaf(f, f, sq);
                        amax = fmaxf(amax, fabsf(f));
                    }
                    tmem_st_32x32(t_r + c0 + 32, vb);
                }
            } else {
                // frames-fastest storage (the reference's (B, d, L) tensor): lanes = consecutive frames
#pragma unroll 1
                for (int c0 = 0; c0 < D; c0 += 32) {
                    uint32_t v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = valid ? __float_as_uint(p.x[off + (long long)(c0 + j) * p.ad.sd]) : 0u;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f = __uint_as_float(v[j]);
                        sq = fmaf(f, f, sq);
                        amax = fmaxf(amax, fabsf(f));
                    }
                    tmem_st_32x32(t_r + c0, v);
                }
            }
            tmem_st_wait();
            const float4 m0 = misc->stage_meta[0];
            const float sb = m0.x, cnmax = m0.y;
            const int b = ilog2f_floor(sb);
            bool force_exact = !isfinite(sq);
            const int a = pick_row_exp(amax, b, force_exact);
            if (a - b < NORM_WINDOW_LO) force_exact = true;
            const float sa = exp2i(a);
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x16(t_r + c0, v);
                tmem_ld_wait();
                store_a16(c0, v, sa);
            }
            store_a_extra(a, b);
            row_consts(D, sq, force_exact, a, b, sb, cnmax, na_row, delta, rs_row);
            store_a_rs(rs_row);
            amax_bound = amax;
            fence_proxy_async_smem();
        };

        // ---------------- prologue: first tile of my slot
        if (s < n_local) {
            load_tile(blockIdx.x + s * gridDim.x);
            mbar_arrive(&misc->a_ready[s]);
        }
        uint32_t uses = 0;  // uses of each accumulator buffer by my slot's finished jobs
        unsigned n_rerank = 0, n_dirty_tot = 0, n_miss = 0;
#ifdef RVQ_FR_PROFILE  // per-phase cycle counters: a profiling build only (they cost ~16 registers per worker thread)
        unsigned n_jobs = 0;
        long long t_full = 0, t_scan = 0, t_rank = 0, t_apply = 0, t_tail = 0;
        const bool prof = p.prof != nullptr;
#define RVQ_CLK() (prof ? clock64() : 0)
#else
#define RVQ_CLK() 0
#endif
        for (JobIterN<NSLOTS> job(n_local, nq); job.valid(); job.next()) {
            if (job.slot != s) continue;
            [[maybe_unused]] const long long tp0 = RVQ_CLK();
            const int q = job.q, q_abs = p.q_begin + q;
            const int tile = blockIdx.x + job.i * gridDim.x;
            const long long n = (long long)tile * TILE_M + row;
            const bool valid = n < p.N;
            const int next_q_abs = (q + 1 < nq) ? q_abs + 1 : -1;
            const float* cbq = p.cb + (size_t)q_abs * p.K * D;
            const float4 mcur = misc->stage_meta[q];
            const int Kv = (int)mcur.w;
            // the next tile's frames start their way from HBM to L2 one stage early
            if (q + 1 == nq && job.i + NSLOTS < n_local && row_major) {
                const long long nn = (long long)(blockIdx.x + (job.i + NSLOTS) * gridDim.x) * TILE_M + row;
                if (nn < p.N) {
                    const float* xr = p.x + p.ad.row(nn);
#pragma unroll
                    for (int c0 = 0; c0 < D; c0 += 32) prefetch_l2(xr + c0);
                }
            }
            // ================================================== scan: two-dimensional running minimum of my frame
            float Cm[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) Cm[j] = BIG;
            float m1 = BIG, m2 = BIG, m3 = BIG, m4 = BIG;
            for (int c = 0; c < n_chunks * SUBS; ++c) {  // accumulator chunks of CH codes
                const uint32_t buf = (uint32_t)c & 1u;
                [[maybe_unused]] const long long tw0 = RVQ_CLK();
                mbar_wait(&misc->tmem_full[s][buf], (uses + ((uint32_t)c >> 1)) & 1u);
                tc_fence_after_sync();
#ifdef RVQ_FR_PROFILE
                if (prof) t_full += clock64() - tw0;
#endif
                const uint32_t taddr = t_lane + buf * CH;
                uint32_t va[16], vb[16];
                tmem_ld_32x16(taddr, va);
                uint32_t it = (uint32_t)c * (CH / 16);
#pragma unroll
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
                if (lane == 0) mbar_arrive(&misc->tmem_empty[s][buf]);
            }
            uses += uses_per_job;
            [[maybe_unused]] const long long tp1 = RVQ_CLK();
            // ================================================== classify my frame (registers only)
            float best = fminf(fminf(Cm[0], Cm[1]), Cm[2]);
#pragma unroll
            for (int j = 3; j < 15; j += 2) best = fminf(fminf(best, Cm[j]), Cm[j + 1]);
            best = fminf(best, Cm[15]);
            int jmin = 15;
#pragma unroll
            for (int j = 14; j >= 0; --j) jmin = (Cm[j] == best) ? j : jmin;
            // allowance of the code behind the best score (zero unless it is above the stage's norm cap, k0_bound)
            const float xbest = best_allowance(best, jmin, m1, m2, m3, m4,
                                               fmaf(rs_row, p.cb_meta[(size_t)q_abs * META_STRIDE + 7],
                                                    na_row * p.cb_meta[(size_t)q_abs * META_STRIDE + 5]),
                                               nl.xb + (size_t)q_abs * p.Kpad, p.Kpad - 1);
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
            const uint32_t rows_word = (__float_as_uint(m1) & IT_MASK) | ((__float_as_uint(m2) & IT_MASK) << 9) |
                                       ((__float_as_uint(m3) & IT_MASK) << 18) | (nr << 27);
            const int nc = (int)nr * __popc(cols);
            const bool dirty = over || nc == 0 || nc > 16;
            // the approximate argmin: the load of the smallest load minimum, the first column attaining the minimum
            int w = (int)((__float_as_uint(m1) & IT_MASK) * 16u) + jmin;
            w = max(0, min(w, Kv - 1));  // cannot bind (padding codes score 2^100); keeps the gather in bounds
            const int w_spec = w;
            // ================================================== exact scores where the filter left a choice (warp only)
            const unsigned fmask = __ballot_sync(FULL, !dirty && nc > 1);
            const unsigned dmask = __ballot_sync(FULL, dirty);
            if (fmask | dmask) {
                const int sub = lane & 7, grp = lane >> 3;
                unsigned pending = fmask;
                while (pending) {  // rounds of up to SCR_ROWS exposed rows
                    const int my_rank = __popc(pending & ((1u << lane) - 1u));
                    const bool mine = ((pending >> lane) & 1u) && my_rank < SCR_ROWS;
                    expose_rows(mine, my_rank);
                    const int n_round = min(SCR_ROWS, __popc(pending));
#pragma unroll 1
                    for (int base = 0; base < n_round; base += 4) {
                        const int r_idx = base + grp;  // exposed row scored by my 8-lane group
                        const bool g_on = r_idx < n_round;
                        // the lane that owns exposed row base + g, for each of the four groups
                        const unsigned o0 = __ballot_sync(FULL, mine && my_rank == base);
                        const unsigned o1 = __ballot_sync(FULL, mine && my_rank == base + 1);
                        const unsigned o2 = __ballot_sync(FULL, mine && my_rank == base + 2);
                        const unsigned o3 = __ballot_sync(FULL, mine && my_rank == base + 3);
                        const unsigned og = grp == 0 ? o0 : (grp == 1 ? o1 : (grp == 2 ? o2 : o3));
                        const int src = g_on ? ((__ffs(og) - 1) & 31) : 0;
                        const uint32_t rw = __shfl_sync(FULL, rows_word, src);
                        const uint32_t cm = __shfl_sync(FULL, cols, src);
                        const Cand1 cs(rw, cm);
                        const int ncs = g_on ? cs.n : 0;
                        int nmax = ncs;
                        nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 8));
                        nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 16));
                        const float* rrow = scr_row(g_on ? r_idx : 0);
                        float bs = __int_as_float(0x7f800000);
                        int bk = 0x7fffffff;
#pragma unroll 1
                        for (int e0 = 0; e0 < nmax; e0 += 4) {
                            int k[4];
                            const float* cc[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                k[t] = cs.code(e0 + t, Kv - 1);
                                cc[t] = cbq + (size_t)k[t] * D;
                            }
                            float sv[4];
                            exact_score8_n<4>(rrow, cc, D, sub, sv);
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if (e0 + t < ncs && better(sv[t], k[t], bs, bk)) {
                                    bs = sv[t];
                                    bk = k[t];
                                }
                        }
                        // owners of rows [base, base + 4) take their winner from the first lane of the scoring group
                        const bool take = mine && my_rank >= base && my_rank < base + 4;
                        const int got = __shfl_sync(FULL, bk, take ? (my_rank - base) * 8 : lane);
                        if (take) w = max(0, min(got, Kv - 1));
                    }
                    pending &= ~__ballot_sync(FULL, mine);
                    __syncwarp();  // the scratch rows are rewritten by the next round
                }
                // frames the filter could not bound: exact scan of the columns in reach, one frame at a time
                unsigned dm = dmask;
                while (dm) {
                    const int L = __ffs(dm) - 1;
                    dm &= dm - 1;
                    expose_rows(lane == L, 0);
                    uint32_t cm = __shfl_sync(FULL, cols, L);
                    if (cm == 0) cm = 0xFFFFu;
                    const ScoreIdx bsc = exact_scan_cols(scr_row(0), cbq, D, 0, (Kv + 15) / 16, cm, Kv, lane);
                    int rk = bsc.k;
                    if (rk < 0 || rk >= Kv) rk = 0;
                    if (lane == L) w = rk;
                    __syncwarp();
                }
                if (p.prof) {
                    n_rerank += __popc(fmask);
                    n_dirty_tot += __popc(dmask);
                    n_miss += __popc(__ballot_sync(FULL, valid && w != w_spec));
                }
            }
            [[maybe_unused]] const long long tp2 = RVQ_CLK();
            // ================================================== update: r <- r - c on my TMEM lane, next operand row
            // 32-feature pieces of the code vector in flight (three slots leave 152 registers per worker thread)
            constexpr int PF = D / 32 < (NSLOTS == 3 ? 2 : 3) ? D / 32 : (NSLOTS == 3 ? 2 : 3);
            const float* crow = cbq + (size_t)w * D;
            uint32_t cpre[PF][32];
#pragma unroll
            for (int i = 0; i < PF; ++i) ldg_nc_32f(crow + 32 * i, cpre[i]);
            // constants of the next stage's operand (scale chosen from a bound known now)
            const bool write_a = next_q_abs >= 0;
            float sb = 1.f, cnmax = 0.f, sa = 0.f;
            int a = 0, b = 0;
            bool force_exact = false;
            if (write_a) {
                const float4 mnext = misc->stage_meta[q + 1];
                sb = mnext.x;
                cnmax = mnext.y;
                b = ilog2f_floor(sb);
                a = pick_row_exp(amax_bound + mcur.z, b, force_exact);
                if (a - b < NORM_WINDOW_LO) force_exact = true;
                sa = exp2i(a);
                store_a_extra(a, b);
            }
            if (valid) {
                p.idx[n * nq + q] = w;
                if (kStats) atomicAdd(p.stats_cnt + (size_t)q_abs * p.K + w, 1.f);
            }
            float* ssum = (kStats && valid) ? p.stats_sum + ((size_t)q_abs * p.K + w) * D : nullptr;
            float sq = 0.f;
            {
                uint32_t va[16], vb[16];
                tmem_ld_32x16(t_r, va);
#pragma unroll
                for (int pc = 0; pc < D / 32; ++pc) {
                    const int c0 = pc * 32;
                    uint32_t(&cv)[32] = cpre[pc % PF];
                    tmem_ld_wait();
                    tmem_ld_32x16(t_r + c0 + 16, vb);
                    if (kStats && ssum) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            red_add_v4(ssum + c0 + j, make_float4(__uint_as_float(va[j]), __uint_as_float(va[j + 1]),
                                                                  __uint_as_float(va[j + 2]), __uint_as_float(va[j + 3])));
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float nf = __uint_as_float(va[j]) - __uint_as_float(cv[j]);
                        sq = fmaf(nf, nf, sq);
                        va[j] = __float_as_uint(nf);
                    }
                    tmem_st_32x16(t_r + c0, va);
                    if (write_a) store_a16(c0, va, sa);
                    tmem_ld_wait();
                    if (c0 + 32 < D) tmem_ld_32x16(t_r + c0 + 32, va);
                    if (kStats && ssum) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            red_add_v4(ssum + c0 + 16 + j,
                                       make_float4(__uint_as_float(vb[j]), __uint_as_float(vb[j + 1]),
                                                   __uint_as_float(vb[j + 2]), __uint_as_float(vb[j + 3])));
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float nf = __uint_as_float(vb[j]) - __uint_as_float(cv[16 + j]);
                        sq = fmaf(nf, nf, sq);
                        vb[j] = __float_as_uint(nf);
                    }
                    tmem_st_32x16(t_r + c0 + 16, vb);
                    if (write_a) store_a16(c0 + 16, vb, sa);
                    if (pc + PF < D / 32) ldg_nc_32f(crow + 32 * (pc + PF), cpre[pc % PF]);
                }
            }
            tmem_st_wait();
            amax_bound = sqrtf(sq) * 1.00002f;  // ||r'||_2 >= max|r'|
            if (write_a) {
                if (!isfinite(sq)) force_exact = true;
                row_consts(D, sq, force_exact, a, b, sb, cnmax, na_row, delta, rs_row);
                store_a_rs(rs_row);
            }
            {
                // commit-loss partial: sum over the valid frames of this warp
                double cs = valid ? (double)sq : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(FULL, cs, o);
                if (lane == 0 && cs != 0.0) atomicAdd(&misc->commit_acc[q], cs);
            }
            [[maybe_unused]] const long long tp3 = RVQ_CLK();
            if (write_a) {
                fence_proxy_async_smem();
                tc_fence_before_sync();
                mbar_arrive(&misc->a_ready[s]);
            } else {
                // ---------------- last stage: xq = x - final residual, then the slot takes its next tile
                const long long off = valid ? p.ad.row(n) : 0;
                if (row_major) {
                    const float* xr = p.x + off;
                    float* qr = p.xq + off;
                    uint32_t xa[32];
                    if (valid) ldg_32f(xr, xa);
#pragma unroll 1
                    for (int c0 = 0; c0 < D; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_r + c0, v);
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = __float_as_uint(__uint_as_float(xa[j]) - __uint_as_float(v[j]));
                            if (c0 + 32 < D) ldg_32f(xr + c0 + 32, xa);
                            stg_32f(qr + c0, v);
                        }
                    }
                } else {
#pragma unroll 1
                    for (int c0 = 0; c0 < D; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_r + c0, v);
                        tmem_ld_wait();
                        if (valid) {
                            // loads before stores: x and xq may alias as far as the compiler knows
                            float xv[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) xv[j] = __ldcs(p.x + off + (long long)(c0 + j) * p.ad.sd);
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                __stcs(p.xq + off + (long long)(c0 + j) * p.ad.sd, xv[j] - __uint_as_float(v[j]));
                        }
                    }
                }
                const int next_i = job.i + NSLOTS;
                if (next_i < n_local) {
                    load_tile(blockIdx.x + next_i * gridDim.x);
                    tc_fence_before_sync();
                    mbar_arrive(&misc->a_ready[s]);
                }
            }
#ifdef RVQ_FR_PROFILE
            if (prof) {
                const long long tp4 = clock64();
                t_scan += tp1 - tp0;
                t_rank += tp2 - tp1;
                t_apply += tp3 - tp2;
                t_tail += tp4 - tp3;
                ++n_jobs;
            }
#endif
        }
        if (p.prof && lane == 0) {
            atomicAdd(p.prof + 0, (unsigned long long)n_rerank);
            atomicAdd(p.prof + 1, (unsigned long long)n_dirty_tot);
            atomicAdd(p.prof + 2, (unsigned long long)n_miss);
#ifdef RVQ_FR_PROFILE
            if (gw == 0) {  // phase cycles of one warp per group, summed over its jobs
                atomicAdd(p.prof + 3, (unsigned long long)n_jobs);
                atomicAdd(p.prof + 4, (unsigned long long)t_scan);
                atomicAdd(p.prof + 5, (unsigned long long)t_full);
                atomicAdd(p.prof + 6, (unsigned long long)t_rank);
                atomicAdd(p.prof + 7, (unsigned long long)t_apply);
                atomicAdd(p.prof + 8, (unsigned long long)t_tail);
            }
#endif
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
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace fr
}  // namespace rvq

// ------------------------------------------------------------------------------------------ host side
using namespace rvq;

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled_fr() {
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

// The tensor map depends on the operand pointer and the shapes only (not on the codebook values): it is encoded
// once per (pointer, shape) and reused by every later launch of the calling thread.
struct TmapKey {
    const void* ptr;
    int d, rows, box_rows;
    bool operator==(const TmapKey& o) const { return ptr == o.ptr && d == o.d && rows == o.rows && box_rows == o.box_rows; }
};
struct TmapCache {
    TmapKey key{nullptr, 0, 0, 0};
    CUtensorMap map;
};

template <int D, bool kStats>
int launch_fr(const fr::Params& p0, const CUtensorMap& tmap, int num_sms, int smem_max, cudaStream_t st) {
    using C = fr::Cfg<D>;
    fr::Params p = p0;
    const uint32_t misc_bytes = (uint32_t)((sizeof(fr::Misc) + 1023) / 1024 * 1024);
    p.off_B = (uint32_t)C::NSLOTS * C::A_TILE_BYTES;
    const uint32_t fixed = p.off_B + misc_bytes + 1024;
    int ns = ((uint32_t)smem_max > fixed) ? (int)(((uint32_t)smem_max - fixed) / C::STAGE_BYTES) : 0;
    if (ns > fr::MAX_RING) ns = fr::MAX_RING;
    if (ns < 2) {
        set_error("rvq_encode: d=%d leaves no room for the codebook ring in %d bytes of shared memory", D, smem_max);
        return RVQ_ERR_ARG;
    }
    p.nstage = ns;
    p.off_misc = p.off_B + (uint32_t)ns * C::STAGE_BYTES;
    const uint32_t smem_total = p.off_misc + misc_bytes + 1024;
    auto kern = fr::rvq_encode_fr_kernel<D, kStats>;
    RVQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    const int CL = p.cluster;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(C::NUM_THREADS, 1, 1);
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
    const int want_clusters = (p.num_tiles + CL - 1) / CL;
    const int grid = (want_clusters < max_clusters ? want_clusters : max_clusters) * CL;
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    RVQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, p));
    RVQ_CUDA(cudaGetLastError());
    return RVQ_OK;
}
}  // namespace

bool rvq_fr_supported(int d) { return d == 64 || d == 128 || d == 256; }

int rvq_launch_fr(const float* x, long long N, long long L, long long sb, long long sl, long long sd, int d, int nq,
                  int K, int q_begin, const float* cb, const void* cb_op, int nq_total, const float* cb_norm,
                  const float* cb_meta, float* xq, long long* idx, double* commit_sq, float* stats_sum,
                  float* stats_cnt, int cluster, unsigned long long* prof, cudaStream_t st) {
    if (K > 32 * CHUNK_N) {
        set_error("rvq_encode: at most %d codes per stage are supported (got %d)", 32 * CHUNK_N, K);
        return RVQ_ERR_ARG;
    }
    if (nq > fr::MAX_NQ) {
        set_error("rvq_encode: at most %d stages are supported (got %d)", fr::MAX_NQ, nq);
        return RVQ_ERR_ARG;
    }
    int dev = 0, num_sms = 0, smem_max = 0;
    RVQ_CUDA(cudaGetDevice(&dev));
    RVQ_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    RVQ_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int Kpad = round_up(K, CHUNK_N);
    const int CL = (cluster == 1 || cluster == 2 || cluster == 4) ? cluster : 2;

    EncodeTiledFn encode = get_encode_tiled_fr();
    if (!encode) {
        set_error("rvq_encode: cuTensorMapEncodeTiled is not available from the driver");
        return RVQ_ERR_CUDA;
    }
    static thread_local TmapCache cache;
    const TmapKey key{cb_op, d, nq_total * Kpad, fr::BCH / CL};
    if (!(cache.key == key)) {
        const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)nq_total * Kpad};
        const cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
        const cuuint32_t box[2] = {(cuuint32_t)KSLICE, (cuuint32_t)(fr::BCH / CL)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode(&cache.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(cb_op), gdim,
                                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            cache.key = TmapKey{nullptr, 0, 0, 0};
            set_error("rvq_encode: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
            return RVQ_ERR_CUDA;
        }
        cache.key = key;
    }
    fr::Params p{};
    p.x = x;
    p.N = N;
    p.ad = RowAddrT{L, sb, sl, sd};
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
    p.num_tiles = (int)((N + TILE_M - 1) / TILE_M);
    p.cluster = CL;
    p.prof = prof;
    const bool stats = stats_sum != nullptr;
#define RVQ_FR_CASE(DD)                                                                              \
    case DD:                                                                                         \
        return stats ? launch_fr<DD, true>(p, cache.map, num_sms, smem_max, st)                      \
                     : launch_fr<DD, false>(p, cache.map, num_sms, smem_max, st)
    switch (d) {
        RVQ_FR_CASE(64);
        RVQ_FR_CASE(128);
        RVQ_FR_CASE(256);
    }
#undef RVQ_FR_CASE
    set_error("rvq_encode: the frame-resident kernel supports d = 64, 128, 256 (got %d)", d);
    return RVQ_ERR_ARG;
}
