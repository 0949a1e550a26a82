// tr_update_spec.cuh -- update group of rvq_encode_tr_kernel when no EMA statistics are requested:
// SPECULATE on the approximate argmin, VERIFY exactly in the shadow of the next stage's MMA / scan.
//
// Per tile slot the chain  scan -> exact re-rank of the uncertified frames -> r <- r - c -> next MMA  is serial,
// and tensor memory holds only two tiles to interleave.  Measured on C2: the exact winner differs from the
// approximate argmin for 0.08 of the 128 frames of a tile-stage (the fp16 filter's real error is ~30x below its
// proven bound), while the re-rank costs 7 k of the update's 17 k cycles.  So every frame is updated at once with
// the approximate argmin (the frames that will need exact scores also drop their stage-input residual row in
// shared memory on the way), the operand of the next stage is released to the MMA, and only then are the
// uncertified frames re-scored exactly.  A frame whose exact winner differs is repaired from its saved row
// (r' = r - c_exact, exactly what the non-speculative path computes); the scores the next stage produced for it
// came from a wrong operand row, so it is flagged and takes the exact scan there.  Results are bit-identical to
// the non-speculative path (tests/test_gpu_kernel_variants.py).
//
// A job falls back to the non-speculative order (re-rank first, then update) when it has a frame without a
// usable approximate argmin (repair pending, NaN / overflow / no fp16 window) or more uncertified frames than the
// row buffer holds.
//
// Included inside namespace rvq::tr by rvq_encode_tr.cu (uses its Params / Misc / constants).
#pragma once

constexpr int RS_CAP = 32;  // residual rows one tile slot can keep for verification

template <int kDummy = 0>
__device__ __forceinline__ void update_group_spec(const Params& p, Misc* misc, uint8_t* smem, float* rstage_all,
                                                  uint32_t tmem_base, int warp, int lane, int n_local) {
    const int d = p.d, nq = p.nq, nslots = p.nslots;
    const uint32_t a_tile_bytes = (uint32_t)(d / KSLICE) * A_SLICE_BYTES;
    const int s = (warp - UPD_WARP0) >> 2;   // tile slot served by this group
    const int row = (warp & 3) * 32 + lane;  // frame of the tile = TMEM lane (UPD_WARP0 % 4 == 0)
    const int gw = warp & 3;                 // warp inside the group
    if (s >= nslots) return;
    const uint32_t t_r = tmem_base + ((uint32_t)(gw * 32) << 16) + TMEM_RES_COL + (uint32_t)(s * d);
    uint8_t* a_tile = smem + (size_t)s * a_tile_bytes;
    uint8_t* a_row = a_tile + (uint32_t)row * 128u;
    const uint32_t rx = ((uint32_t)row & 7u) << 4;  // 16-byte chunk j of a swizzle row sits at (j ^ (row & 7)) << 4
    float* rstage = rstage_all + (size_t)s * RS_CAP * p.pitch;
    const bool row_major = (p.ad.sd == 1);
    const uint32_t bar_grp = BAR_GRP0 + (uint32_t)s;
    constexpr int NORM_WINDOW_LO = -10;  // below: 2^(a-b-4) leaves the fp16 normal range -> exact scan
    uint32_t sphase = 0;
    long long t_upd = 0, t_wait = 0, t_front = 0, t_apply = 0, t_verify = 0, t_tail = 0;
    unsigned long long n_dirty_tot = 0, n_multi_tot = 0, n_jobs = 0, n_legacy = 0, n_repair = 0;

    // 16 consecutive features (c0 .. c0+15) of my frame -> fp16 operand (scaled by sa) in the UMMA A tile
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
        v.z = v.w = 0u;
        *reinterpret_cast<uint4*>(misc->a_extra[s] + (row >> 3) * 256 + (row & 7) * 16) = v;
    };

    // load tile `tile` into this slot: residual <- x, fp16 operand + row constants of stage 0 (thread = frame,
    // 256-bit loads of its own 4d bytes, two 16-feature pieces ahead)
    auto load_tile = [&](int tile) {
        const long long n = (long long)tile * TILE_M + row;
        const bool valid = n < p.N;
        const long long off = valid ? p.ad.row(n) : 0;
        const float* xr = p.x + off;
        float sq = 0.f, amax = 0.f;
        auto take16 = [&](uint32_t (&v)[16], int c0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float f = __uint_as_float(v[j]);
                sq = fmaf(f, f, sq);
                amax = fmaxf(amax, fabsf(f));
            }
            tmem_st_32x16(t_r + c0, v);
        };
        if (row_major) {
            uint32_t xa[16], xb[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) xa[j] = xb[j] = 0u;
            if (valid) {
                ldg_nc_16f(xr, xa);
                ldg_nc_16f(xr + 16, xb);
            }
#pragma unroll 1
            for (int c0 = 0; c0 < d; c0 += 32) {
                take16(xa, c0);
                if (valid && c0 + 32 < d) ldg_nc_16f(xr + c0 + 32, xa);
                take16(xb, c0 + 16);
                if (valid && c0 + 32 < d) ldg_nc_16f(xr + c0 + 48, xb);
            }
        } else {
            // frames-fastest storage (the reference's (B, d, L) tensor): lanes = consecutive frames
#pragma unroll 1
            for (int c0 = 0; c0 < d; c0 += 16) {
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = valid ? __float_as_uint(xr[(long long)(c0 + j) * p.ad.sd]) : 0u;
                take16(v, c0);
            }
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
        for (int c0 = 0; c0 < d; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(t_r + c0, v);
            tmem_ld_wait();
            store_a16(c0, v, sa);
        }
        store_a_extra(a, b);
        float na, delta;
        row_consts(d, sq, force_exact, a, b, sb, cnmax, na, delta);
        misc->row_amax[s][row] = amax;
        misc->row_na[s][row] = na;
        misc->row_delta[s][row] = delta;
        misc->repair[s][row] = 0;
        fence_proxy_async_smem();
    };

    // ---------------- prologue: first tile of this slot
    if (s < n_local) {
        load_tile(blockIdx.x + s * gridDim.x);
        mbar_arrive(&misc->a_ready[s]);
    }
    for (JobIter job(n_local, nq, nslots); job.valid(); job.next()) {
        if (job.slot % nslots != s) continue;
        const int q = job.q, q_abs = p.q_begin + q;
        const int tile = blockIdx.x + job.i * gridDim.x;
        const long long n = (long long)tile * TILE_M + row;
        const bool valid = n < p.N;
        const int next_q_abs = (q + 1 < nq) ? q_abs + 1 : -1;
        const bool write_a = next_q_abs >= 0;
        const float* cbq = p.cb + (size_t)q_abs * p.K * d;
        const int Kv = (int)p.cb_meta[(size_t)q_abs * META_STRIDE + 3];
        const long long tj0 = clock64();
        mbar_wait(&misc->scan_done[s], sphase);
        sphase ^= 1u;
        const long long tj1 = clock64();
        // ---------------- classify my frame: certified (one candidate), several candidates (ceil(nc / 4) re-rank
        // entries of four candidates), exact scan (more than three loads in reach, no usable filter, repair)
        const uint32_t r0 = misc->g_rows[s][q & 1][0][row], r1 = misc->g_rows[s][q & 1][1][row];
        const uint32_t c0m = misc->g_cols[s][q & 1][0][row], c1m = misc->g_cols[s][q & 1][1][row];
        const int n0 = (int)((r0 >> 27) & 3u) * __popc(c0m), n1 = (int)((r1 >> 27) & 3u) * __popc(c1m);
        const int nc = n0 + n1;
        const bool repair_me = misc->repair[s][row] != 0;
        const bool hard = ((r0 | r1) & G_NOFILTER) != 0 || nc == 0 || repair_me;  // no usable approximate argmin
        const bool dirty = hard || ((r0 | r1) & G_OVER) != 0 || nc > 16;
        const int w_spec = max(0, min(Kv - 1, misc->vbest[s][0][row] <= misc->vbest[s][1][row]
                                                  ? (int)misc->wbest[s][0][row]
                                                  : (int)misc->wbest[s][1][row]));
        int w = w_spec, mypos = -1, rpos = -1;
        const int myk = dirty ? 1 : (nc == 1 ? 0 : (nc + 3) >> 2);
        if (myk > 0) {
            mypos = atomicAdd(&misc->n_special[s], myk);
            for (int i = 0; i < myk; ++i)
                misc->special_rows[s][mypos + i] = (uint16_t)(row | (i << 8) | (dirty ? 0x8000 : 0));
            if (dirty) atomicAdd(&misc->n_dirty[s], 1);
            rpos = atomicAdd(&misc->n_srows[s], 1);
            misc->pos_of[s][row] = (uint8_t)min(rpos, 255);
            if (hard) {
                atomicAdd(&misc->n_hard[s], 1);
                if (p.prof) atomicAdd(p.prof + (((r0 | r1) & G_NOFILTER) ? 23 : (nc == 0 ? 24 : 25)), 1ull);
            }
        }
        named_bar_sync(bar_grp, GRP_THREADS);
        const int n_special = misc->n_special[s], n_dirty = misc->n_dirty[s];
        const int n_srows = misc->n_srows[s];
        const bool legacy = misc->n_hard[s] > 0 || n_srows > RS_CAP;  // uniform over the group
        float bs = __int_as_float(0x7f800000);
        int kwin = 0x7fffffff;
        // ---------------- exact scores of the uncertified frames, RS_ROWS entries per round.  expose_now: the rows
        // are read from tensor memory into the row buffer round by round (non-speculative order); otherwise they
        // were dropped there (one row per frame, at pos_of) while the frame was updated.
        auto rerank = [&](bool expose_now) {
            const uint32_t cols_all = 0xFFFFu;
#pragma unroll 1
            for (int base = 0; base < n_special; base += RS_ROWS) {
                const int lo = max(mypos, base), hi = min(mypos + myk, base + RS_ROWS);
                const bool in_round = mypos >= 0 && lo < hi;
                if (expose_now) {
                    if (__any_sync(0xffffffffu, in_round)) {
                        float* dst = rstage + (size_t)(in_round ? lo - base : 0) * p.pitch;
#pragma unroll 1
                        for (int c0 = 0; c0 < d; c0 += 16) {
                            uint32_t v[16];
                            tmem_ld_32x16(t_r + c0, v);
                            tmem_ld_wait();
                            if (in_round) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4)
                                    *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            }
                        }
                    }
                    named_bar_sync(bar_grp, GRP_THREADS);  // rows of this round exposed
                }
                const int n_round = min(RS_ROWS, n_special - base);
                {
                    const int sub = lane & 7, slot16 = (gw * 32 + lane) >> 3;
                    const int ent = slot16 < n_round ? misc->special_rows[s][base + slot16] : 0x8000;
                    if (__any_sync(0xffffffffu, !(ent & 0x8000))) {
                        const int rr = ent & 0x7f, blk = (ent >> 8) & 0x7f;
                        const CandSet cs(misc->g_rows[s][q & 1][0][rr], misc->g_cols[s][q & 1][0][rr], misc->g_rows[s][q & 1][1][rr],
                                         misc->g_cols[s][q & 1][1][rr]);
                        const int ri = expose_now ? max(slot16 - blk, 0) : (int)misc->pos_of[s][rr];
                        const float* rrow = rstage + (size_t)min(ri, RS_CAP - 1) * p.pitch;
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
                // frames the filter could not bound: exact scan of the columns in reach, all four warps
                if (n_dirty > 0) {
#pragma unroll 1
                    for (int i = 0; i < n_round; ++i) {
                        const int ent = misc->special_rows[s][base + i];
                        if (!(ent & 0x8000)) continue;
                        const int rr = ent & 0x7f;
                        const uint32_t gr = misc->g_rows[s][q & 1][0][rr] | misc->g_rows[s][q & 1][1][rr];
                        uint32_t cols = (uint32_t)misc->g_cols[s][q & 1][0][rr] | (uint32_t)misc->g_cols[s][q & 1][1][rr];
                        if (cols == 0 || (gr & G_NOFILTER) || misc->repair[s][rr]) cols = cols_all;
                        const int ri = expose_now ? i : (int)misc->pos_of[s][rr];
                        const int n_it = (Kv + 15) / 16, per_w = (n_it + 3) / 4;
                        const int it0 = min(n_it, gw * per_w), it1 = min(n_it, it0 + per_w);
                        const ScoreIdx bsc =
                            exact_scan_cols(rstage + (size_t)min(ri, RS_CAP - 1) * p.pitch, cbq, d, it0, it1, cols, Kv, lane);
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
                if (base + RS_ROWS < n_special) named_bar_sync(bar_grp, GRP_THREADS);  // round buffers are reused
            }
        };
        if (legacy) {
            if (p.prof && gw == 0 && lane == 0 && n_srows > RS_CAP) atomicAdd(p.prof + 22, 1ull);
            rerank(true);
            if (myk > 0) w = dirty ? misc->win[s][row] : min(kwin, Kv - 1);
            ++n_legacy;
        }
        const long long tj2 = clock64();
        // ---------------- r <- r - c with the (approximate, or in the fallback order exact) winner: the code
        // vector comes straight into registers with 256-bit loads, two 16-feature pieces ahead of their use
        float sa = 0.f, sb = 1.f, cnmax = 0.f;
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
        // a frame to be verified keeps its stage-input residual row in shared memory
        const bool keep = !legacy && rpos >= 0;
        float* keep_row = rstage + (size_t)(keep ? rpos : 0) * p.pitch;
        float sq = 0.f;
        {
            const float* crow = cbq + (size_t)w * d;
            uint32_t ca[16], cb_[16], va[16], vb[16];
            ldg_nc_16f(crow, ca);
            ldg_nc_16f(crow + 16, cb_);
            auto apply16 = [&](uint32_t (&v)[16], const uint32_t (&c)[16], int c0) {
                if (keep) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<uint4*>(keep_row + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float nf = __uint_as_float(v[j]) - __uint_as_float(c[j]);
                    sq = fmaf(nf, nf, sq);
                    v[j] = __float_as_uint(nf);
                }
                tmem_st_32x16(t_r + c0, v);
                if (write_a) store_a16(c0, v, sa);
            };
            tmem_ld_32x16(t_r, va);
#pragma unroll 1
            for (int c0 = 0; c0 < d; c0 += 32) {
                tmem_ld_wait();
                tmem_ld_32x16(t_r + c0 + 16, vb);
                apply16(va, ca, c0);
                if (c0 + 32 < d) ldg_nc_16f(crow + c0 + 32, ca);
                tmem_ld_wait();
                if (c0 + 32 < d) tmem_ld_32x16(t_r + c0 + 32, va);
                apply16(vb, cb_, c0 + 16);
                if (c0 + 32 < d) ldg_nc_16f(crow + c0 + 48, cb_);
            }
        }
        tmem_st_wait();
        misc->row_amax[s][row] = sqrtf(sq) * 1.00002f;  // ||r'||_2 >= max|r'|
        if (write_a) {
            if (!isfinite(sq)) force_exact = true;
            float na, delta;
            row_consts(d, sq, force_exact, a, b, sb, cnmax, na, delta);
            misc->row_na[s][row] = na;
            misc->row_delta[s][row] = delta;
            fence_proxy_async_smem();
        }
        if (repair_me) misc->repair[s][row] = 0;  // resolved by the exact scan above (fallback order)
        if (write_a) mbar_arrive(&misc->a_ready[s]);  // the next stage's MMA may start
        const long long tj3 = clock64();
        // ---------------- verification (speculative order): exact scores in the shadow of the next stage
        if (!legacy && n_srows > 0) {
            named_bar_sync(bar_grp, GRP_THREADS);  // every kept row is complete
            rerank(false);
            bool miss = false;
            if (myk > 0) {
                const int w_exact = dirty ? misc->win[s][row] : min(kwin, Kv - 1);
                miss = w_exact != w;
                w = w_exact;
            }
            if (__any_sync(0xffffffffu, miss)) {
                // repair: r' = (saved stage input) - c_exact, exactly as the non-speculative order computes it; the
                // other lanes of the warp pass their residual through (tcgen05.st is warp-wide)
                const float* crow = cbq + (size_t)w * d;
                float sq2 = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < d; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_r + c0, v);
                    tmem_ld_wait();
                    if (miss) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 rv = *reinterpret_cast<const float4*>(keep_row + c0 + j);
                            const float4 cv = ldg_nc_v4(crow + c0 + j);
                            const float n0f = rv.x - cv.x, n1f = rv.y - cv.y, n2f = rv.z - cv.z, n3f = rv.w - cv.w;
                            sq2 = fmaf(n0f, n0f, sq2);
                            sq2 = fmaf(n1f, n1f, sq2);
                            sq2 = fmaf(n2f, n2f, sq2);
                            sq2 = fmaf(n3f, n3f, sq2);
                            v[j + 0] = __float_as_uint(n0f);
                            v[j + 1] = __float_as_uint(n1f);
                            v[j + 2] = __float_as_uint(n2f);
                            v[j + 3] = __float_as_uint(n3f);
                        }
                    }
                    tmem_st_32x16(t_r + c0, v);
                }
                tmem_st_wait();
                if (miss) {
                    sq = sq2;
                    misc->row_amax[s][row] = sqrtf(sq) * 1.00002f;
                    // the next stage scored this frame with a wrong operand row: it takes the exact scan there
                    if (write_a) misc->repair[s][row] = 1;
                    ++n_repair;
                }
            }
        }
        // ---------------- bookkeeping: code index, commit-loss partial; counters of the next job of this slot
        if (valid) p.idx[n * nq + q] = w;
        {
            double cs = valid ? (double)sq : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
            if (lane == 0 && cs != 0.0) atomicAdd(&misc->commit_acc[q], cs);
        }
        named_bar_sync(bar_grp, GRP_THREADS);  // every thread has read the lists and counters of this job
        if (gw == 0 && lane == 0) {
            misc->n_special[s] = 0;
            misc->n_dirty[s] = 0;
            misc->n_srows[s] = 0;
            misc->n_hard[s] = 0;
        }
        const long long tj4 = clock64();
        if (!write_a) {
            // ---------------- last stage: xq = x - final residual, then the slot takes its next tile
            const long long off = valid ? p.ad.row(n) : 0;
            if (row_major) {
                // every lane runs the warp-wide TMEM loads; global accesses are predicated on the frame being real
                const float* xr = p.x + off;
                float* qr = p.xq + off;
                uint32_t xa[16], xb[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) xa[j] = xb[j] = 0u;
                if (valid) {
                    ldg_nc_16f(xr, xa);
                    ldg_nc_16f(xr + 16, xb);
                }
                auto out16 = [&](uint32_t (&xv)[16], const uint32_t (&v)[16], int c0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) xv[j] = __float_as_uint(__uint_as_float(xv[j]) - __uint_as_float(v[j]));
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; j += 8)
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(qr + c0 + j),
                                         "r"(xv[j]), "r"(xv[j + 1]), "r"(xv[j + 2]), "r"(xv[j + 3]), "r"(xv[j + 4]),
                                         "r"(xv[j + 5]), "r"(xv[j + 6]), "r"(xv[j + 7])
                                         : "memory");
                    }
                };
#pragma unroll 1
                for (int c0 = 0; c0 < d; c0 += 32) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_r + c0, v);
                    tmem_ld_wait();
                    out16(xa, v, c0);
                    if (valid && c0 + 32 < d) ldg_nc_16f(xr + c0 + 32, xa);
                    tmem_ld_32x16(t_r + c0 + 16, v);
                    tmem_ld_wait();
                    out16(xb, v, c0 + 16);
                    if (valid && c0 + 32 < d) ldg_nc_16f(xr + c0 + 48, xb);
                }
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < d; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_r + c0, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const long long o = off + (long long)(c0 + j) * p.ad.sd;
                            p.xq[o] = p.x[o] - __uint_as_float(v[j]);
                        }
                    }
                }
            }
            const int next_i = job.i + nslots;
            if (next_i < n_local) {
                load_tile(blockIdx.x + next_i * gridDim.x);
                mbar_arrive(&misc->a_ready[s]);
            }
        }
        const long long tj5 = clock64();
        t_wait += tj1 - tj0;
        t_front += tj2 - tj1;
        t_apply += tj3 - tj2;
        t_verify += tj4 - tj3;
        t_tail += tj5 - tj4;
        t_upd += tj5 - tj1;
        n_multi_tot += n_special;
        ++n_jobs;
    }
    if (p.prof && gw == 0 && lane == 0) {
        atomicAdd(p.prof + 2, (unsigned long long)t_upd);
        atomicAdd(p.prof + 4, n_dirty_tot);
        atomicAdd(p.prof + 5, n_jobs);
        atomicAdd(p.prof + 6, n_multi_tot);
        atomicAdd(p.prof + 7, (unsigned long long)t_wait);
        atomicAdd(p.prof + 8, (unsigned long long)t_front);
        atomicAdd(p.prof + 9, (unsigned long long)t_apply);
        atomicAdd(p.prof + 10, (unsigned long long)t_verify);
        atomicAdd(p.prof + 12, (unsigned long long)t_tail);
        atomicAdd(p.prof + 21, n_legacy);
    }
    if (p.prof && lane == 0 && n_repair) atomicAdd(p.prof + 20, n_repair);
}
