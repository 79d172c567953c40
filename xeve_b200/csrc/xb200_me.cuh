// xb200_me.cuh -- motion search kernel (integer EPZS diamond + sub-pel pattern).
// Replaces reference src_base/xeve_pinter.c:122-140 (get_range_ipel), 363-551 (me_ipel_diamond),
// 272-361 (me_ipel_refinement), 553-697 (me_spel_pattern), 699-869 (pinter_me_epzs) over xeve_sad
// (src_base/xeve_sad.c:40-61) and xeve_mc_l (src_base/xeve_mc.c:99-254).
//
// Design (B200)
//  * One TEAM of threads per pi->fn_me call: a warp for 8x8 / 16x16 CUs (4 calls per CTA, warp
//    synchronous), 128 threads for 32x32, 256 for 64x64; the four size classes run as four
//    concurrent grids.
//  * The search window of the CU is staged once into shared memory by the TMA engine (one
//    cp.async.bulk per window row, completion on an mbarrier) and re-staged only when a refinement
//    pass leaves it.
//  * A diamond run of the reference is a sequence of rounds (dense window around the start, then
//    4-, 8-, 16-point diamonds of growing radius centred on the START point) whose candidate
//    POSITIONS do not depend on earlier rounds -- only their validity (the range is re-centred on
//    the round-1 winner) and the early exit do.  The kernel therefore evaluates the SAD + MV cost
//    of the whole candidate table of a run in one parallel batch (G lanes per candidate, packed
//    16-bit |a-b| with VIMNMX.U16x2, group-local shuffle reduction) and then REPLAYS the reference's
//    control flow on the cost table: per round a min over key = cost << 8 | order index restricted
//    to the candidates the reference would have evaluated (range test with the range as of that
//    round), strict-less update, not_found_best / beststep bookkeeping.  Over-evaluated candidates
//    never influence the result.
//  * Sub-pel candidates interpolate the whole CU from the staged window and accumulate the SAD in
//    the same pass.
#pragma once
#include "xb200_common.cuh"

#define ME_THREADS 128
#ifndef CU_PROF
#define CU_PROF(k) do { } while(0)
#endif
#define ME_MAX_CAND 160
#ifndef ME_ISSUE_SINGLE
#define ME_ISSUE_SINGLE 0
#endif
#ifndef ME_LAZY
#define ME_LAZY 1
#endif

template <int L2> struct MeGeom {
    static constexpr int W     = 1 << L2;
    static constexpr int QPR   = W / 4;                                  // 4-sample quads per row
    static constexpr int G     = L2 == 3 ? 1 : (L2 == 4 ? 4 : (L2 == 5 ? 8 : 32)); // lanes per candidate
    static constexpr int LPR   = QPR < G ? QPR : G;                      // lanes along a row
    static constexpr int RG    = G / LPR;                                // row groups inside the lane group
    static constexpr int QPL   = QPR / LPR;                              // quads per lane per row
    static constexpr int T     = L2 <= 4 ? 32 : (L2 == 5 ? 128 : XB200_T64);   // threads per call (team)
    static constexpr int CTA   = L2 <= 4 ? ME_THREADS : T;
    static constexpr int TEAMS = CTA / T;
    static constexpr int NG    = T / G;                                  // candidates in flight per team
};
template <int T> XB_DEV void me_team_sync() { team_bar<T>(); }

// shared memory of one team: mbarrier, cost table, reduction slots, original block, interpolation
// scratch, staged window (multiple of 16 bytes)
__host__ __device__ inline size_t me_team_bytes(int l2, int win_cap_elems)
{
    const int W = 1 << l2;
    return (16 + ME_MAX_CAND * 4 + 32 + (size_t)(W * W + (W + 8) * W + win_cap_elems + 16) * 2 + 15) & ~(size_t)15;
}
static inline size_t me_smem_bytes(int l2, int win_cap_elems) { return me_team_bytes(l2, win_cap_elems) * (l2 <= 4 ? 4 : 1); }

struct MeWin { // staged region [x0, x0+pitch) x [y0, y0+rows) in reference-plane coordinates
    int x0, y0, pitch, rows, staged, biased;
};

// SAD of one candidate by a group of G lanes; every lane of the group returns the total
template <int L2, bool ODD>
XB_DEV uint32_t me_group_sad_impl(const int16_t *__restrict__ win, int pitch, const int16_t *__restrict__ org, int ox, int oy, int j,
                                  unsigned gmask)
{
    using Gm = MeGeom<L2>;
    constexpr int ROWS  = Gm::W / Gm::RG;        // rows handled by one lane
    constexpr int QUADS = ROWS * Gm::QPL;        // 4-sample quads per lane
    constexpr int FLUSH = QUADS < 8 ? QUADS : 8; // packed halves hold <= 16 differences of <= 3069
    const int      col_lane = j % Gm::LPR, row_lane = j / Gm::LPR;
    const int16_t *base = win + (oy + row_lane) * pitch + (ox & ~1) + col_lane * 4;
    const int16_t *ob   = org + row_lane * Gm::W + col_lane * 4;
    uint32_t       total = 0;
#pragma unroll 1
    for(int q0 = 0; q0 < QUADS; q0 += FLUSH) {
        uint32_t acc = 0;
#pragma unroll
        for(int qq = 0; qq < FLUSH; qq++) {
            const int       q = q0 + qq, rr = (q / Gm::QPL) * Gm::RG, qc = (q % Gm::QPL) * Gm::LPR * 4;
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(base + rr * pitch + qc);
            const uint2     o  = *reinterpret_cast<const uint2 *>(ob + rr * Gm::W + qc);
            if(ODD) {
                const uint32_t w0 = rp[0], w1 = rp[1], w2 = rp[2];
                acc += absdiff_u16x2(o.x, __funnelshift_r(w0, w1, 16));
                acc += absdiff_u16x2(o.y, __funnelshift_r(w1, w2, 16));
            }
            else {
                acc += absdiff_u16x2(o.x, rp[0]);
                acc += absdiff_u16x2(o.y, rp[1]);
            }
        }
        total += sum_halves(acc);
    }
#pragma unroll
    for(int m = Gm::G >> 1; m > 0; m >>= 1) total += __shfl_xor_sync(gmask, total, m);
    return total;
}

// offsets of the reference's search patterns (src_base/xeve_pinter.c:57-65)
XB_DEV void me_d8(int i, int &dx, int &dy) // (-2,0) (-1,1) (0,2) (1,1) (2,0) (1,-1) (0,-2) (-1,-1) (0,0)
{
    dx = i == 8 ? 0 : (i < 4 ? i - 2 : 6 - i);
    dy = i == 8 ? 0 : (i < 2 ? i : (i < 6 ? 4 - i : i - 8));
}
XB_DEV void me_d16(int c, int &dx, int &dy) // (-4,0) (-3,1) .. (0,4) .. (4,0) .. (0,-4) .. (-3,-1)
{
    dx = c <= 8 ? c - 4 : 12 - c;
    dy = c <= 4 ? c : (c <= 12 ? 8 - c : c - 16);
}

// One pi->fn_me call by one team.  smem_raw: the team's me_team_bytes() area (its first 16 bytes are reserved: the batched kernels
// keep the mbarrier there); bar: the mbarrier the window copies complete on, initialised once (count 1) and never re-initialised --
// re-initialising a live mbarrier object is undefined (PTX mbarrier.init) and hangs the copy engine in practice; `phase` is its
// parity and persists across calls.  Results are uniform across the team's threads.
template <int L2>
XB_DEV void me_search(unsigned char *smem_raw, uint64_t *bar, const PicDev *__restrict__ pics, const xb200_me_item *__restrict__ it,
                      const int16_t *__restrict__ side, const SeqDev &sq, int win_cap_elems, int *__restrict__ err_flag, int tid,
                      uint32_t &phase, int &o_mv_x, int &o_mv_y, uint32_t &o_cost, int &o_mot_bits)
{
    using Gm = MeGeom<L2>;
    constexpr int W = Gm::W, T = Gm::T, G = Gm::G;
    uint32_t *costs = reinterpret_cast<uint32_t *>(smem_raw + 16);                       // cost table of the current run
    int32_t  *red   = reinterpret_cast<int32_t *>(smem_raw + 16 + ME_MAX_CAND * 4);      // 8 ints
    int16_t  *org   = reinterpret_cast<int16_t *>(smem_raw + 16 + ME_MAX_CAND * 4 + 32);
    int16_t  *tmp   = org + W * W;         // (W + 7) * W, horizontal pass of 2-D interpolation
    int16_t  *win   = tmp + (W + 8) * W;   // staged reference window (+ slack)
    const int      lane = tid & 31;
    const int      grp = tid / G, j = tid % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));

    const PicDev &cur = pics[it->cur_pic], &ref = pics[it->ref_pic];
    const int      x = it->x, y = it->y, bi = it->bi, lidx = it->lidx, bd = sq.bd;
    const int16_t *refy = ref.p[0];
    const int      sref = ref.s[0], ref_pad = ref.pad_l, ref_w = ref.w, ref_h = ref.h;

    // original block -> shared (picture rows, or the contiguous 2*org - pred block of the bi search).
    // The bi block holds signed values: both SAD operands are biased by ^0x8000 so that the packed
    // unsigned min/max sees them in the right order.
    {
        const uint32_t bias = bi ? 0x80008000u : 0u;
        const int16_t *src = bi ? side + it->org_bi_off : cur.p[0] + (size_t)y * cur.s[0] + x;
        const int      so  = bi ? W : cur.s[0];
        for(int e = tid; e < W * W / 4; e += T) {
            const int r = e / (W / 4), c = (e % (W / 4)) * 4;
            uint2 v = *reinterpret_cast<const uint2 *>(src + (size_t)r * so + c);
            v.x ^= bias; v.y ^= bias;
            *reinterpret_cast<uint2 *>(org + r * W + c) = v;
        }
    }
    me_team_sync<T>();

    // ---- search parameters (uniform) -----------------------------------------------------------
    const uint32_t lambda_mv = it->lambda_mv;
    const int      num_refp = it->num_refp, refi = it->refi;
    const int      other_bits = it->mot_bits_in[lidx ? 0 : 1];
    const int      static_range = it->max_search_range;
    int            dyn_range;
    {
        int d = it->poc - it->ref_poc;
        d     = d < 0 ? -d : d;
        dyn_range = clip3i(static_range >> 2, static_range, (static_range * d + (it->gop_size >> 1)) / it->gop_size);
    }
    const int gmvp_x = (int16_t)(it->mvp[0] + (x << 2)), gmvp_y = (int16_t)(it->mvp[1] + (y << 2));
    const int mvp_x = it->mvp[0], mvp_y = it->mvp[1];
    int       mot_bits_l = it->mot_bits_in[lidx];
    int       mv_x = it->mv_in[0], mv_y = it->mv_in[1];
    int       lo0, lo1, hi0, hi1; // current search range (get_range_ipel)

    MeWin wn;
    wn.staged = 0; wn.x0 = wn.y0 = wn.pitch = wn.rows = 0; wn.biased = 0;

    auto set_range = [&](int cx, int cy, int bi_mode) {
        const int r = bi_mode ? 5 : dyn_range;
        lo0 = clip3i(sq.min_clip[0], sq.max_clip[0], cx - r); hi0 = clip3i(sq.min_clip[0], sq.max_clip[0], cx + r);
        lo1 = clip3i(sq.min_clip[1], sq.max_clip[1], cy - r); hi1 = clip3i(sq.min_clip[1], sq.max_clip[1], cy + r);
    };
    auto mv_cost = [&](int qx, int qy, int &bits) -> uint32_t {
        bits = xb200_mv_bits(qx - gmvp_x, qy - gmvp_y, num_refp, refi);
        if(bi) bits += other_bits;
        return (uint32_t)((lambda_mv * (uint32_t)bits + (1u << 15)) >> 16);
    };
    // make sure [cx - m - 3, cx + m + W + 4) x [cy - m - 3, cy + m + W + 4) is resident in `win`
    auto ensure_window = [&](int cx, int cy, int m) {
        const int nx0 = max(cx - m - 3, -ref_pad), nx1 = min(cx + m + W + 4, ref_w + ref_pad);
        const int ny0 = max(cy - m - 3, -ref_pad), ny1 = min(cy + m + W + 4, ref_h + ref_pad);
        if(wn.staged && nx0 >= wn.x0 && nx1 <= wn.x0 + wn.pitch && ny0 >= wn.y0 && ny1 <= wn.y0 + wn.rows) return;
        const int ax0 = nx0 & ~7, ax1 = (nx1 + 7) & ~7; // -pad (144) is a multiple of 8
        int       pitch = ax1 - ax0, rows = ny1 - ny0;
        if(pitch * rows + 8 > win_cap_elems) { // cannot happen with host-side sizing; fail loudly
            if(tid == 0) atomicExch(err_flag, 1);
            rows = (win_cap_elems - 8) / pitch;
        }
        me_team_sync<T>(); // every reader of the previous window is done
#if ME_ISSUE_SINGLE
        if(tid == 0) { // one elected thread drives the copy engine: the row loop stays in uniform registers
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(bar, (uint32_t)(pitch * rows * 2));
            const int16_t *src = refy + (ptrdiff_t)ny0 * sref + ax0;
            int16_t       *dst = win;
#pragma unroll 4
            for(int r = 0; r < rows; r++, src += sref, dst += pitch) bulk_g2s(dst, src, (uint32_t)(pitch * 2), bar);
        }
#else
        if(tid < 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if(tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(pitch * rows * 2));
            __syncwarp();
            for(int r = lane; r < rows; r += 32)
                bulk_g2s(win + r * pitch, refy + (ptrdiff_t)(ny0 + r) * sref + ax0, (uint32_t)(pitch * 2), bar);
        }
#endif
        mbar_wait(bar, phase);
        phase ^= 1;
        { const int tt = tid; CU_PROF(22); }
        wn.x0 = ax0; wn.y0 = ny0; wn.pitch = pitch; wn.rows = rows; wn.staged = 1; wn.biased = 0;
    };
    auto set_bias = [&](int want) { // see the comment at the original-block load
        if(!bi || wn.biased == want) return;
        uint32_t *w32 = reinterpret_cast<uint32_t *>(win);
        for(int e = tid; e < (wn.pitch * wn.rows) >> 1; e += T) w32[e] ^= 0x80008000u;
        wn.biased = want;
        me_team_sync<T>();
    };

    // ---- candidate table of one diamond run around (c0x, c0y) --------------------------------------------
    // [0, NW): dense window (2R+1)^2, row-major; then step 4 (5 points), step 8 (9), steps 16, 32, ... (16 each)
    const int WR = bi == 1 ? 5 : 2, WN = 2 * WR + 1, NW = WN * WN;
    const int wrcp = (65536 + WN - 1) / WN; // exact i / WN for i < 128
    int       n_tab = NW;
    if(bi != 1) {
        n_tab += 5 + 9;
        for(int s = 16; s <= static_range; s <<= 1) n_tab += 16;
        if(static_range < 8) n_tab = NW + (static_range >= 4 ? 5 : 0);
        if(n_tab > ME_MAX_CAND) n_tab = ME_MAX_CAND;
    }
    auto cand_off = [&](int i, int &dx, int &dy) {
        if(i < NW) { const int q = (i * wrcp) >> 16; dx = i - q * WN - WR; dy = q - WR; return; }
        i -= NW;
        if(i < 5) { me_d8(i * 2, dx, dy); dx *= 2; dy *= 2; return; }
        i -= 5;
        if(i < 9) { me_d8(i, dx, dy); dx *= 4; dy *= 4; return; }
        i -= 9;
        me_d16(i & 15, dx, dy);
        const int m = 4 << (i >> 4);
        dx *= m; dy *= m;
    };
    // batch evaluation: costs[i] = MV cost + SAD for every table entry that lies inside the clip range and the
    // staged window (anything else can never be in range and stays UINT32_MAX)
    auto eval_table = [&](int c0x, int c0y, int from, int ntab) {
        for(int c = from + grp; c < ntab; c += Gm::NG) {
            int dx, dy;
            cand_off(c, dx, dy);
            const int  px = c0x + dx, py = c0y + dy, ox = px - wn.x0, oy = py - wn.y0;
            const bool ok = px >= sq.min_clip[0] && px <= sq.max_clip[0] && py >= sq.min_clip[1] && py <= sq.max_clip[1] && ox >= 0 &&
                            oy >= 0 && ox + W + 2 <= wn.pitch && oy + W <= wn.rows;
            uint32_t cost = 0xffffffffu;
            if(ok) {
                uint32_t sad = (ox & 1) ? me_group_sad_impl<L2, true>(win, wn.pitch, org, ox, oy, j, gmask)
                                        : me_group_sad_impl<L2, false>(win, wn.pitch, org, ox, oy, j, gmask);
                int bits;
                cost = mv_cost(px << 2, py << 2, bits);
                sad >>= (bd - 8);
                cost += bi ? (sad >> 1) : sad;
            }
            if(j == 0) costs[c] = cost;
        }
        me_team_sync<T>();
    };
    // min over table entries [a, b) of key = cost << 8 | (i - a), restricted to candidates inside the current
    // range (and, for the dense window, inside the extents the reference loops over); every warp computes it
    auto round_min = [&](int a, int b, int c0x, int c0y, int wx0, int wx1, int wy0, int wy1) -> uint64_t {
        uint64_t k = ~0ull;
        for(int i = a + lane; i < b; i += 32) {
            int dx, dy;
            cand_off(i, dx, dy);
            const int px = c0x + dx, py = c0y + dy;
            if(px < wx0 || px > wx1 || py < wy0 || py > wy1) continue; // not a candidate of this round
            uint32_t c = costs[i];
            if(px < lo0 || px > hi0 || py < lo1 || py > hi1) c = 0xffffffffu;
            k = min(k, ((uint64_t)c << 8) | (uint32_t)(i - a));
        }
#pragma unroll
        for(int m = 16; m > 0; m >>= 1) k = min(k, shfl_xor_u64(k, m));
        return k;
    };

    // one me_ipel_diamond run (start position already clipped by the caller's ensure_window)
    auto diamond = [&](int sx, int sy, int patience, int &bx, int &by, int &found_step, int &best_bits) -> uint32_t {
        const int c0x = clip3i(sq.min_clip[0], sq.max_clip[0], sx), c0y = clip3i(sq.min_clip[1], sq.max_clip[1], sy);
        // the table is evaluated lazily in three batches (window + step 4 | steps 8, 16 | the rest): most runs end
        // after the first or second batch (not_found_best early exit)
        int done = ME_LAZY ? min(n_tab, NW + 5) : n_tab;
        eval_table(c0x, c0y, 0, done);
        bx = c0x; by = c0y;
        uint32_t best_cost = 0xffffffffu;
        int      misses = 0, step = 0, tab = 0;
        best_bits = 0;
        for(;;) {
            misses++;
            uint64_t k;
            int      a, this_step;
            if(step <= 2) {
                const int wx0 = bx <= lo0 ? bx : bx - WR, wx1 = bx >= hi0 ? bx : bx + WR;
                const int wy0 = by <= lo1 ? by : by - WR, wy1 = by >= hi1 ? by : by + WR;
                a = 0;
                k = round_min(0, NW, c0x, c0y, wx0, wx1, wy0, wy1);
                this_step = 2;
                tab = NW;
            }
            else {
                const int cnt = step == 4 ? 5 : (step == 8 ? 9 : 16);
                a = tab;
                if(a + cnt > done && a + cnt <= n_tab) {
                    const int upto = (done < NW + 30) ? min(n_tab, NW + 30) : n_tab;
                    eval_table(c0x, c0y, done, upto);
                    done = upto;
                }
                k = (a + cnt <= n_tab) ? round_min(a, a + cnt, c0x, c0y, -32768, 32767, -32768, 32767) : ~0ull;
                this_step = step;
                tab += cnt;
            }
            const uint32_t kc = (uint32_t)(k >> 8);
            if(kc < best_cost) {
                int dx, dy;
                cand_off(a + (int)(k & 0xff), dx, dy);
                bx = c0x + dx; by = c0y + dy;
                best_cost = kc; found_step = this_step; misses = 0;
                int bits;
                mv_cost(bx << 2, by << 2, bits);
                best_bits = bits;
            }
            if(step <= 2) {
                set_range(bx, by, bi == 1);
                step += 2;
            }
            if(misses == patience || bi == 1) break;
            step <<= 1;
            if(step > static_range) break;
        }
        me_team_sync<T>(); // cost table free for the next run
        return best_cost;
    };

    // ---- pinter_me_epzs ----------------------------------------------------------------------------
    const int start_x = bi == 1 ? mv_x : mvp_x, start_y = bi == 1 ? mv_y : mvp_y;
    set_range(clip3i(sq.min_clip[0], sq.max_clip[0], x + (start_x >> 2)), clip3i(sq.min_clip[1], sq.max_clip[1], y + (start_y >> 2)),
              bi == 1);
    uint32_t best = 0xffffffffu, c;
    int      found = 0, beststep = 0, bits = 0, bx, by;
    {
        const int sx = ((int16_t)(start_x + (x << 2))) >> 2, sy = ((int16_t)(start_y + (y << 2))) >> 2;
        ensure_window(clip3i(sq.min_clip[0], sq.max_clip[0], sx), clip3i(sq.min_clip[1], sq.max_clip[1], sy),
                      bi == 1 ? 5 : dyn_range + 2);
        set_bias(1);
        c = diamond(sx, sy, 3, bx, by, found, bits);
    }
    if(bi != 1 && bits > 0) mot_bits_l = bits;
    if(c < best) {
        best = c; mv_x = (int16_t)((bx - x) << 2); mv_y = (int16_t)((by - y) << 2);
        beststep = (abs(mvp_x - mv_x) < 2 && abs(mvp_y - mv_y) < 2) ? 0 : found;
    }
    { const int tt = tid; CU_PROF(23); }
    while(bi != 1 && beststep > 0 && sq.me_complexity > 0) {
        set_range(x + (mv_x >> 2), y + (mv_y >> 2), 0);
        beststep = 0;
        const int sx = ((int16_t)(mv_x + (x << 2))) >> 2, sy = ((int16_t)(mv_y + (y << 2))) >> 2;
        ensure_window(clip3i(sq.min_clip[0], sq.max_clip[0], sx), clip3i(sq.min_clip[1], sq.max_clip[1], sy), dyn_range + 2);
        c = diamond(sx, sy, 2, bx, by, found, bits);
        if(bits > 0) mot_bits_l = bits;
        if(c < best) {
            best = c; mv_x = (int16_t)((bx - x) << 2); mv_y = (int16_t)((by - y) << 2);
            beststep = (abs(mvp_x - mv_x) < 2 && abs(mvp_y - mv_y) < 2) ? 0 : found;
        }
    }

    { const int tt = tid; CU_PROF(24); }
    if(sq.me_level > 1) {
        // ---- me_spel_pattern: every candidate = 8-tap interpolation of the whole CU + SAD -----------
        int           smv_x = mv_x, smv_y = mv_y, sbits = 0;
        const int16_t obias = bi ? (int16_t)0x8000 : (int16_t)0;
        uint32_t      sbest = 0xffffffffu;
        ensure_window(x + (mv_x >> 2), y + (mv_y >> 2), 2);
        set_bias(0);
        for(int stage = 0; stage < 2; stage++) {
            if(stage == 1 && sq.me_level <= 2) break;
            const int cx = (int16_t)(smv_x + (x << 2)), cy = (int16_t)(smv_y + (y << 2));
            const int cnt = stage ? sq.qpel_cnt : sq.hpel_cnt;
            for(int i = 0; i < cnt; i++) {
                int ox, oy;
                if(stage == 0) { // (-2,0) (-2,2) (0,2) (2,2) (2,0) (2,-2) (0,-2) (-2,-2)
                    ox = (i <= 1 || i == 7) ? -2 : ((i == 2 || i == 6) ? 0 : 2);
                    oy = (i == 0 || i == 4) ? 0 : ((i >= 1 && i <= 3) ? 2 : -2);
                }
                else {           // (-1,0) (0,1) (1,0) (0,-1) (-1,1) (1,1) (-1,-1) (1,-1)
                    ox = (i == 0 || i == 4 || i == 6) ? -1 : ((i == 1 || i == 3) ? 0 : 1);
                    oy = (i == 0 || i == 2) ? 0 : ((i == 1 || i == 4 || i == 5) ? 1 : -1);
                }
                const int qx = (int16_t)(cx + ox), qy = (int16_t)(cy + oy);
                const int dx = qx & 3, dy = qy & 3, ix = (qx >> 2) - wn.x0, iy = (qy >> 2) - wn.y0;
                const int maxv = (1 << bd) - 1;
                uint32_t  part = 0;
                if(dx && dy) {
                    const int s1 = min(4, bd - 8);
                    for(int e = tid; e < (W + 7) * W; e += T) {
                        const int r = e / W, cc = e % W;
                        const int16_t *p = win + (iy + r - 3) * wn.pitch + ix + cc - 3;
                        int acc = 0;
#pragma unroll
                        for(int t = 0; t < 8; t++) acc += c_mc_l[dx][t] * p[t];
                        tmp[e] = (int16_t)(acc >> s1);
                    }
                    me_team_sync<T>();
                    const int s2 = max(8, 20 - bd);
                    for(int e = tid; e < W * W; e += T) {
                        int acc = 0;
#pragma unroll
                        for(int t = 0; t < 8; t++) acc += c_mc_l[dy][t] * tmp[e + t * W];
                        const int v = clip3i(0, maxv, (acc + (1 << (s2 - 1))) >> s2);
                        part += (uint32_t)abs((int)(int16_t)(org[e] ^ obias) - v);
                    }
                }
                else {
                    const int stp = dx ? 1 : wn.pitch, ph = dx ? dx : dy;
                    for(int e = tid; e < W * W; e += T) {
                        const int r = e / W, cc = e % W;
                        int       v;
                        if(!dx && !dy) v = win[(iy + r) * wn.pitch + ix + cc];
                        else {
                            const int16_t *p = dx ? win + (iy + r) * wn.pitch + ix + cc - 3 : win + (iy + r - 3) * wn.pitch + ix + cc;
                            int acc = 0;
#pragma unroll
                            for(int t = 0; t < 8; t++) acc += c_mc_l[ph][t] * p[t * stp];
                            v = clip3i(0, maxv, acc >> 6);
                        }
                        part += (uint32_t)abs((int)(int16_t)(org[e] ^ obias) - v);
                    }
                }
#pragma unroll
                for(int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
                uint32_t sad = part;
                if(T > 32) {
                    team_bar<T>(); // red[] free, tmp[] reads done
                    if(lane == 0) red[tid >> 5] = (int32_t)part;
                    team_bar<T>();
                    sad = 0;
#pragma unroll
                    for(int wi = 0; wi < T / 32; wi++) sad += (uint32_t)red[wi];
                }
                else __syncwarp();
                sad >>= (bd - 8);
                int      cb;
                uint32_t cost = mv_cost(qx, qy, cb) + (bi ? (sad >> 1) : sad);
                if(cost < sbest) {
                    sbest = cost; smv_x = (int16_t)(qx - (x << 2)); smv_y = (int16_t)(qy - (y << 2));
                    if(stage) sbits = cb;
                }
            }
        }
        if(!bi && sbits > 0) mot_bits_l = sbits;
        if(sbest < best) { best = sbest; mv_x = smv_x; mv_y = smv_y; }
    }
    else {
        // me_ipel_refinement (src_base/xeve_pinter.c:272-361): 3x3 around the best integer position = the dense
        // window of a run with radius 1; evaluated through the same table (entries of the radius-2 window)
        set_range(x + (mv_x >> 2), y + (mv_y >> 2), bi == 1);
        const int ix = clip3i(sq.min_clip[0], sq.max_clip[0], ((int16_t)(mv_x + (x << 2))) >> 2);
        const int iy = clip3i(sq.min_clip[1], sq.max_clip[1], ((int16_t)(mv_y + (y << 2))) >> 2);
        ensure_window(ix, iy, WR + 1);
        set_bias(1);
        eval_table(ix, iy, 0, NW);
        // reference order: (0,0) (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)  [dx first]
        uint32_t rb = 0xffffffffu;
        int      rx = ix, ry = iy, rbits = 0;
        for(int cc = 0; cc < 9; cc++) {
            const int dx9 = cc == 0 ? 0 : (cc <= 3 ? -1 : (cc <= 5 ? 0 : 1));
            const int dy9 = cc == 0 ? 0 : (cc <= 3 ? cc - 2 : (cc == 4 ? -1 : (cc == 5 ? 1 : cc - 7)));
            const int px = ix + dx9, py = iy + dy9;
            uint32_t  cv = costs[(dy9 + WR) * WN + dx9 + WR];
            if(px < lo0 || px > hi0 || py < lo1 || py > hi1) cv = 0xffffffffu;
            if(cv < rb) { rb = cv; rx = px; ry = py; mv_cost(px << 2, py << 2, rbits); }
        }
        if(bi != 1 && rbits > 0) mot_bits_l = rbits;
        if(rb < best) { best = rb; mv_x = (int16_t)((rx - x) << 2); mv_y = (int16_t)((ry - y) << 2); }
    }

    { const int tt = tid; CU_PROF(25); }
    o_mv_x = mv_x; o_mv_y = mv_y; o_cost = best; o_mot_bits = mot_bits_l;
    me_team_sync<T>(); // the team's shared area is free again
}

template <int L2>
__global__ void __launch_bounds__(MeGeom<L2>::CTA) k_me(const PicDev *__restrict__ pics, xb200_me_item *__restrict__ items,
                                                        const int32_t *__restrict__ order, int n, const int16_t *__restrict__ side,
                                                        SeqDev sq, int win_cap_elems, int *__restrict__ err_flag)
{
    using Gm = MeGeom<L2>;
    constexpr int T = Gm::T;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int      team = threadIdx.x / T, tid = threadIdx.x % T; // tid = thread index inside the team
    unsigned char *smem_raw = smem_all + (size_t)team * me_team_bytes(L2, win_cap_elems);
    const int      item_no = blockIdx.x * Gm::TEAMS + team;
    if(item_no >= n) return; // whole team leaves (teams never meet at a block barrier when TEAMS > 1)
    xb200_me_item *it = &items[order[item_no]];
    if(tid == 0) mbar_init(reinterpret_cast<uint64_t *>(smem_raw), 1);
    uint32_t phase = 0, best;
    int      mv_x, mv_y, mot_bits_l;
    const int lidx = it->lidx, other_bits = it->mot_bits_in[lidx ? 0 : 1];
    me_search<L2>(smem_raw, reinterpret_cast<uint64_t *>(smem_raw), pics, it, side, sq, win_cap_elems, err_flag, tid, phase, mv_x, mv_y, best, mot_bits_l);
    if(tid == 0) {
        it->mv_out[0] = (int16_t)mv_x; it->mv_out[1] = (int16_t)mv_y; it->cost = best;
        it->mot_bits_out[lidx] = mot_bits_l;
        it->mot_bits_out[lidx ? 0 : 1] = other_bits;
    }
}
