// xb200_me.cuh -- motion search kernel (integer EPZS diamond + sub-pel pattern), one CTA per
// pi->fn_me call.  Replaces reference src_base/xeve_pinter.c:122-140 (get_range_ipel),
// 363-551 (me_ipel_diamond), 553-697 (me_spel_pattern), 699-869 (pinter_me_epzs) over
// xeve_sad (src_base/xeve_sad.c:40-61) and xeve_mc_l (src_base/xeve_mc.c:99-254).
//
// Design (B200): the reference window of the CU is staged ONCE into shared memory by the TMA
// engine (one cp.async.bulk per row, completion on an mbarrier); every search round evaluates
// its whole candidate set in parallel -- a group of G lanes per candidate, packed s16x2
// |a-b| (VIMNMX.S16x2), warp-shuffle reduction -- and then every warp redundantly reduces the
// round's (cost, order) keys so the data-dependent control (re-centring, early exits, refinement
// passes) runs uniformly in all threads with one block barrier per round.  Tie-breaking follows
// the reference's evaluation order: the key is cost << 32 | candidate index.
#pragma once
#include "xb200_common.cuh"

#define ME_THREADS 128

// shared memory of one team: mbarrier, round keys, reduction slots, original block, interpolation
// scratch, staged window (multiple of 16 bytes)
__host__ __device__ inline size_t me_team_bytes(int l2, int win_cap_elems)
{
    const int W = 1 << l2;
    return (16 + 2 * 128 * 8 + 32 + (size_t)(W * W + (W + 8) * W + win_cap_elems + 16) * 2 + 15) & ~(size_t)15;
}
static inline size_t me_smem_bytes(int l2, int win_cap_elems) { return me_team_bytes(l2, win_cap_elems) * (l2 <= 4 ? 4 : 1); }

struct MeState { // uniform per CTA (kept in registers by every thread)
    int lo[2], hi[2];
};

template <int L2> struct MeGeom {
    static constexpr int W    = 1 << L2;
    static constexpr int QPR  = W / 4;                                  // 4-sample quads per row
    static constexpr int G    = L2 == 3 ? 4 : (L2 == 4 ? 16 : 32);      // lanes per candidate
    static constexpr int LPR  = QPR < G ? QPR : G;                      // lanes along a row
    static constexpr int RG   = G / LPR;                                // row groups inside the lane group
    static constexpr int QPL  = QPR / LPR;                              // quads per lane per row
    static constexpr int T    = L2 <= 4 ? 32 : (L2 == 5 ? 128 : 256);   // threads per item (team)
    static constexpr int CTA   = L2 <= 4 ? ME_THREADS : T;
    static constexpr int TEAMS = CTA / T;
    static constexpr int NG   = T / G;                                  // candidates in flight per team
};
template <int T> XB_DEV void me_team_sync()
{
    if(T == 32) __syncwarp();
    else __syncthreads();
}

// window bookkeeping: staged region [x0, x0+pitch) x [y0, y0+rows) in reference-plane coordinates
struct MeWin {
    int x0, y0, pitch, rows, staged, biased;
};

template <int L2, bool ODD>
XB_DEV uint32_t me_group_sad_impl(const int16_t *__restrict__ win, int pitch, const int16_t *__restrict__ org, int ox, int oy, int j)
{
    using Gm = MeGeom<L2>;
    constexpr int ROWS = Gm::W / Gm::RG;          // rows handled by one lane
    constexpr int QUADS = ROWS * Gm::QPL;         // 4-sample quads per lane: 4 / 4 / 8 / 32
    constexpr int FLUSH = QUADS < 8 ? QUADS : 8;  // packed halves hold <= 16 differences of <= 3069
    const int     col_lane = j % Gm::LPR, row_lane = j / Gm::LPR;
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
    for(int m = Gm::G >> 1; m > 0; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
    return total;
}
template <int L2>
XB_DEV uint32_t me_group_sad(const int16_t *__restrict__ win, const MeWin &wn, const int16_t *__restrict__ org, int px, int py, int j)
{
    const int ox = px - wn.x0, oy = py - wn.y0;
    return (ox & 1) ? me_group_sad_impl<L2, true>(win, wn.pitch, org, ox, oy, j) : me_group_sad_impl<L2, false>(win, wn.pitch, org, ox, oy, j);
}

template <int L2>
__global__ void __launch_bounds__(MeGeom<L2>::CTA) k_me(const PicDev *__restrict__ pics, xb200_me_item *__restrict__ items,
                                                    const int32_t *__restrict__ order, int n, const int16_t *__restrict__ side,
                                                    SeqDev sq, int win_cap_elems, int *__restrict__ err_flag)
{
    using Gm = MeGeom<L2>;
    constexpr int W = Gm::W;
    constexpr int T = Gm::T;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int team = threadIdx.x / T;
    unsigned char *smem_raw = smem_all + (size_t)team * me_team_bytes(L2, win_cap_elems);
    uint64_t *bar   = reinterpret_cast<uint64_t *>(smem_raw);                 // 8 B (padded to 16)
    uint64_t *keys  = reinterpret_cast<uint64_t *>(smem_raw + 16);            // 2 x 128 keys
    int32_t  *red   = reinterpret_cast<int32_t *>(smem_raw + 16 + 2 * 128 * 8); // 8 ints
    int16_t  *org   = reinterpret_cast<int16_t *>(smem_raw + 16 + 2 * 128 * 8 + 32);
    int16_t  *tmp   = org + W * W;                  // (W + 7) * W, horizontal pass of 2-D interpolation
    int16_t  *win   = tmp + (W + 8) * W;            // staged reference window (+ slack)

    const int tid = threadIdx.x % T, lane = tid & 31; // tid = thread index inside the team
    const int grp = tid / Gm::G, j = tid % Gm::G;
    const int item_no = blockIdx.x * Gm::TEAMS + team;
    if(item_no >= n) return; // whole team leaves (teams never meet at a block barrier when TEAMS > 1)
    const int it_idx = order[item_no];
    xb200_me_item *it = &items[it_idx];

    if(tid == 0) mbar_init(bar, 1);
    uint32_t phase = 0;

    const PicDev cur = pics[it->cur_pic], ref = pics[it->ref_pic];
    const int    x = it->x, y = it->y, bi = it->bi, lidx = it->lidx;
    const int    bd = sq.bd;
    const int16_t *refy = ref.p[0];
    const int      sref = ref.s[0];

    // original block -> shared (picture rows, or the contiguous 2*org - pred block of bi search).
    // The bi block holds signed values: both operands are biased by ^0x8000 so that the packed
    // unsigned max/min of the SAD sees them in the right order.
    const uint32_t bias = bi ? 0x80008000u : 0u;
    {
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

    MeState st;
    MeWin   wn;
    wn.staged = 0; wn.x0 = wn.y0 = wn.pitch = wn.rows = 0; wn.biased = 0;
    int key_buf = 0;

    auto set_window = [&](int cx, int cy, int bi_mode) {
        const int r = bi_mode ? 5 : dyn_range;
        st.lo[0] = clip3i(sq.min_clip[0], sq.max_clip[0], cx - r);
        st.hi[0] = clip3i(sq.min_clip[0], sq.max_clip[0], cx + r);
        st.lo[1] = clip3i(sq.min_clip[1], sq.max_clip[1], cy - r);
        st.hi[1] = clip3i(sq.min_clip[1], sq.max_clip[1], cy + r);
    };
    auto mv_cost = [&](int qx, int qy, int &bits) -> uint32_t {
        bits = xb200_mv_bits(qx - gmvp_x, qy - gmvp_y, num_refp, refi);
        if(bi) bits += other_bits;
        return (uint32_t)((lambda_mv * (uint32_t)bits + (1u << 15)) >> 16);
    };
    // make sure [cx - m - 3, cx + m + W + 4) x [cy - m - 3, cy + m + W + 4) is resident in `win`
    auto ensure_window = [&](int cx, int cy, int m) {
        int nx0 = max(cx - m - 3, -ref.pad_l), nx1 = min(cx + m + W + 4, ref.w + ref.pad_l);
        int ny0 = max(cy - m - 3, -ref.pad_l), ny1 = min(cy + m + W + 4, ref.h + ref.pad_l);
        if(wn.staged && nx0 >= wn.x0 && nx1 <= wn.x0 + wn.pitch && ny0 >= wn.y0 && ny1 <= wn.y0 + wn.rows) return;
        const int ax0 = nx0 & ~7, ax1 = (nx1 + 7) & ~7; // -pad_l (144) is a multiple of 8
        int       pitch = ax1 - ax0, rows = ny1 - ny0;
        if(pitch * rows + 8 > win_cap_elems) { // cannot happen with host-side sizing; fail loudly
            if(tid == 0) atomicExch(err_flag, 1);
            rows = (win_cap_elems - 8) / pitch;
        }
        me_team_sync<T>(); // every reader of the previous window is done
        if(tid < 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if(tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(pitch * rows * 2));
            __syncwarp();
            for(int r = lane; r < rows; r += 32)
                bulk_g2s(win + r * pitch, refy + (ptrdiff_t)(ny0 + r) * sref + ax0, (uint32_t)(pitch * 2), bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        wn.x0 = ax0; wn.y0 = ny0; wn.pitch = pitch; wn.rows = rows; wn.staged = 1; wn.biased = 0;
    };
    // bi search: the signed 2*org - pred block is compared through unsigned packed min/max, so both
    // operands carry a ^0x8000 bias during the integer rounds; the window is un-biased for sub-pel
    auto set_bias = [&](int want) {
        if(!bi || wn.biased == want) return;
        uint32_t *w32 = reinterpret_cast<uint32_t *>(win);
        for(int e = tid; e < (wn.pitch * wn.rows) >> 1; e += T) w32[e] ^= 0x80008000u;
        wn.biased = want;
        me_team_sync<T>();
    };

    // Evaluate `ncand` integer candidates (positions produced by `pos(c, px, py)`), then reduce.
    // Returns the winning key (cost << 32 | index), UINT64_MAX if every candidate was out of range.
    // (safe_x, safe_y): any position known to be inside the staged window; groups whose candidate is
    // out of range still run the SAD there so that every lane takes part in the shuffles.
    auto eval_round = [&](int ncand, int safe_x, int safe_y, auto pos) -> uint64_t {
        uint64_t *kb = keys + key_buf * 128;
        key_buf ^= 1;
        uint64_t kmine = ~0ull;
        for(int c = grp; c < ((ncand + Gm::NG - 1) / Gm::NG) * Gm::NG; c += Gm::NG) {
            int  px = 0, py = 0;
            bool live = c < ncand;
            if(live) pos(c, px, py);
            px = (int16_t)px; py = (int16_t)py;
            const bool inr = live && px >= st.lo[0] && px <= st.hi[0] && py >= st.lo[1] && py <= st.hi[1];
            uint32_t   sad = me_group_sad<L2>(win, wn, org, inr ? px : safe_x, inr ? py : safe_y, j);
            if(live && j == 0) {
                uint32_t cost = 0xffffffffu;
                if(inr) {
                    int bits;
                    cost = mv_cost(px << 2, py << 2, bits);
                    sad >>= (bd - 8);
                    cost += bi ? (sad >> 1) : sad;
                }
                if(T == 32) kmine = min(kmine, ((uint64_t)cost << 32) | (uint32_t)c);
                else kb[c] = ((uint64_t)cost << 32) | (uint32_t)c;
            }
        }
        uint64_t k = kmine;
        if(T > 32) {
            __syncthreads();
            for(int c = lane; c < ncand; c += 32) k = min(k, kb[c]);
        }
#pragma unroll
        for(int m = 16; m > 0; m >>= 1) k = min(k, shfl_xor_u64(k, m));
        return k;
    };

    // one me_ipel_diamond run
    auto diamond = [&](int sx, int sy, int patience, int &bx, int &by, int &found_step, int &best_bits) -> uint32_t {
        const int c0x = clip3i(sq.min_clip[0], sq.max_clip[0], sx), c0y = clip3i(sq.min_clip[1], sq.max_clip[1], sy);
        bx = c0x; by = c0y;
        uint32_t best_cost = 0xffffffffu;
        int      misses = 0, step = 0;
        best_bits = 0;
        for(;;) {
            misses++;
            uint64_t k;
            int      this_step, wx0 = 0, wy0 = 0, wnx = 1;
            if(step <= 2) {
                const int r = bi == 1 ? 5 : 2;
                wx0 = bx <= st.lo[0] ? bx : bx - r;
                wy0 = by <= st.lo[1] ? by : by - r;
                const int wx1 = bx >= st.hi[0] ? bx : bx + r, wy1 = by >= st.hi[1] ? by : by + r;
                wnx = wx1 - wx0 + 1;
                const int ncand = wnx * (wy1 - wy0 + 1);
                const int rcp = (65536 + wnx - 1) / wnx; // exact c / wnx for c < 128, wnx <= 11
                k = eval_round(ncand, c0x, c0y, [&](int c, int &px, int &py) { const int qy = (c * rcp) >> 16; px = wx0 + c - qy * wnx; py = wy0 + qy; });
                this_step = 2;
            }
            else if(step <= 8) {
                const int hs = step >> 1, st4 = step == 4;
                k = eval_round(st4 ? 5 : 9, c0x, c0y, [&](int c, int &px, int &py) {
                    const int i = st4 ? c * 2 : c; // step 4 keeps the even entries 0,2,4,6,8
                    // (-2,0) (-1,1) (0,2) (1,1) (2,0) (1,-1) (0,-2) (-1,-1) (0,0)
                    const int dx = i == 8 ? 0 : (i < 4 ? i - 2 : 6 - i);
                    const int dy = i == 8 ? 0 : (i < 2 ? i : (i < 6 ? 4 - i : i - 8));
                    px = c0x + hs * dx; py = c0y + hs * dy;
                });
                this_step = step;
            }
            else {
                const int qs = step >> 2;
                k = eval_round(16, c0x, c0y, [&](int c, int &px, int &py) {
                    // 16-point diamond of radius 4: (-4,0) (-3,1) ... (0,4) ... (4,0) ... (0,-4) ... (-1,-3)... (-3,-1)
                    const int dx = c <= 8 ? c - 4 : 12 - c;
                    const int dy = c <= 4 ? c : (c <= 12 ? 8 - c : c - 16);
                    px = c0x + qs * dx; py = c0y + qs * dy;
                });
                this_step = step;
            }
            const uint32_t kc = (uint32_t)(k >> 32);
            if(kc < best_cost) {
                const int c = (int)(uint32_t)k;
                int       px, py;
                if(step <= 2) { px = wx0 + c % wnx; py = wy0 + c / wnx; }
                else if(step <= 8) {
                    const int i = step == 4 ? c * 2 : c;
                    const int dx = i == 8 ? 0 : (i < 4 ? i - 2 : 6 - i);
                    const int dy = i == 8 ? 0 : (i < 2 ? i : (i < 6 ? 4 - i : i - 8));
                    px = c0x + (step >> 1) * dx; py = c0y + (step >> 1) * dy;
                }
                else {
                    const int dx = c <= 8 ? c - 4 : 12 - c;
                    const int dy = c <= 4 ? c : (c <= 12 ? 8 - c : c - 16);
                    px = c0x + (step >> 2) * dx; py = c0y + (step >> 2) * dy;
                }
                bx = (int16_t)px; by = (int16_t)py;
                best_cost = kc; found_step = this_step; misses = 0;
                int bits;
                mv_cost(bx << 2, by << 2, bits);
                best_bits = bits;
            }
            if(step <= 2) {
                set_window(bx, by, bi == 1);
                step += 2;
            }
            if(misses == patience || bi == 1) break;
            step <<= 1;
            if(step > static_range) break;
        }
        return best_cost;
    };

    // ---- pinter_me_epzs ----------------------------------------------------------------------------
    const int start_x = bi == 1 ? mv_x : mvp_x, start_y = bi == 1 ? mv_y : mvp_y;
    {
        const int cx = clip3i(sq.min_clip[0], sq.max_clip[0], x + (start_x >> 2));
        const int cy = clip3i(sq.min_clip[1], sq.max_clip[1], y + (start_y >> 2));
        set_window(cx, cy, bi == 1);
    }
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
    while(bi != 1 && beststep > 0 && sq.me_complexity > 0) {
        set_window(x + (mv_x >> 2), y + (mv_y >> 2), 0);
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

    if(sq.me_level > 1) {
        // ---- me_spel_pattern: every candidate = 8-tap interpolation of the whole CU + SAD -----------
        int       smv_x = mv_x, smv_y = mv_y, sbits = 0;
        const int16_t obias = bi ? (int16_t)0x8000 : (int16_t)0;
        uint32_t  sbest = 0xffffffffu;
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
                    __syncthreads(); // red[] free, tmp[] reads done
                    if(lane == 0) red[tid >> 5] = (int32_t)part;
                    __syncthreads();
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
        // me_ipel_refinement (src_base/xeve_pinter.c:272-361): 3x3 around the best integer position
        set_window(x + (mv_x >> 2), y + (mv_y >> 2), bi == 1);
        const int ix = clip3i(sq.min_clip[0], sq.max_clip[0], ((int16_t)(mv_x + (x << 2))) >> 2);
        const int iy = clip3i(sq.min_clip[1], sq.max_clip[1], ((int16_t)(mv_y + (y << 2))) >> 2);
        ensure_window(ix, iy, 2);
        set_bias(1);
        const uint64_t k = eval_round(9, ix, iy, [&](int cc, int &px, int &py) {
            // (0,0) (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)
            const int dx9 = cc == 0 ? 0 : (cc <= 3 ? -1 : (cc <= 5 ? 0 : 1));
            const int dy9 = cc == 0 ? 0 : (cc <= 3 ? cc - 2 : (cc == 4 ? -1 : (cc == 5 ? 1 : cc - 7)));
            px = ix + dx9; py = iy + dy9;
        });
        const uint32_t kc = (uint32_t)(k >> 32);
        int rx = ix, ry = iy, rbits = 0;
        if(kc != 0xffffffffu) {
            const int cc = (int)(uint32_t)k;
            const int dx9 = cc == 0 ? 0 : (cc <= 3 ? -1 : (cc <= 5 ? 0 : 1));
            const int dy9 = cc == 0 ? 0 : (cc <= 3 ? cc - 2 : (cc == 4 ? -1 : (cc == 5 ? 1 : cc - 7)));
            rx = ix + dx9; ry = iy + dy9;
            mv_cost(rx << 2, ry << 2, rbits);
        }
        if(bi != 1 && rbits > 0) mot_bits_l = rbits;
        if(kc < best) { best = kc; mv_x = (int16_t)((rx - x) << 2); mv_y = (int16_t)((ry - y) << 2); }
    }

    if(tid == 0) {
        it->mv_out[0] = (int16_t)mv_x; it->mv_out[1] = (int16_t)mv_y; it->cost = best;
        it->mot_bits_out[lidx] = mot_bits_l;
        it->mot_bits_out[lidx ? 0 : 1] = other_bits;
    }
}


