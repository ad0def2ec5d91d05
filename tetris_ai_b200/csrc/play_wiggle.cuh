// play_wiggle.cuh — K2x: the wiggle move sets (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "play_common.cuh"

namespace tb {

// ---------------------------------------------------------------------------------------------
// K2x: the wiggle move sets (simulator.py:49-85), one warp per game.
// The moves of a piece are the resting cells reachable from the starts by down / left / right steps.  As a SET they come
// from a warp-parallel bit flood fill (wiggle_rest: one (rotation, column) slot per lane, reachable heights as a bit
// mask, fill_down + neighbour exchange until stable), and every lane scores the moves of its own slots.  The reference's
// visiting ORDER (a depth-first recursion) only matters for its first-of-max tie rule, so the serial reference-order
// search (wiggle_rot, explicit stack in shared memory) runs only when several moves of the deciding rotation tie at
// the maximum, for that one rotation, and stops at the first tied move (wiggle_pick).  With lookahead 2 the first layer
// is listed in reference order (wiggle_collect, one search per rotation on lanes 0..3) because the first move of the
// winning pair is what gets committed; the second layer is order-free.
// ---------------------------------------------------------------------------------------------
struct WiggleShared {
    uint32_t fm[4][12];        // free map per rotation and column (bit y)
    uint32_t seen[4][12];
    uint16_t stack[4][256];
    uint16_t list[4][128];     // moves of each rotation in visiting order: y | col << 8
    uint8_t  starty[4][12];    // hard-drop landing height per rotation and column, 255 = illegal
    uint32_t tie[4][12];       // fast path: moves that tie at the maximum (bit y), per rotation and column
    int cnt[4], base[4], wph[4];
};

// Move list of `piece` on board c (warp-cooperative): free maps + hard-drop landings by all lanes, one flood fill per
// rotation by lanes 0..nrot-1.  off[r] = number of moves of rotations < r; returns the total.
TB_D int wiggle_collect(WiggleShared& wg, const Tables& tab, const uint32_t c[COLS], const Heights& H, int piece, bool from_top, int lane, int off[5])
{
    const int ns = tab.nslots[piece];
    const int nrot = (int)((tab.slot[piece][ns - 1].w[0] >> 4) & 3u) + 1;
    __syncwarp(FULLMASK);
    for (int i = lane; i < 48; i += 32) { wg.fm[i / 12][i % 12] = 0u; wg.starty[i / 12][i % 12] = 255; }
    __syncwarp(FULLMASK);
    for (int s = lane; s < ns; s += 32) {
        const SlotDesc d = load_slot12(tab, piece, s);
        const int rot = (int)((d.w[0] >> 4) & 3u), c0 = (int)(d.w[0] & 15u);
        wg.fm[rot][c0] = free_mask(c, d);
        Placement P; uint32_t full;
        if (place_pre(c, H, d, P, &full)) wg.starty[rot][c0] = (uint8_t)P.y;
        if (c0 == 0) { wg.base[rot] = s; wg.wph[rot] = (int)((d.w[0] >> 6) & 7u) | (int)((d.w[0] >> 9) & 7u) << 8; }
    }
    __syncwarp(FULLMASK);
    if (lane < nrot) {
        const int wd = wg.wph[lane] & 255, ph = wg.wph[lane] >> 8;
        uint16_t starts[COLS]; int nst = 0;
        for (int col = 0; col <= COLS - wd; ++col) {
            if (from_top) starts[nst++] = (uint16_t)((uint32_t)(ROWS - ph) | (uint32_t)col << 8);
            else if (wg.starty[lane][col] != 255) starts[nst++] = (uint16_t)((uint32_t)wg.starty[lane][col] | (uint32_t)col << 8);
        }
        wg.cnt[lane] = wiggle_rot(wg.fm[lane], wd, starts, nst, wg.stack[lane], wg.seen[lane], wg.list[lane], 128);
    }
    __syncwarp(FULLMASK);
    off[0] = 0;
    off[1] = wg.cnt[0];
    off[2] = off[1] + (nrot > 1 ? wg.cnt[1] : 0);
    off[3] = off[2] + (nrot > 2 ? wg.cnt[2] : 0);
    off[4] = off[3] + (nrot > 3 ? wg.cnt[3] : 0);
    return off[4];
}

TB_D void wiggle_move_at(const WiggleShared& wg, const int off[5], int idx, int& slot, int& y)
{
    const int r = idx < off[1] ? 0 : (idx < off[2] ? 1 : (idx < off[3] ? 2 : 3));
    const uint32_t e = wg.list[r][idx - off[r]];
    y = (int)(e & 31u); slot = wg.base[r] + (int)(e >> 8);
}

// score every move of the list as a leaf on board c (previous-state statistics pv); tag = tag_hi | position in the list
template <int MODE>
TB_D void wiggle_score(const WiggleShared& wg, const int off[5], const Tables& tab, const uint32_t c[COLS], const PrevStat& pv, int piece,
                       const Pressure& pr, int level, const uint8_t* ids, const double* w, int F, uint64_t rmask, uint32_t tag_hi, int lane,
                       uint64_t& key, uint32_t& tag, uint32_t& cnt)
{
    for (int idx = lane; idx < off[4]; idx += 32) {
        int slot, y;
        wiggle_move_at(wg, off, idx, slot, y);
        const SlotDesc d = load_slot12(tab, piece, slot);
        if (pr.on && !is_time(d, y, pr, level)) continue;
        Placement P; uint32_t full;
        place_at(c, d, y, P, &full);
        if (full != 0u) place_clear(d, P, full);
        Feat f;
        eval_features<ModeTraits<MODE>::CMASK>(P, d, pv, rmask, f);
        const uint64_t k = score_key(score_features<MODE>(f, ids, w, F));
        if (k > key) { key = k; tag = tag_hi | (uint32_t)idx; }
        ++cnt;
    }
}

// Order-free fast path: the SET of resting cells of every slot by a warp-parallel flood fill.  Lane l owns slots l and
// l + 32; each keeps the reachable cells of its slot as a bit mask (bit y), grows it downwards with fill_down and
// sideways from its two neighbours of the same rotation (exchanged through wg.seen, index column + 1, zero padded)
// until no lane changes.  rest[j] = reachable cells with no free cell below = the moves of slot l + 32 j.  Also leaves
// the free maps, hard-drop landings and per-rotation slot bases in wg for the tie break.
TB_D void wiggle_rest(WiggleShared& wg, const Tables& tab, const uint32_t c[COLS], const Heights& H, int piece, bool from_top, int lane, uint32_t rest[2])
{
    const int ns = tab.nslots[piece];
    __syncwarp(FULLMASK);
    for (int i = lane; i < 48; i += 32) { wg.seen[i / 12][i % 12] = 0u; wg.fm[i / 12][i % 12] = 0u; wg.starty[i / 12][i % 12] = 255; }
    __syncwarp(FULLMASK);
    uint32_t fm[2] = {0u, 0u}, rv[2] = {0u, 0u};
    int at[2] = {0, 0};                      // rot * 12 + c0 + 1 of the lane's slots
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int sl = lane + 32 * j;
        if (sl < ns) {
            const SlotDesc d = load_slot12(tab, piece, sl);
            const int rot = (int)((d.w[0] >> 4) & 3u), c0 = (int)(d.w[0] & 15u), ph = (int)((d.w[0] >> 9) & 7u);
            fm[j] = free_mask(c, d);
            wg.fm[rot][c0] = fm[j];
            if (c0 == 0) { wg.base[rot] = sl; wg.wph[rot] = (int)((d.w[0] >> 6) & 7u) | ph << 8; }
            uint32_t seed;
            if (from_top) seed = 1u << (ROWS - ph);
            else {
                Placement P; uint32_t full;
                const bool ok = place_pre(c, H, d, P, &full);
                seed = ok ? 1u << P.y : 0u;
                if (ok) wg.starty[rot][c0] = (uint8_t)P.y;
            }
            rv[j] = fill_down(seed & fm[j], fm[j]);
            at[j] = rot * 12 + c0 + 1;
            (&wg.seen[0][0])[at[j]] = rv[j];
        }
    }
    for (;;) {
        __syncwarp(FULLMASK);
        uint32_t nv[2] = {0u, 0u};
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (lane + 32 * j < ns) {
                const uint32_t* sn = &wg.seen[0][0] + at[j];
                nv[j] = fill_down((sn[-1] | sn[1] | rv[j]) & fm[j], fm[j]);
            }
        const bool changed = nv[0] != rv[0] || nv[1] != rv[1];
        __syncwarp(FULLMASK);
        if (!__any_sync(FULLMASK, changed)) break;
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (nv[j] != rv[j]) { rv[j] = nv[j]; (&wg.seen[0][0])[at[j]] = nv[j]; }
    }
    rest[0] = rv[0] & ~(fm[0] << 1);
    rest[1] = rv[1] & ~(fm[1] << 1);
}

// score the lane's moves (rest masks of wiggle_rest) as leaves on board c; tm[j] = the moves of slot lane + 32 j that
// reach the lane's best key
template <int MODE>
TB_D void wiggle_score_rest(const uint32_t rest[2], const Tables& tab, const uint32_t c[COLS], const PrevStat& pv, int piece,
                            const Pressure& pr, int level, const uint8_t* ids, const double* w, int F, uint64_t rmask, int lane,
                            uint64_t& key, uint32_t tm[2], uint32_t& cnt)
{
    tm[0] = 0u; tm[1] = 0u;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t m = rest[j];
        if (m == 0u) continue;
        const SlotDesc d = load_slot12(tab, piece, lane + 32 * j);
        while (m != 0u) {
            const int y = __ffs((int)m) - 1;
            m &= m - 1u;
            if (pr.on && !is_time(d, y, pr, level)) continue;
            Placement P; uint32_t full;
            place_at(c, d, y, P, &full);
            if (full != 0u) place_clear(d, P, full);
            Feat f;
            eval_features<ModeTraits<MODE>::CMASK>(P, d, pv, rmask, f);
            const uint64_t k = score_key(score_features<MODE>(f, ids, w, F));
            ++cnt;
            if (k > key) { key = k; tm[0] = 0u; tm[1] = 0u; tm[j] = 1u << y; }
            else if (k == key) tm[j] |= 1u << y;
        }
    }
}

// The winner among the lanes' best moves under the reference's first-of-max rule.  A unique maximum needs no order;
// otherwise the tied moves of the lowest rotation that has any decide: one such move wins outright, several are
// ordered by running the reference-order search of that one rotation (lane 0) up to the first of them.
// Returns false when there is no move at all; else slot / y of the winner and key = the maximum.
TB_D bool wiggle_pick(WiggleShared& wg, const Tables& tab, int piece, bool from_top, int lane, uint64_t& key, const uint32_t tm_in[2], int& slot, int& y)
{
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mhi = __reduce_max_sync(FULLMASK, hi);
    const uint32_t mlo = __reduce_max_sync(FULLMASK, hi == mhi ? lo : 0u);
    key = ((uint64_t)mhi << 32) | mlo;
    if (key == 0) return false;
    const bool win = hi == mhi && lo == mlo;
    const uint32_t tm[2] = {win ? tm_in[0] : 0u, win ? tm_in[1] : 0u};
    // lowest tied slot (slots are rotation-major) -> the rotation that decides
    const uint32_t smin = __reduce_min_sync(FULLMASK, tm[0] != 0u ? (uint32_t)lane : (tm[1] != 0u ? (uint32_t)lane + 32u : 99u));
    const int rstar = (int)((tab.slot[piece][smin].w[0] >> 4) & 3u);
    uint32_t mine = 0u, mymv = 0u;
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (tm[j] != 0u && (int)((tab.slot[piece][lane + 32 * j].w[0] >> 4) & 3u) == rstar) {
            mine += (uint32_t)__popc(tm[j]);
            mymv = (uint32_t)(lane + 32 * j) | (uint32_t)(__ffs((int)tm[j]) - 1) << 8;
        }
    const uint32_t total = __reduce_add_sync(FULLMASK, mine);
    if (total == 1u) {
        const uint32_t who = __ballot_sync(FULLMASK, mine != 0u);
        const uint32_t mv = __shfl_sync(FULLMASK, mymv, __ffs((int)who) - 1);
        slot = (int)(mv & 255u); y = (int)(mv >> 8);
        return true;
    }
    // several tied moves in rotation rstar: visiting order decides
    __syncwarp(FULLMASK);
    if (lane < 12) wg.tie[rstar][lane] = 0u;
    __syncwarp(FULLMASK);
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (tm[j] != 0u) {
            const uint32_t w0 = tab.slot[piece][lane + 32 * j].w[0];
            if ((int)((w0 >> 4) & 3u) == rstar) wg.tie[rstar][w0 & 15u] = tm[j];
        }
    __syncwarp(FULLMASK);
    if (lane == 0) {
        const int wd = wg.wph[rstar] & 255, ph = wg.wph[rstar] >> 8;
        uint16_t starts[COLS]; int nst = 0;
        for (int col = 0; col <= COLS - wd; ++col) {
            if (from_top) starts[nst++] = (uint16_t)((uint32_t)(ROWS - ph) | (uint32_t)col << 8);
            else if (wg.starty[rstar][col] != 255) starts[nst++] = (uint16_t)((uint32_t)wg.starty[rstar][col] | (uint32_t)col << 8);
        }
        wiggle_rot(wg.fm[rstar], wd, starts, nst, wg.stack[0], wg.seen[0], wg.list[0], 128, wg.tie[rstar]);
    }
    __syncwarp(FULLMASK);
    const uint32_t e = wg.list[0][0];
    y = (int)(e & 31u); slot = wg.base[rstar] + (int)(e >> 8);
    return true;
}

template <int MODE, int LOOKAHEAD>
__global__ void __launch_bounds__(BLOCK) play_games_wiggle_kernel(const PlayParams prm)
{
    constexpr int GROUPS = BLOCK / 32;
    __shared__ Tables s_tab;
    __shared__ WiggleShared s_wg[GROUPS][LOOKAHEAD];
    __shared__ double s_w[GROUPS][TB200_MAX_WEIGHTED];
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x >> 5;
    WiggleShared& wg1 = s_wg[grp][0];
    WiggleShared& wg2 = s_wg[grp][LOOKAHEAD - 1];
    const int F = prm.F;
    const uint32_t T = prm.T;
    const bool from_top = (prm.move_set & 4) != 0;
    const double* w = s_w[grp];

    for (;;) {
        uint32_t game = 0;
        if (lane == 0) game = atomicAdd(prm.ticket, 1u);
        game = __shfl_sync(FULLMASK, game, 0);
        if (game >= prm.n_games) break;
        const uint32_t sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[game] : (prm.per_game_weights != 0 ? game : game % (uint32_t)prm.G);
        const uint64_t stream = sq;
        const uint8_t* seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
        uint32_t c[COLS];
        if (prm.init_rows) {
            uint16_t rows[ROWS];
            for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
            rows_to_cols(rows, c);
        } else {
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = 0;
        }
        PrevStat pv = prev_stat(c);
        Heights H = pack_gh(pv.gh);
        {
            const size_t wrow = prm.per_game_weights > 0 ? (size_t)game : (prm.per_game_weights < 0 ? 0 : (size_t)(game / (uint32_t)prm.G));
            const double* wsrc = prm.weights + wrow * (size_t)F;
            __syncwarp(FULLMASK);
            for (int i = lane; i < F; i += 32) s_w[grp][i] = __ldg(wsrc + i);
            __syncwarp(FULLMASK);
        }
        auto fetch_piece = [&](uint32_t tt) -> int {
            if (tt >= T) return 7;
            return seq ? (int)__ldg(seq + tt) : hashed_piece(prm.seed, stream, tt);
        };
        uint32_t t = 0, lines = prm.init_lines ? prm.init_lines[game] : 0u, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
        uint64_t gscore = 0;
        const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
        int p_cur = fetch_piece(0), p_next = fetch_piece(1);

        while (t < stop && p_cur < 7) {
            const int p_after = fetch_piece(t + 2);
            const int level = (int)(lines / 10u);
            uint64_t key = 0;
            int slot = 0, y = 0;
            if (LOOKAHEAD == 1 || !(t + 1 < T && p_next < 7)) {
                // order-free fast path: the move SET by warp-parallel flood fill; the reference's visiting order only
                // matters when several moves tie at the maximum (wiggle_pick)
                uint32_t rest[2], tm[2];
                wiggle_rest(wg1, s_tab, c, H, p_cur, from_top, lane, rest);
                wiggle_score_rest<MODE>(rest, s_tab, c, pv, p_cur, prm.press, level, prm.ids, w, F, prm.rmask, lane, key, tm, cnt);
                if (!wiggle_pick(wg1, s_tab, p_cur, from_top, lane, key, tm, slot, y)) break;
            } else {
                // lookahead 2: every first move in the reference's list order (warp-uniform), expanded by the set flood fill of the next piece
                int off1[5];
                uint32_t tag = 0;
                wiggle_collect(wg1, s_tab, c, H, p_cur, from_top, lane, off1);
                for (int i1 = 0; i1 < off1[4]; ++i1) {
                    int slot1, y1;
                    wiggle_move_at(wg1, off1, i1, slot1, y1);
                    const SlotDesc d1 = load_slot12(s_tab, p_cur, slot1);
                    if (prm.press.on && !is_time(d1, y1, prm.press, level)) continue;
                    Placement P1; uint32_t full;
                    place_at(c, d1, y1, P1, &full);
                    if (full != 0u) place_clear(d1, P1, full);
                    const PrevStat pv1 = next_stat(pv, P1, d1);
                    const Heights H1 = pack_gh(pv1.gh);
                    // the second layer is order-free: only the FIRST move of the winning pair is committed, so among equal
                    // keys the earliest i1 wins whatever the order inside its second layer
                    uint32_t rest[2], tm[2];
                    uint64_t k2 = 0;
                    wiggle_rest(wg2, s_tab, P1.n, H1, p_next, from_top, lane, rest);
                    wiggle_score_rest<MODE>(rest, s_tab, P1.n, pv1, p_next, prm.press, (int)((lines + (uint32_t)P1.k) / 10u), prm.ids, w, F, prm.rmask,
                                            lane, k2, tm, cnt);
                    if (k2 > key) { key = k2; tag = (uint32_t)i1 << 11; }
                }
                warp_argmax32(key, tag);
                bool leaf_is_first = key == 0;
                if (leaf_is_first) {              // no second-layer leaf anywhere: score the first layer itself
                    tag = 0;
                    wiggle_score<MODE>(wg1, off1, s_tab, c, pv, p_cur, prm.press, level, prm.ids, w, F, prm.rmask, 0u, lane, key, tag, cnt);
                    warp_argmax32(key, tag);
                }
                if (key == 0) break;
                wiggle_move_at(wg1, off1, (int)(leaf_is_first ? tag : (tag >> 11)), slot, y);
            }
            // ---- commit the (first) move of the winner ----
            const SlotDesc d = load_slot12(s_tab, p_cur, slot);
            Placement P; uint32_t full;
            place_at(c, d, y, P, &full);
            if (full != 0u) place_clear(d, P, full);
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = P.n[k];
            pv = next_stat(pv, P, d); H = pack_gh(pv.gh);
            const uint32_t wk = (uint32_t)P.k;
            if (wk != 0u) {
                lines += wk;
                h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                const uint32_t pts = prm.pts[wk];
                gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
            } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
            if (lane == 0) {
                if (prm.trace) {
                    const uint32_t row = (uint32_t)(ROWS - P.y - P.ph);
                    reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = (uint32_t)P.rot | row << 8 | (uint32_t)P.c0 << 16 | wk << 24;
                }
                if (prm.score_trace) {
                    const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                    prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                }
            }
            ++t;
            p_cur = p_next; p_next = p_after;
        }
        cnt = __reduce_add_sync(FULLMASK, cnt);
        if (lane == 0) {
            tb200_game_result* r = prm.results + game;
            r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4;
            r->score = gscore; r->placements = cnt;
            if (prm.final_rows) {
                uint16_t rows[ROWS];
                cols_to_rows(c, rows);
                for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2];
            }
        }
    }
}

}  // namespace tb
