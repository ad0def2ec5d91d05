// play_generic.cuh — K2 generic (any feature list, lookahead 1 / 2, overhang slides, pressure) and K2w (one warp per game) (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "play_common.cuh"

namespace tb {

template <int SLIDE>
TB_D int collect_slides(const uint32_t c[COLS], const Heights& H, const Tables& tab, int piece, const SlotDesc& d0, int s0, int y0, bool slide_on, uint16_t* list)
{
    int n = 0;
    if (SLIDE && slide_on)
        for_each_slide(c, H, d0, s0, y0, [&](int s2) { return load_slot12(tab, piece, s2); },
                       [&](const SlotDesc&, int s2, int yy, int) { list[n++] = (uint16_t)((uint32_t)s2 | (uint32_t)yy << 8); });
    return n;
}

// layer search: every move of `piece` on board c (hard drops, plus overhang slides with SLIDE), each scored as a
// leaf whose previous state is the board itself.  Lanes take slots gl, gl+LANES, ...; inside a lane moves are
// visited in enumeration order, so a strict > keeps the earliest.  tag = tag_hi | slot << 5 | slide index;
// mv = the placement to commit if this leaf wins: the leaf's own (leaf_is_first) or the caller's first move mv1.
template <int LANES, int MODE, int SLIDE>
TB_D void search_one(const uint32_t c[COLS], const Heights& H, const PrevStat& pv, int piece, int gl,
                     const Tables& tab, const uint8_t* ids, const double* w, int F, uint64_t rmask,
                     bool slide_on, const Pressure& pr, int level,
                     uint32_t tag_hi, uint32_t mv1, bool leaf_is_first, uint64_t& key, uint32_t& tag, uint32_t& mv, uint32_t& cnt)
{
    const int ns = tab.nslots[piece];
    for (int s = gl; s < ns; s += LANES) {
        const SlotDesc d0 = load_slot12(tab, piece, s);
        Placement P0;
        if (!place_slot(c, H, d0, P0)) continue;
        uint16_t list[SLIDE ? 18 : 1];
        const int n = collect_slides<SLIDE>(c, H, tab, piece, d0, s, P0.y, slide_on, list);
        for (int i = 0; i <= n; ++i) {
            const uint32_t m = i == 0 ? ((uint32_t)s | (uint32_t)P0.y << 8) : (uint32_t)list[SLIDE ? i - 1 : 0];
            Placement P; SlotDesc d;
            if (SLIDE) {
                d = load_slot12(tab, piece, (int)(m & 255u));
                if (pr.on && !is_time(d, (int)(m >> 8), pr, level)) continue;      // move_with_pressure filter
                uint32_t full;
                place_at(c, d, (int)(m >> 8), P, &full);
                if (full != 0u) place_clear(d, P, full);
            } else {
                d = d0; P = P0;
            }
            Feat f;
            eval_features<ModeTraits<MODE>::CMASK>(P, d, pv, rmask, f);
            const uint64_t k = score_key(score_features<MODE>(f, ids, w, F));
            if (k > key) { key = k; tag = tag_hi | (uint32_t)s << 5 | (uint32_t)i; mv = leaf_is_first ? m : mv1; }
            ++cnt;
        }
    }
}

template <int LANES, int MODE, int LOOKAHEAD, int SLIDE>
__global__ void __launch_bounds__(BLOCK) play_games_kernel(const PlayParams prm)
{
    constexpr int GROUPS = BLOCK / LANES;
    __shared__ Tables s_tab;
    __shared__ double s_w[MODE == MODE_W6 ? 1 : GROUPS][MODE == MODE_W6 ? 1 : TB200_MAX_WEIGHTED];
    load_tables<BLOCK>(s_tab);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const int grp = threadIdx.x / LANES;
    const int F = prm.F;
    const uint32_t T = prm.T;

    // per-game state (identical in every lane of the group)
    bool active = false, exhausted = false;
    uint32_t game = 0, t = 0, lines = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
    uint64_t gscore = 0, stream = 0;
    const uint8_t* seq = nullptr;
    uint32_t c[COLS];
    Heights H; PrevStat pv;
    double wreg[6];
    int p_cur = 0, p_next = 0;

    auto fetch_piece = [&](uint32_t tt) -> int {
        if (tt >= T) return 0;
        return seq ? (int)__ldg(seq + tt) : hashed_piece(prm.seed, stream, tt);
    };

    for (;;) {
        // ---- ticket: groups without a game fetch the next one (shuffles stay warp-convergent) ----
        const bool want = !active && !exhausted;
        uint32_t g = 0;
        if (want && gl == 0) g = atomicAdd(prm.ticket, 1u);
        g = __shfl_sync(FULLMASK, g, 0, LANES);
        __syncwarp(FULLMASK);                    // previous game's reads of s_w are complete
        if (want) {
            if (g < prm.n_games) {
                if (prm.seq_major_C) g = (g % prm.seq_major_C) * (uint32_t)prm.G + g / prm.seq_major_C;     // see PlayParams
                game = g; active = true; t = 0; h1 = h2 = h3 = h4 = 0; cnt = 0; gscore = 0;
                lines = prm.init_lines ? prm.init_lines[g] : 0u;
                const uint32_t sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[g] : (prm.per_game_weights != 0 ? g : g % (uint32_t)prm.G);
                stream = sq;
                seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
                if (prm.init_rows) {
                    uint16_t rows[ROWS];
                    for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)g * ROWS + r];
                    rows_to_cols(rows, c);
                } else {
#pragma unroll
                    for (int k = 0; k < COLS; ++k) c[k] = 0;
                }
                pv = prev_stat(c); H = pack_gh(pv.gh);
                const size_t wrow = prm.per_game_weights > 0 ? (size_t)g : (prm.per_game_weights < 0 ? 0 : (size_t)(g / (uint32_t)prm.G));
                const double* wsrc = prm.weights + wrow * (size_t)F;
                if (MODE == MODE_W6) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) wreg[i] = __ldg(wsrc + i);
                } else {
                    for (int i = gl; i < F; i += LANES) s_w[MODE == MODE_W6 ? 0 : grp][i] = __ldg(wsrc + i);
                }
                p_cur = fetch_piece(0); p_next = fetch_piece(1);
            } else {
                exhausted = true;
            }
        }
        __syncwarp(FULLMASK);                    // s_w of a freshly started game is visible to its group
        if (!__any_sync(FULLMASK, active)) break;

        const double* w = MODE == MODE_W6 ? wreg : s_w[MODE == MODE_W6 ? 0 : grp];
        const int p_after = active ? fetch_piece(t + 2) : 0;      // prefetch: consumed two steps later

        uint64_t key = 0; uint32_t tag = 0, mv = 0;
        const bool slide_on = (prm.move_set & 1) != 0;
        if (LOOKAHEAD == 2) {
            if (active && t + 1 < T && p_cur < 7 && p_next < 7) {
                // layer 1 (group-uniform): every move of the current piece in enumeration order, expanded by the next piece
                const int ns1 = s_tab.nslots[p_cur];
                for (int s1 = 0; s1 < ns1; ++s1) {
                    const SlotDesc d0 = load_slot12(s_tab, p_cur, s1);
                    Placement P0;
                    if (!place_slot(c, H, d0, P0)) continue;
                    uint16_t list[SLIDE ? 18 : 1];
                    const int n = collect_slides<SLIDE>(c, H, s_tab, p_cur, d0, s1, P0.y, slide_on, list);
                    for (int i = 0; i <= n; ++i) {
                        const uint32_t m1 = i == 0 ? ((uint32_t)s1 | (uint32_t)P0.y << 8) : (uint32_t)list[SLIDE ? i - 1 : 0];
                        Placement P1; SlotDesc d1;
                        if (SLIDE) {
                            d1 = load_slot12(s_tab, p_cur, (int)(m1 & 255u));
                            if (prm.press.on && !is_time(d1, (int)(m1 >> 8), prm.press, (int)(lines / 10u))) continue;
                            uint32_t full;
                            place_at(c, d1, (int)(m1 >> 8), P1, &full);
                            if (full != 0u) place_clear(d1, P1, full);
                        } else {
                            d1 = d0; P1 = P0;
                        }
                        const PrevStat pv1 = next_stat(pv, P1, d1);
                        const Heights H1 = pack_gh(pv1.gh);
                        search_one<LANES, MODE, SLIDE>(P1.n, H1, pv1, p_next, gl, s_tab, prm.ids, w, F, prm.rmask,
                                                       slide_on, prm.press, (int)((lines + (uint32_t)P1.k) / 10u),
                                                       ((uint32_t)s1 << 5 | (uint32_t)i) << 11, m1, false, key, tag, mv, cnt);
                    }
                }
            }
            group_argmax_mv<LANES>(key, tag, mv);
        }
        const bool need_l1 = active && p_cur < 7 && (LOOKAHEAD == 1 || key == 0);
        if (LOOKAHEAD == 1 || __any_sync(FULLMASK, need_l1)) {
            uint64_t k1 = 0; uint32_t t1 = 0, m1 = 0;
            if (need_l1) search_one<LANES, MODE, SLIDE>(c, H, pv, p_cur, gl, s_tab, prm.ids, w, F, prm.rmask, slide_on, prm.press, (int)(lines / 10u), 0u, 0u, true, k1, t1, m1, cnt);
            group_argmax_mv<LANES>(k1, t1, m1);
            if (need_l1) { key = k1; tag = t1; mv = m1; }
        }

        if (active) {
            bool finish = false;
            if (key == 0) {
                finish = true;                                   // no legal placement: game over (simulator.py:187-189)
            } else {
                const SlotDesc d1 = load_slot12(s_tab, p_cur, (int)(mv & 255u));
                Placement P;
                {
                    uint32_t full;
                    place_at(c, d1, (int)(mv >> 8), P, &full);
                    if (full != 0u) place_clear(d1, P, full);
                }
#pragma unroll
                for (int k = 0; k < COLS; ++k) c[k] = P.n[k];
                pv = next_stat(pv, P, d1); H = pack_gh(pv.gh);
                lines += (uint32_t)P.k;
                h1 += P.k == 1; h2 += P.k == 2; h3 += P.k == 3; h4 += P.k == 4;
                const uint32_t pts = prm.pts[P.k];
                gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);          // game.py:159-164
                if (gl == 0) {
                    if (prm.trace) {
                        const uint32_t row = (uint32_t)(ROWS - P.y - P.ph);
                        const uint32_t rec = (uint32_t)P.rot | row << 8 | (uint32_t)P.c0 << 16 | (uint32_t)P.k << 24;
                        reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = rec;
                    }
                    if (prm.score_trace) {
                        const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                        prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                    }
                }
                ++t;
                p_cur = p_next; p_next = p_after;
                if (t >= T || t >= prm.max_moves) finish = true;
            }
            if (finish) {
                if (gl == 0) {
                    tb200_game_result* r = prm.results + game;
                    r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4;
                    r->score = gscore;
                    if (prm.final_rows) {
                        uint16_t rows[ROWS];
                        cols_to_rows(c, rows);
                        for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2];
                    }
                }
                if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&prm.results[game].placements), (unsigned long long)cnt);
                active = false;
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(BLOCK) play_games_w32_kernel(const PlayParams prm)
{
    constexpr int GROUPS = BLOCK / 32;
    __shared__ Tables s_tab;
    __shared__ double s_w[MODE == MODE_W6 ? 1 : GROUPS][MODE == MODE_W6 ? 1 : TB200_MAX_WEIGHTED];
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x >> 5;
    const int F = prm.F;
    const uint32_t T = prm.T;
    double wreg[6];
    const double* w = MODE == MODE_W6 ? wreg : s_w[MODE == MODE_W6 ? 0 : grp];

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
            if (MODE == MODE_W6) {
#pragma unroll
                for (int i = 0; i < 6; ++i) wreg[i] = __ldg(wsrc + i);
            } else {
                __syncwarp(FULLMASK);
                for (int i = lane; i < F; i += 32) s_w[MODE == MODE_W6 ? 0 : grp][i] = __ldg(wsrc + i);
                __syncwarp(FULLMASK);
            }
        }
        auto fetch_piece = [&](uint32_t tt) -> int {
            if (tt >= T) return 7;
            return seq ? (int)__ldg(seq + tt) : hashed_piece(prm.seed, stream, tt);
        };
        uint32_t t = 0, lines = prm.init_lines ? prm.init_lines[game] : 0u, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
        uint64_t gscore = 0;
        int p_cur = fetch_piece(0), p_next = fetch_piece(1);
        const uint32_t max_moves = prm.max_moves;

#ifdef TB_PHASES
        long long ph[8] = {0,0,0,0,0,0,0,0}; long long last_ = clock64();
#endif
        while (t < T && t < max_moves && p_cur < 7) {
            PH(5)
            const int p_after = fetch_piece(t + 2);
            const int ns = s_tab.nslots[p_cur];
            // ---- stream A: slot = lane ----
            const bool hasA = lane < ns;
            const SlotDesc dA = load_slot12(s_tab, p_cur, hasA ? lane : 0);
            Placement PA; uint32_t fA;
            const bool legalA = place_pre(c, H, dA, PA, &fA) && hasA;
            uint64_t key; uint32_t tag = (uint32_t)lane;
            PH(0)
            if (MODE == MODE_W6 && ns > 32) {
                // ---- stream B: slot = lane + 32 (only lanes 0, 1 of the 34-slot pieces), interleaved with A ----
                const bool hasB = lane + 32 < ns;
                const SlotDesc dB = load_slot12(s_tab, p_cur, hasB ? lane + 32 : 0);
                Placement PB; uint32_t fB;
                const bool legalB = place_pre(c, H, dB, PB, &fB) && hasB;
                if ((legalA && fA != 0u) || (legalB && fB != 0u)) {
                    if (legalA && fA != 0u) place_clear(dA, PA, fA);
                    if (legalB && fB != 0u) place_clear(dB, PB, fB);
                }
                const uint64_t kA = eval_key<MODE>(PA, dA, legalA, pv, prm.ids, w, F, prm.rmask);
                const uint64_t kB = eval_key<MODE>(PB, dB, legalB, pv, prm.ids, w, F, prm.rmask);
                cnt += (uint32_t)legalA + (uint32_t)legalB;
                key = kA;
                if (kB > kA) {                      // strict: the earlier slot keeps ties
                    key = kB; tag = (uint32_t)lane + 32u;
#pragma unroll
                    for (int k = 0; k < COLS; ++k) PA.n[k] = PB.n[k];
                    PA.y = PB.y; PA.k = PB.k; PA.rot = PB.rot; PA.c0 = PB.c0; PA.ph = PB.ph;
                }
            } else {
                if (legalA && fA != 0u) place_clear(dA, PA, fA);
                key = eval_key<MODE>(PA, dA, legalA, pv, prm.ids, w, F, prm.rmask);
                cnt += (uint32_t)legalA;
                if (MODE != MODE_W6 && ns > 32) {   // big feature sets: second pass after the first (register budget)
                    const bool hasB = lane + 32 < ns;
                    const SlotDesc dB = load_slot12(s_tab, p_cur, hasB ? lane + 32 : 0);
                    Placement PB;
                    const bool legalB = place_slot(c, H, dB, PB) && hasB;
                    if (__any_sync(FULLMASK, legalB)) {
                        const uint64_t kB = eval_key<MODE>(PB, dB, legalB, pv, prm.ids, w, F, prm.rmask);
                        cnt += (uint32_t)legalB;
                        if (kB > key) {
                            key = kB; tag = (uint32_t)lane + 32u;
#pragma unroll
                            for (int k = 0; k < COLS; ++k) PA.n[k] = PB.n[k];
                            PA.y = PB.y; PA.k = PB.k; PA.rot = PB.rot; PA.c0 = PB.c0; PA.ph = PB.ph;
                        }
                    }
                }
            }
            PH(ns > 32 ? 2 : 1)
            warp_argmax32(key, tag);
            PH(3)
            if (key == 0) break;                    // no legal placement: game over (simulator.py:187-189)
            // ---- commit: broadcast the winner's candidate board ----
            const int src = (int)(tag & 31u);
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = __shfl_sync(FULLMASK, PA.n[k], src);
            const uint32_t info = __shfl_sync(FULLMASK, (uint32_t)PA.y | (uint32_t)PA.k << 8 | (uint32_t)PA.rot << 12 | (uint32_t)PA.c0 << 16 | (uint32_t)PA.ph << 24, src);
            const uint32_t wk = (info >> 8) & 15u;
            {   // statistics of the committed board: arithmetic from the winner's move unless it cleared rows
                Placement Pw;
                Pw.y = (int)(info & 255u); Pw.k = (int)wk;
#pragma unroll
                for (int k = 0; k < COLS; ++k) Pw.n[k] = c[k];
                const SlotDesc dw = load_slot12(s_tab, p_cur, (int)tag);
                pv = next_stat(pv, Pw, dw);
                H = pack_gh(pv.gh);
            }
            lines += wk;
            h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
            const uint32_t pts = prm.pts[wk];
            gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
            if (lane == 0) {
                if (prm.trace) {
                    const uint32_t wy = info & 255u, wph = info >> 24;
                    const uint32_t row = (uint32_t)ROWS - wy - wph;
                    reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = ((info >> 12) & 15u) | row << 8 | ((info >> 16) & 255u) << 16 | wk << 24;
                }
                if (prm.score_trace) {
                    const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                    prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                }
            }
            ++t;
            p_cur = p_next; p_next = p_after;
            PH(4)
        }
#ifdef TB_PHASES
        if (lane == 0) { for (int i = 0; i < 6; ++i) atomicAdd(&g_phase[i], (unsigned long long)ph[i]); atomicAdd(&g_phase[6], (unsigned long long)t); }
#endif
        // ---- game finished (over, cap reached or max_moves) ----
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
