// play_population.cuh — K2g, K2s (lockstep) and K2b (piece-binned): the w6 throughput kernels (BASELINE configs[2] and [4]) (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "play_common.cuh"

namespace tb {

// ---------------------------------------------------------------------------------------------
// K2g: the throughput kernel for the w6 set, lookahead 1: LANES (8 or 16) lanes per game, 32/LANES
// games per warp, for populations larger than one resident wave (BASELINE configs[2] / [4]).
// Same fast evaluation as K2f (w6_pre / w6_clear / w6_post: no clz, popc only for the transition
// counts), slots in rounds of LANES with a per-lane running best, shuffle-butterfly argmax inside
// the group, commit by re-placing the winning slot with w6_pre (its heights come out of the
// placement arithmetic, so the commit needs no clz either).  Persistent, ticket per group.
// ---------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(BLOCK) play_games_w6grp_kernel(const PlayParams prm)
{
    __shared__ Tables s_tab;
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const uint32_t T = prm.T;
    const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;

    bool active = false, exhausted = false;
    uint32_t game = 0, t = 0, lines = 0, lines0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
    uint64_t gscore = 0, stream = 0;
    const uint8_t* seq = nullptr;
    uint32_t c[COLS];
    GameStat g;
    int cells0 = 0;
    double w[6];
    int p_cur = 7, p_next = 7;
#pragma unroll
    for (int k = 0; k < COLS; ++k) { c[k] = 0; g.gh[k] = 0; }
    g.H.H0 = g.H.H1 = g.H.H2 = 0; g.cells = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = 0.0;

    auto fetch_piece = [&](uint32_t tt) -> int {
        if (tt >= T) return 7;
        return seq ? (int)__ldg(seq + tt) : hashed_piece(prm.seed, stream, tt);
    };

    for (;;) {
        const bool want = !active && !exhausted;
        uint32_t tk = 0;
        if (want && gl == 0) tk = atomicAdd(prm.ticket, 1u);
        tk = __shfl_sync(FULLMASK, tk, 0, LANES);
        if (want) {
            if (tk < prm.n_games) {
                game = prm.seq_major_C ? (tk % prm.seq_major_C) * (uint32_t)prm.G + tk / prm.seq_major_C : tk;
                active = true; t = 0; h1 = h2 = h3 = h4 = 0; cnt = 0; gscore = 0;
                lines = prm.init_lines ? prm.init_lines[game] : 0u; lines0 = lines;
                const uint32_t sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[game] : (prm.per_game_weights != 0 ? game : game % (uint32_t)prm.G);
                stream = sq;
                seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
                if (prm.init_rows) {
                    uint16_t rows[ROWS];
                    for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
                    rows_to_cols(rows, c);
                } else {
#pragma unroll
                    for (int k = 0; k < COLS; ++k) c[k] = 0;
                }
                g = game_stat(c);
                cells0 = g.cells;
                const size_t wrow = prm.per_game_weights > 0 ? (size_t)game : (prm.per_game_weights < 0 ? 0 : (size_t)(game / (uint32_t)prm.G));
                const double* wsrc = prm.weights + wrow * 6;
#pragma unroll
                for (int i = 0; i < 6; ++i) w[i] = __ldg(wsrc + i);
                p_cur = fetch_piece(0); p_next = fetch_piece(1);
                if (stop == 0 || p_cur >= 7) {            // nothing to play: finish immediately
                    if (gl == 0) {
                        tb200_game_result* r = prm.results + game;
                        r->pieces = 0; r->lines = lines; r->score = 0;
                        if (prm.final_rows) {
                            uint16_t rows[ROWS];
                            cols_to_rows(c, rows);
                            for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2];
                        }
                    }
                    active = false;
                }
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(FULLMASK, active)) break;

        const int p_after = active ? fetch_piece(t + 2) : 7;
        uint64_t key = 0; uint32_t tag = 0;
        if (active) {
            const int pc = p_cur < 7 ? p_cur : 0;
            const int ns = p_cur < 7 ? s_tab.nslots[pc] : 0;
            for (int s = gl; s < ns; s += LANES) {
                const SlotDesc d = load_slot12(s_tab, pc, s);
                W6Cand A;
                w6_pre(c, g, d, A);
                if (A.legal && A.full != 0u) w6_clear(d, A);
                Heights Hn;
                const uint64_t k = w6_post(A, w, Hn);
                cnt += (uint32_t)A.legal;
                if (k > key) { key = k; tag = (uint32_t)s; }
            }
        }
        group_argmax<LANES>(key, tag);
        if (active) {
            bool finish = false;
            if (key == 0) {
                finish = true;                            // no legal placement: game over
            } else {
                const SlotDesc d = load_slot12(s_tab, p_cur, (int)tag);
                W6Cand A;
                w6_pre(c, g, d, A);
                if (A.full != 0u) w6_clear(d, A);
#pragma unroll
                for (int k = 0; k < COLS; ++k) { c[k] = A.P.n[k]; g.gh[k] = A.h[k]; }
                g.H.H0 = (uint32_t)A.h[0] | (uint32_t)A.h[1] << 8 | (uint32_t)A.h[2] << 16 | (uint32_t)A.h[3] << 24;
                g.H.H1 = (uint32_t)A.h[4] | (uint32_t)A.h[5] << 8 | (uint32_t)A.h[6] << 16 | (uint32_t)A.h[7] << 24;
                g.H.H2 = (uint32_t)A.h[8] | (uint32_t)A.h[9] << 8;
                const uint32_t wk = (uint32_t)A.P.k;
                if (wk != 0u) {
                    lines += wk;
                    h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                    const uint32_t pts = prm.pts[wk];
                    gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
                } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
                g.cells = cells0 + 4 * (int)(t + 1) - 10 * (int)(lines - lines0);
                if (gl == 0) {
                    if (prm.trace) {
                        const uint32_t row = (uint32_t)(ROWS - A.P.y - A.P.ph);
                        reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = (uint32_t)A.P.rot | row << 8 | (uint32_t)A.P.c0 << 16 | wk << 24;
                    }
                    if (prm.score_trace) {
                        const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                        prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                    }
                }
                ++t;
                p_cur = p_next; p_next = p_after;
                if (t >= stop || p_cur >= 7) finish = true;
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

// ---------------------------------------------------------------------------------------------
// K2s: the lockstep throughput kernel for CE populations (BASELINE configs[2]): C candidates x G games on G SHARED piece
// sequences, w6, lookahead 1.  All games of sequence g see the same piece at the same move, so a warp that holds only
// games of one sequence at one move index runs every round with all of its lanes on the same piece: no lane waits for
// another game's longer slot list (K2g on independent streams keeps 18 of 32 threads busy).  Games die, so the play is
// cut into stages of K pieces (the first two shorter); a task = (stage, sequence, chunk of 32 / LANES games); at the end of a task the
// survivors are appended to the sequence's list for the next stage, which re-packs them into full warps.  One
// persistent launch: warps walk the buckets (stage-major, sequence-minor) and claim tasks by ticket; bucket (s, g)
// opens when every task of (s - 1, g) is done — 29 other buckets lie in between, so nobody actually waits except at
// the very end, and a running task never waits, so the scheme cannot deadlock with a fully resident grid.
// Hand-over state: GameSave (96 bytes per game) through L2 (__ldcg / fence + counter).
// ---------------------------------------------------------------------------------------------
// first move of stage s: two short stages first (K/4, K/2) — a fresh CE population loses most of its games within a
// few dozen pieces and they should be re-packed early — then K pieces per stage
TB_HD uint32_t lock_stage_begin(uint32_t s, uint32_t K) { return s == 0u ? 0u : (s == 1u ? K / 4u : (s == 2u ? K / 2u : K * (s - 2u))); }

// MODE_W6: the fast w6 evaluation.  MODE_ALL45 / MODE_GENERIC: the general placement / feature / scorer code of tetris_core.cuh,
// weights read through L1 from the candidate's row (they do not fit in registers).
// One task of the lockstep schedule: up to 32 / LANES games of sequence sq, all at the first move of stage st, played to the end of
// the stage (or of the game); survivors are appended to the sequence's list of the next stage.  `slot` = task index in the bucket.
template <int LANES, int MODE, bool PRESS = false>      // PRESS: move_with_pressure(move_drop) — the per-move time predicate of simulator.py:88-120
TB_D void lock_task(const PlayParams& prm, const Tables& s_tab, uint32_t st, uint32_t sq, uint32_t slot, uint32_t n_in, uint32_t* ctl)
{
    constexpr bool W6 = MODE == MODE_W6;
    constexpr uint32_t GPW = 32 / LANES;           // games per warp
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const uint32_t T = prm.T, C = prm.lk_C, K = prm.lk_K, S = prm.lk_S;
    const uint32_t G = (uint32_t)prm.G;
    const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
    {
    // ---- task (st, sq, slot): up to GPW games of sequence sq, all at move st * K ----
    const uint32_t idx = slot * GPW + (uint32_t)(lane / LANES);
    bool active = idx < n_in;
    uint32_t game = 0;
    if (active) game = st == 0 ? idx * G + sq : __ldcg(prm.lk_list + ((size_t)(st & 1u) * G + sq) * C + idx);
    const uint8_t* seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
    auto fetch_piece = [&](uint32_t tt) -> int {                     // warp-uniform
        if (tt >= T) return 7;
        return seq ? (int)__ldg(seq + tt) : hashed_piece(prm.seed, (uint64_t)sq, tt);
    };
    uint32_t c[COLS];
    uint32_t t = lock_stage_begin(st, K), lines = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
    uint64_t gscore = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) c[k] = 0;
    if (active) {
        if (st > 0) {
            const uint4* sv = reinterpret_cast<const uint4*>(prm.st_save + game);
            const uint4 a0 = __ldcg(sv), a1 = __ldcg(sv + 1), a2 = __ldcg(sv + 2), a3 = __ldcg(sv + 3), a4 = __ldcg(sv + 4);
            c[0] = a0.x; c[1] = a0.y; c[2] = a0.z; c[3] = a0.w; c[4] = a1.x; c[5] = a1.y; c[6] = a1.z; c[7] = a1.w; c[8] = a2.x; c[9] = a2.y;
            lines = a2.w; h1 = a3.x; h2 = a3.y; h3 = a3.z; h4 = a3.w;
            if (gl == 0) cnt = a4.x;
            gscore = ((uint64_t)a4.w << 32) | a4.z;
        } else {
            if (prm.init_rows) {
                uint16_t rows[ROWS];
                for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
                rows_to_cols(rows, c);
            }
            if (prm.init_lines) lines = prm.init_lines[game];
        }
    }
    GameStat g;                                   // w6: heights / cells of the board
    PrevStat pv; Heights H;                       // general: statistics of the board the piece is placed on
    int cells0 = 0;
    const uint32_t lines0 = lines;
    double w[6];
    const double* wrow = prm.weights + (size_t)(active ? game / G : 0u) * (size_t)prm.F;
    if constexpr (W6) {
        g = game_stat(c);
        cells0 = g.cells - 4 * (int)t;
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = __ldg(wrow + i);
    } else {
        pv = prev_stat(c);
        H = pack_gh(pv.gh);
    }
    const uint32_t t_next = lock_stage_begin(st + 1u, K);
    const uint32_t t_end = (st + 1 == S || t_next > stop) ? stop : t_next;
    int p_cur = fetch_piece(t), p_next = fetch_piece(t + 1u);
    if (t >= stop || p_cur >= 7) {                                   // nothing to play (stop == 0 or the sequence ends here)
        if (active && gl == 0) {
            tb200_game_result* r = prm.results + game;
            r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4; r->score = gscore;
            if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&r->placements), (unsigned long long)cnt);
            if (prm.final_rows) { uint16_t rows[ROWS]; cols_to_rows(c, rows); for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2]; }
        }
        active = false;
    }
    while (t < t_end && p_cur < 7 && __any_sync(FULLMASK, active)) {
        const int p_after = fetch_piece(t + 2u);
        // LANES == 1: every lane is on the same slot of the same piece, so the descriptor is warp-uniform: read it from
        // constant memory with a provably uniform index (REDUX result) and let its field extraction run on the uniform
        // datapath instead of the ALU pipe, which is the binding resource of this kernel
        const int p_u = LANES == 1 ? (int)__reduce_max_sync(FULLMASK, (unsigned)p_cur) : p_cur;
        const int ns = LANES == 1 ? c_tables.nslots[p_u] : s_tab.nslots[p_cur];
        uint64_t key = 0; uint32_t tag = 0;
        for (int s2 = gl; s2 < ns; s2 += LANES) {                    // same piece, same trip count in every lane
            const SlotDesc d = LANES == 1 ? c_tables.slot[p_u][s2] : load_slot12(s_tab, p_cur, s2);
            uint64_t k = 0;
            if constexpr (W6) {
                W6Cand A;
                w6_pre(c, g, d, A);
                if (PRESS && A.legal && !is_time(d, A.P.y, prm.press, (int)(lines / 10u))) A.legal = false;      // a filtered move is never scored
                if (A.legal && A.full != 0u) w6_clear(d, A);
                Heights Hn;
                k = w6_post(A, w, Hn);
            } else {
                Placement P;
                if (place_slot(c, H, d, P) && !(PRESS && !is_time(d, P.y, prm.press, (int)(lines / 10u)))) {
                    Feat f;
                    eval_features<ModeTraits<MODE>::CMASK>(P, d, pv, prm.rmask, f);
                    k = score_key(score_features<MODE>(f, prm.ids, wrow, prm.F));
                }
            }
            cnt += (uint32_t)(k != 0ull && active);
            if (k > key) { key = k; tag = (uint32_t)s2; }
        }
        group_argmax<LANES>(key, tag);
        bool finish = false;
        if (active) {
            if (key == 0) {
                finish = true;                                       // no legal placement: game over (simulator.py:187-189)
            } else {
                const SlotDesc d = load_slot12(s_tab, p_cur, (int)tag);
                W6Cand A;
                if constexpr (W6) {
                    w6_pre(c, g, d, A);
                    if (A.full != 0u) w6_clear(d, A);
#pragma unroll
                    for (int k = 0; k < COLS; ++k) { c[k] = A.P.n[k]; g.gh[k] = A.h[k]; }
                    g.H.H0 = (uint32_t)A.h[0] | (uint32_t)A.h[1] << 8 | (uint32_t)A.h[2] << 16 | (uint32_t)A.h[3] << 24;
                    g.H.H1 = (uint32_t)A.h[4] | (uint32_t)A.h[5] << 8 | (uint32_t)A.h[6] << 16 | (uint32_t)A.h[7] << 24;
                    g.H.H2 = (uint32_t)A.h[8] | (uint32_t)A.h[9] << 8;
                } else {
                    place_slot(c, H, d, A.P);
#pragma unroll
                    for (int k = 0; k < COLS; ++k) c[k] = A.P.n[k];
                    pv = next_stat(pv, A.P, d);
                    H = pack_gh(pv.gh);
                }
                const uint32_t wk = (uint32_t)A.P.k;
                if (wk != 0u) {
                    lines += wk;
                    h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                    const uint32_t pts = prm.pts[wk];
                    gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
                } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
                if constexpr (W6) g.cells = cells0 + 4 * (int)(t + 1) - 10 * (int)(lines - lines0);
                if (gl == 0) {
                    if (prm.trace) {
                        const uint32_t row = (uint32_t)(ROWS - A.P.y - A.P.ph);
                        reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = (uint32_t)A.P.rot | row << 8 | (uint32_t)A.P.c0 << 16 | wk << 24;
                    }
                    if (prm.score_trace) {
                        const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                        prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                    }
                }
            }
        }
        const uint32_t tn = t + 1u;
        if (active && !finish && (tn >= stop || p_next >= 7)) finish = true;
        if (finish) {
            const uint32_t played = key == 0 ? t : tn;
            if (gl == 0) {
                tb200_game_result* r = prm.results + game;
                r->pieces = played; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4; r->score = gscore;
                if (prm.final_rows) { uint16_t rows[ROWS]; cols_to_rows(c, rows); for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2]; }
            }
            if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&prm.results[game].placements), (unsigned long long)cnt);
            active = false;
        }
        t = tn;
        p_cur = p_next; p_next = p_after;
    }
    // ---- survivors (alive at move t_end < stop) go to the sequence's list of the next stage ----
    {
#pragma unroll
        for (int off = LANES / 2; off > 0; off >>= 1) cnt += __shfl_xor_sync(FULLMASK, cnt, off);
        const uint32_t live = __ballot_sync(FULLMASK, active && gl == 0);
        if (live != 0u) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(prm.lk_ctl + ((size_t)(st + 1) * G + sq) * 4 + 2, (uint32_t)__popc(live));
            base = __shfl_sync(FULLMASK, base, 0);
            if (active && gl == 0) {
                uint4* sv = reinterpret_cast<uint4*>(prm.st_save + game);
                __stcg(sv, make_uint4(c[0], c[1], c[2], c[3]));
                __stcg(sv + 1, make_uint4(c[4], c[5], c[6], c[7]));
                __stcg(sv + 2, make_uint4(c[8], c[9], t, lines));
                __stcg(sv + 3, make_uint4(h1, h2, h3, h4));
                __stcg(sv + 4, make_uint4(cnt, 0u, (uint32_t)gscore, (uint32_t)(gscore >> 32)));
                const uint32_t rank = (uint32_t)__popc(live & ((1u << lane) - 1u));
                __stcg(prm.lk_list + ((size_t)((st + 1) & 1u) * G + sq) * C + base + rank, game);
            }
        }
        __syncwarp(FULLMASK);
        if (lane == 0) { __threadfence(); atomicAdd(ctl + 1, 1u); }
    }
    }
}

template <int LANES, int MODE = MODE_W6>
__global__ void __launch_bounds__(BLOCK) play_games_w6lock_kernel(const PlayParams prm)
{
    constexpr uint32_t GPW = 32 / LANES;           // games per warp
    __shared__ Tables s_tab;
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t C = prm.lk_C, S = prm.lk_S;
    const uint32_t G = (uint32_t)prm.G;

    uint32_t st = 0, sq = 0;                       // current bucket
    uint32_t n_in = C, ntasks = (C + GPW - 1) / GPW;
    for (;;) {
        uint32_t* ctl = prm.lk_ctl + ((size_t)st * G + sq) * 4;
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(ctl, 1u);
        slot = __shfl_sync(FULLMASK, slot, 0);
        if (slot >= ntasks) {                      // bucket handed out: open the next one
            if (++sq == G) { sq = 0; if (++st == S) break; }
            if (st > 0) {
                if (lane == 0) {
                    const uint32_t* prev = prm.lk_ctl + ((size_t)(st - 1) * G + sq) * 4;
                    const uint32_t n_prev = st == 1 ? C : __ldcg(prev + 2);
                    const uint32_t need = (n_prev + GPW - 1) / GPW;
                    uint32_t spins = 0;
                    while (*reinterpret_cast<volatile const uint32_t*>(prev + 1) < need) {
                        __nanosleep(200);
                        if (++spins > (1u << 23)) { atomicExch(prm.lk_ctl + (size_t)S * G * 4 + 3, 1u); break; }   // safety valve (> 1.5 s): never hang the device
                    }
                    __threadfence();
                    n_in = spins > (1u << 23) ? 0u : __ldcg(prm.lk_ctl + ((size_t)st * G + sq) * 4 + 2);
                }
                n_in = __shfl_sync(FULLMASK, n_in, 0);
            }
            ntasks = (n_in + GPW - 1) / GPW;
            continue;
        }
        lock_task<LANES, MODE>(prm, s_tab, st, sq, slot, n_in, ctl);
    }
}

// ---------------------------------------------------------------------------------------------
// K2sW: the lockstep kernel with WINDOW-LOCAL evaluation (one lane per game, w6, lookahead 1).  Same task / bucket / stage
// machinery as K2s; what changes is the inner loop.  Every lane of a warp is on the same slot of the same piece, so the
// slot's first column c0 and width are warp-uniform: a placement is scored from the window of columns it touches
// (tetris_core.cuh: win_eval) — ~9 popc and ~140 instructions instead of ~19 popc and ~260 for all ten columns.
// The boards live in shared memory, word-major ([word][lane]: every lane reads and writes only its own bank, no
// conflicts, no synchronisation), so that a uniform run-time column index is a plain LDS; per game there are the ten
// columns and the prefix / suffix ANDs of the columns (full-row test of a window placement = two loads and one AND).
// The per-pair / per-column contributions of the current board and their totals (WinCache, 18 registers) are rebuilt at
// every commit.  Placements that complete a row move every column: they are DEFERRED (one bit per slot) and evaluated
// after the slot loop with the general w6 code, so the common path carries no divergent clear branch at all.
// ---------------------------------------------------------------------------------------------
constexpr int LW_BOARD = 30;           // per game: columns 0..9 | P[k] = AND of columns < k, k = 0..9 | S[k] = AND of columns >= k, k = 1..10
constexpr int LW_ITEMS = 256;          // deferred (game, slot) placements a warp shares per piece (more than that: their owners evaluate them)
constexpr int LW_XCH = LW_BOARD + LW_ITEMS / 32 + 2 * LW_ITEMS / 32;        // K2sP: this warp's best (key lo, key hi, slot) per game, double-buffered by piece parity
constexpr int LW_WORDS = LW_XCH + 6;                                        // board + item table (owner | slot << 8) + 64-bit keys of the items + exchange rows

// K2sP — a group of warps of one block playing the same 32 games (see lockwin_task<.., SPLIT>): they meet once per piece at
// an mbarrier in shared memory.  The wait is bounded (try_wait + %globaltimer): a warp whose partner never arrives raises the
// launch's error flag and leaves the kernel instead of hanging the device.
struct PairLink {
    uint32_t bar;                  // shared-memory address of the group's mbarrier (arrival count = nparts: lane 0 of every warp)
    uint32_t phase;                // parity of the phase this warp waits for next
    int part;                      // this warp's index in the group (0 = the one that writes results); warp-uniform BY CONSTRUCTION (a vote result),
                                   // so that the slot ranges derived from it keep the descriptor loads on the uniform datapath
    uint32_t (*peer)[32];          // the partner warp's shared-memory rows
    uint32_t par;                  // exchange-row parity
};
constexpr int PAIR_PARTS = 2;
// Everything this warp wrote to shared memory before the call is visible to the partner after its matching call returns
// (__syncwarp orders the lanes' writes before lane 0's arrive.release; lane 0's try_wait.acquire + the shuffle order the partner's
// writes before every lane's later reads).  The result is a shuffle from lane 0: uniform, also in the compiler's eyes.
TB_D bool pair_sync(uint32_t bar, uint32_t& phase)
{
    __syncwarp(FULLMASK);
    uint32_t ok = 0;
    if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
        unsigned long long t0 = 0ull;
        for (;;) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
            if (ok) break;
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0ull) t0 = now; else if (now - t0 > 4000000000ull) break;      // 4 s: longer than every other bounded wait of the kernel
        }
    }
    ok = __shfl_sync(FULLMASK, ok, 0);
    phase ^= 1u;
    return ok != 0u;
}

TB_D void lw_store_board(uint32_t (*sb)[32], int lane, const uint32_t c[COLS])
{
    uint32_t acc = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < COLS; ++k) { sb[k][lane] = c[k]; sb[10 + k][lane] = acc; acc &= c[k]; }
    acc = 0xffffffffu;
#pragma unroll
    for (int k = COLS; k >= 1; --k) { sb[20 + k - 1][lane] = acc; acc &= c[k - 1]; }
}

// Best placement of piece p_cur (warp-uniform) for the game of this lane, window-local evaluation + deferred row-completing
// placements; key = 0: no legal placement.  cnt counts the legal placements of active lanes (= scorer calls).
// The slots cc .. ce - 1 of one rotation (piece width WDT, warp-uniform), windows sliding one column per slot.
template <int WDT, bool PRESS>
TB_D void lw_slots(uint32_t (*sb)[32], int lane, const WinCache& wc, WinSlide& ws, const double* w, int p_u, int& s2, int cc, int ce, bool active,
                   const Pressure& press, int speed, uint32_t& cnt, uint64_t& key, uint32_t& tag, uint32_t& dlo, uint32_t& dhi)
{
#pragma unroll 1
    for (; cc < ce; ++cc, ++s2) {
        const SlotDesc d = c_tables.slot[p_u][s2];
        const WinDesc wdsc = c_win[p_u][s2];
        const int c0 = (int)(d.w[0] & 15u);
        const uint32_t and_out = sb[10 + c0][lane] & sb[20 + c0 + WDT - 1][lane];
        uint32_t full; bool legal; int y;
        uint64_t k = win_eval_w<WDT>(wc, ws, d, wdsc, w, [&](int kk) -> uint32_t { return sb[kk][lane]; }, and_out, &full, &legal, &y);
        win_advance(ws);
        if (PRESS && legal && !is_time_at_speed(d, y, press, speed)) { legal = false; k = 0ull; }      // move_with_pressure: filtered before it is scored
        cnt += (uint32_t)(legal && active);
        if (legal && full != 0u) { k = 0ull; if (s2 < 32) dlo |= 1u << s2; else dhi |= 1u << (s2 - 32); }
        if (k > key) { key = k; tag = (uint32_t)s2; }
    }
}

// SPLIT: the slots of the piece are shared out between the `nparts` (1 or 2) warps that play these 32 games together (K2sP, see
// lockwin_task): this warp takes half (part + rotation / 2) mod 2 of every rotation's columns, so the parts stay balanced.
template <bool PRESS, bool SPLIT = false>
TB_D void lw_search(const Tables& s_tab, uint32_t (*sb)[32], int lane, const WinCache& wc, const double* w, int p_cur, bool active,
                    const Pressure& press, int level, uint32_t& cnt, uint64_t& key, uint32_t& tag, const int nparts = 1, const int part = 0)
{
    uint32_t c[COLS];
    const int p_u = (int)__reduce_max_sync(FULLMASK, (unsigned)p_cur);      // provably uniform: descriptors from constant memory, fields on the uniform datapath
    const int ns = c_tables.nslots[p_u];
    uint32_t dlo = 0u, dhi = 0u;                                 // slots deferred to the general evaluation (they complete a row)
    const int speed = PRESS ? speed_of_level(level) : 0;         // the level bracket of simulator.py:110-114, once per piece
    // slots are enumerated rotation-major, column-minor: per rotation the windows start at column 0 and slide
    for (int s2 = 0, rot = 0; s2 < ns; ++rot) {
        WinSlide ws;
        win_reset(wc, ws);
        const int ncol = COLS + 1 - (int)((c_tables.slot[p_u][s2].w[0] >> 6) & 7u);      // columns of this rotation (uniform)
        int cc = 0, ce = ncol;
        const int s2_end = s2 + ncol;
        if (SPLIT) {
            const int q = (part + (rot >> 1)) & (nparts - 1);      // nparts is 1 (the whole rotation) or 2; all warp-uniform
            cc = (q * ncol) >> (nparts - 1); ce = ((q + 1) * ncol) >> (nparts - 1);
            win_skip(ws, cc);
            s2 += cc;
        }
        switch (COLS + 1 - ncol) {                 // the rotation's piece width: uniform, so one jump per rotation buys a slot loop without width tests
        case 1: lw_slots<1, PRESS>(sb, lane, wc, ws, w, p_u, s2, cc, ce, active, press, speed, cnt, key, tag, dlo, dhi); break;
        case 2: lw_slots<2, PRESS>(sb, lane, wc, ws, w, p_u, s2, cc, ce, active, press, speed, cnt, key, tag, dlo, dhi); break;
        case 3: lw_slots<3, PRESS>(sb, lane, wc, ws, w, p_u, s2, cc, ce, active, press, speed, cnt, key, tag, dlo, dhi); break;
        default: lw_slots<4, PRESS>(sb, lane, wc, ws, w, p_u, s2, cc, ce, active, press, speed, cnt, key, tag, dlo, dhi); break;
        }
        s2 = s2_end;
    }
    // ---- deferred placements: rows completed, every column moves -> general w6 evaluation.  The lanes of a warp hold very
    //      different numbers of them (0 .. 10 per piece), so they are SHARED: every lane lists its (game, slot) items in shared
    //      memory, the warp evaluates 32 items per round whatever game they belong to (the board comes from the owner's
    //      shared-memory column, the weights / heights / cell count by shuffle from the owner's registers), and every owner
    //      collects the keys of its own items in slot order. ----
    const uint32_t mine = (uint32_t)__popc(dlo) + (uint32_t)__popc(dhi);
    if (!__any_sync(FULLMASK, mine != 0u)) return;
    uint32_t off = mine;                                         // inclusive prefix sum over the lanes
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) { const uint32_t v = __shfl_up_sync(FULLMASK, off, dlt); if (lane >= dlt) off += v; }
    const uint32_t total = __shfl_sync(FULLMASK, off, 31);
    off -= mine;
    uint32_t* const items = &sb[LW_BOARD][0];
    unsigned long long* const keys = reinterpret_cast<unsigned long long*>(&sb[LW_BOARD + LW_ITEMS / 32][0]);
    {
        uint32_t a = dlo, b = dhi, i = off;
        while ((a | b) != 0u) {
            int s2;
            if (a != 0u) { s2 = ctz(a); a &= a - 1u; } else { s2 = 32 + ctz(b); b &= b - 1u; }
            if (i < (uint32_t)LW_ITEMS) items[i] = (uint32_t)lane | (uint32_t)s2 << 8;
            ++i;
        }
    }
    __syncwarp(FULLMASK);
    const uint32_t shared_n = total < (uint32_t)LW_ITEMS ? total : (uint32_t)LW_ITEMS;
    for (uint32_t base = 0; base < shared_n; base += 32u) {
        const uint32_t j = base + (uint32_t)lane;
        const bool have = j < shared_n;
        const uint32_t it = have ? items[j] : 0u;
        const int owner = (int)(it & 31u), s2 = (int)(it >> 8);
        double wo[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) wo[i] = __shfl_sync(FULLMASK, w[i], owner);
        const uint32_t x0 = __shfl_sync(FULLMASK, wc.X[0], owner), x1 = __shfl_sync(FULLMASK, wc.X[1], owner),
                       x2 = __shfl_sync(FULLMASK, wc.X[2], owner), x3 = __shfl_sync(FULLMASK, wc.X[3], owner);
        GameStat g;
        g.H.H0 = funnel_r(x0, x1, 16u); g.H.H1 = funnel_r(x1, x2, 16u); g.H.H2 = funnel_r(x2, x3, 16u) & 0xffffu;
        unpack_heights(g.H, g.gh);
        g.cells = __shfl_sync(FULLMASK, wc.cells, owner);
        if (have) {
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = sb[k][owner];
            const SlotDesc d = load_slot12(s_tab, p_cur, s2);
            W6Cand A;
            w6_pre(c, g, d, A);
            w6_clear(d, A);
            Heights Hn;
            keys[j] = w6_post(A, wo, Hn);
        }
    }
    __syncwarp(FULLMASK);
    if (mine != 0u) {
        GameStat g;
        bool loaded = false;
        uint32_t i = off;
        while ((dlo | dhi) != 0u) {
            int s2;
            if (dlo != 0u) { s2 = ctz(dlo); dlo &= dlo - 1u; } else { s2 = 32 + ctz(dhi); dhi &= dhi - 1u; }
            uint64_t k;
            if (i < (uint32_t)LW_ITEMS) {
                k = keys[i];
            } else {                                             // beyond the shared table: the owner evaluates it itself
                if (!loaded) {
                    g.H.H0 = funnel_r(wc.X[0], wc.X[1], 16u); g.H.H1 = funnel_r(wc.X[1], wc.X[2], 16u); g.H.H2 = funnel_r(wc.X[2], wc.X[3], 16u) & 0xffffu;
                    unpack_heights(g.H, g.gh);
                    g.cells = wc.cells;
#pragma unroll
                    for (int kk = 0; kk < COLS; ++kk) c[kk] = sb[kk][lane];
                    loaded = true;
                }
                const SlotDesc d = load_slot12(s_tab, p_cur, s2);
                W6Cand A;
                w6_pre(c, g, d, A);
                w6_clear(d, A);
                Heights Hn;
                k = w6_post(A, w, Hn);
            }
            ++i;
            if (k > key || (k == key && (uint32_t)s2 < tag)) { key = k; tag = (uint32_t)s2; }       // first of max: the lower slot keeps a tie
        }
    }
    __syncwarp(FULLMASK);                                        // the item / key rows are reused by the next piece
}

// Commit the winning slot `tag`: apply the placement to the lane's board (shared memory), rebuild the window cache.
TB_D void lw_commit(const Tables& s_tab, uint32_t (*sb)[32], int lane, WinCache& wc, int p_cur, uint32_t tag, W6Cand& A)
{
    uint32_t c[COLS];
    const SlotDesc d = load_slot12(s_tab, p_cur, (int)tag);
    GameStat g;
    g.H.H0 = funnel_r(wc.X[0], wc.X[1], 16u); g.H.H1 = funnel_r(wc.X[1], wc.X[2], 16u); g.H.H2 = funnel_r(wc.X[2], wc.X[3], 16u) & 0xffffu;
    unpack_heights(g.H, g.gh);
    g.cells = wc.cells;
#pragma unroll
    for (int k = 0; k < COLS; ++k) c[k] = sb[k][lane];
    w6_pre(c, g, d, A);
    if (A.full != 0u) w6_clear(d, A);
#pragma unroll
    for (int k = 0; k < COLS; ++k) c[k] = A.P.n[k];
    win_cache_init(c, wc);
    lw_store_board(sb, lane, c);
}

// The window-local task: 32 games of sequence sq, one lane per game; sb = this warp's board words in shared memory.
// SPLIT (K2sP, round 2): TWO warps play the task together.  Both hold the complete state of the 32 games (own board rows, own
// window cache) and do everything identically, except that each searches only its share of the piece's slots (lw_search);
// per piece they swap their best (key, slot) per game through shared memory — one mbarrier phase — and then both commit the
// same winner (higher key, then lower slot: first of max).  Only part 0 writes results / traces / hand-over state; the
// scorer-call count is additive, so each part adds its own.  For shards that fill less than the machine with one warp per 32
// games (a rank's 1 250 x 30 slice: 1 200 tasks on 2 368 resident warps) this halves the per-piece latency of the search, which
// is all that matters there.  Returns false only when the partner never arrived (bounded wait, error flag raised).
// (SPLIT = the kernel has the pair task at all; `split` = this task is one, warp-uniform.  ONE instantiation serves both: a second
//  inlined copy of this function pushes the kernel past the size up to which ptxas keeps the slot loop on the uniform datapath —
//  measured: the plain one-lane path lost 40 %.)
template <bool PRESS = false, bool SPLIT = false>
TB_D bool lockwin_task(const PlayParams& prm, const Tables& s_tab, uint32_t (*sb)[32], uint32_t st, uint32_t sq, uint32_t slot, uint32_t n_in, uint32_t* ctl,
                       bool split = false, PairLink* pl = nullptr)
{
    const bool lead = !SPLIT || !split || pl->part == 0;
    constexpr uint32_t GPW = 32;
    const int lane = threadIdx.x & 31;
    const uint32_t T = prm.T, C = prm.lk_C, K = prm.lk_K, S = prm.lk_S;
    const uint32_t G = (uint32_t)prm.G;
    const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
    {
    // ---- task (st, sq, slot): up to 32 games of sequence sq, all at the first move of stage st ----
    const uint32_t idx = slot * GPW + (uint32_t)lane;
    bool active = idx < n_in;
    uint32_t game = 0;
    if (active) game = st == 0 ? idx * G + sq : __ldcg(prm.lk_list + ((size_t)(st & 1u) * G + sq) * C + idx);
    const uint8_t* seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
    auto fetch_piece = [&](uint32_t tt) -> int {                     // warp-uniform
        if (tt >= T) return 7;
        return seq ? (int)__ldg(seq + tt) : hashed_piece(prm.seed, (uint64_t)sq, tt);
    };
    uint32_t c[COLS];
    uint32_t t = lock_stage_begin(st, K), lines = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
    uint64_t gscore = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) c[k] = 0;
    if (active) {
        if (st > 0) {
            const uint4* sv = reinterpret_cast<const uint4*>(prm.st_save + game);
            const uint4 a0 = __ldcg(sv), a1 = __ldcg(sv + 1), a2 = __ldcg(sv + 2), a3 = __ldcg(sv + 3), a4 = __ldcg(sv + 4);
            c[0] = a0.x; c[1] = a0.y; c[2] = a0.z; c[3] = a0.w; c[4] = a1.x; c[5] = a1.y; c[6] = a1.z; c[7] = a1.w; c[8] = a2.x; c[9] = a2.y;
            lines = a2.w; h1 = a3.x; h2 = a3.y; h3 = a3.z; h4 = a3.w;
            cnt = lead ? a4.x : 0u;
            gscore = ((uint64_t)a4.w << 32) | a4.z;
        } else {
            if (prm.init_rows) {
                uint16_t rows[ROWS];
                for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
                rows_to_cols(rows, c);
            }
            if (prm.init_lines) lines = prm.init_lines[game];
        }
    }
    WinCache wc;
    win_cache_init(c, wc);
    lw_store_board(sb, lane, c);
    double w[6];
    {
        const double* wrow = prm.weights + (size_t)(active ? game / G : 0u) * 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = __ldg(wrow + i);
    }
    const uint32_t t_next = lock_stage_begin(st + 1u, K);
    const uint32_t t_end = (st + 1 == S || t_next > stop) ? stop : t_next;
    int p_cur = fetch_piece(t), p_next = fetch_piece(t + 1u);
    if (t >= stop || p_cur >= 7) {                                   // nothing to play (stop == 0 or the sequence ends here)
        if (active && lead) {
            tb200_game_result* r = prm.results + game;
            r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4; r->score = gscore;
            if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&r->placements), (unsigned long long)cnt);
            if (prm.final_rows) { uint16_t rows[ROWS]; cols_to_rows(c, rows); for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2]; }
        }
        active = false;
    }
    while (t < t_end && p_cur < 7 && __any_sync(FULLMASK, active)) {
        const int p_after = fetch_piece(t + 2u);
        uint64_t key = 0; uint32_t tag = 0;
        if constexpr (SPLIT) {
            lw_search<PRESS, true>(s_tab, sb, lane, wc, w, p_cur, active, prm.press, (int)(lines / 10u), cnt, key, tag, split ? PAIR_PARTS : 1, pl->part);
          if (split) {
            const uint32_t xr = (uint32_t)LW_XCH + 3u * pl->par;
            sb[xr][lane] = (uint32_t)key; sb[xr + 1][lane] = (uint32_t)(key >> 32); sb[xr + 2][lane] = tag;
            if (!pair_sync(pl->bar, pl->phase)) {
                if (lane == 0) atomicExch(prm.lk_ctl + (size_t)S * G * 4 + 3, 1u);
                return false;
            }
            const uint64_t okey = ((uint64_t)pl->peer[xr + 1][lane] << 32) | pl->peer[xr][lane];
            const uint32_t otag = pl->peer[xr + 2][lane];
            if (okey > key || (okey == key && otag < tag)) { key = okey; tag = otag; }      // first of max: the lower slot keeps a tie
            pl->par ^= 1u;
          }
        } else {
            lw_search<PRESS>(s_tab, sb, lane, wc, w, p_cur, active, prm.press, (int)(lines / 10u), cnt, key, tag);
        }
        bool finish = false;
        if (active) {
            if (key == 0) {
                finish = true;                                       // no legal placement: game over (simulator.py:187-189)
            } else {
                W6Cand A;
                lw_commit(s_tab, sb, lane, wc, p_cur, tag, A);
                const uint32_t wk = (uint32_t)A.P.k;
                if (wk != 0u) {
                    lines += wk;
                    h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                    const uint32_t pts = prm.pts[wk];
                    gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
                } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
                if (prm.trace && lead) {
                    const uint32_t row = (uint32_t)(ROWS - A.P.y - A.P.ph);
                    reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = (uint32_t)A.P.rot | row << 8 | (uint32_t)A.P.c0 << 16 | wk << 24;
                }
                if (prm.score_trace && lead) {
                    const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                    prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                }
            }
        }
        const uint32_t tn = t + 1u;
        if (active && !finish && (tn >= stop || p_next >= 7)) finish = true;
        if (finish) {
            const uint32_t played = key == 0 ? t : tn;
            tb200_game_result* r = prm.results + game;
            if (lead) { r->pieces = played; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4; r->score = gscore; }
            if (prm.final_rows && lead) {
                uint16_t rows[ROWS];
#pragma unroll
                for (int k = 0; k < COLS; ++k) c[k] = sb[k][lane];
                cols_to_rows(c, rows);
                for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2];
            }
            if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&prm.results[game].placements), (unsigned long long)cnt);
            active = false;
        }
        t = tn;
        p_cur = p_next; p_next = p_after;
    }
    // ---- survivors (alive at move t_end < stop) go to the sequence's list of the next stage ----
    {
        const uint32_t live = __ballot_sync(FULLMASK, active);
        if (!lead) {                                   // the other part hands the games over; this part's scorer calls are added where they end up
            if (active && cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&prm.results[game].placements), (unsigned long long)cnt);
        } else if (live != 0u) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(prm.lk_ctl + ((size_t)(st + 1) * G + sq) * 4 + 2, (uint32_t)__popc(live));
            base = __shfl_sync(FULLMASK, base, 0);
            if (active) {
#pragma unroll
                for (int k = 0; k < COLS; ++k) c[k] = sb[k][lane];
                uint4* sv = reinterpret_cast<uint4*>(prm.st_save + game);
                __stcg(sv, make_uint4(c[0], c[1], c[2], c[3]));
                __stcg(sv + 1, make_uint4(c[4], c[5], c[6], c[7]));
                __stcg(sv + 2, make_uint4(c[8], c[9], t, lines));
                __stcg(sv + 3, make_uint4(h1, h2, h3, h4));
                __stcg(sv + 4, make_uint4(cnt, 0u, (uint32_t)gscore, (uint32_t)(gscore >> 32)));
                const uint32_t rank = (uint32_t)__popc(live & ((1u << lane) - 1u));
                __stcg(prm.lk_list + ((size_t)((st + 1) & 1u) * G + sq) * C + base + rank, game);
            }
        }
        __syncwarp(FULLMASK);
        if (lane == 0 && lead) { __threadfence(); atomicAdd(ctl + 1, 1u); }
    }
    }
    return true;
}

__global__ void __launch_bounds__(BLOCK, 4) play_games_w6lockwin_kernel(const PlayParams prm)
{
    constexpr uint32_t GPW = 32;                   // games per warp
    __shared__ Tables s_tab;
    __shared__ alignas(16) uint32_t s_board[BLOCK / 32][LW_WORDS][32];
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t (*sb)[32] = s_board[threadIdx.x >> 5];
    const uint32_t C = prm.lk_C, S = prm.lk_S;
    const uint32_t G = (uint32_t)prm.G;

    uint32_t st = 0, sq = 0;                       // current bucket
    uint32_t n_in = C, ntasks = (C + GPW - 1) / GPW;
    for (;;) {
        uint32_t* ctl = prm.lk_ctl + ((size_t)st * G + sq) * 4;
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(ctl, 1u);
        slot = __shfl_sync(FULLMASK, slot, 0);
        if (slot >= ntasks) {                      // bucket handed out: open the next one
            if (++sq == G) { sq = 0; if (++st == S) break; }
            if (st > 0) {
                if (lane == 0) {
                    const uint32_t* prev = prm.lk_ctl + ((size_t)(st - 1) * G + sq) * 4;
                    const uint32_t n_prev = st == 1 ? C : __ldcg(prev + 2);
                    const uint32_t need = (n_prev + GPW - 1) / GPW;
                    uint32_t spins = 0;
                    while (*reinterpret_cast<volatile const uint32_t*>(prev + 1) < need) {
                        __nanosleep(200);
                        if (++spins > (1u << 23)) { atomicExch(prm.lk_ctl + (size_t)S * G * 4 + 3, 1u); break; }   // safety valve (> 1.5 s): never hang the device
                    }
                    __threadfence();
                    n_in = spins > (1u << 23) ? 0u : __ldcg(prm.lk_ctl + ((size_t)st * G + sq) * 4 + 2);
                }
                n_in = __shfl_sync(FULLMASK, n_in, 0);
            }
            ntasks = (n_in + GPW - 1) / GPW;
            continue;
        }
        lockwin_task(prm, s_tab, sb, st, sq, slot, n_in, ctl);
    }
}

// ---------------------------------------------------------------------------------------------
// K2sA: lanes per game chosen PER BUCKET from the live count ("choose lanes per game per stage").  A CE population
// thins out as it plays — a fresh one loses most of its games within a few dozen pieces — and a rank's shard of a sharded
// population is small to begin with, so one launch-wide width is wrong for part of every play: with few games alive one
// lane per game leaves most warps without work (and the rest latency-bound), with many alive eight lanes per game waste
// lanes.  Here every (stage, sequence) bucket picks its own width from the number of games that enter it, n_in (every
// warp reads the same counter once the bucket is open, so the choice — and hence the task count — is the same everywhere):
//     games alive over all G sequences per resident warp  >= 24 : one lane per game, window-local evaluation (K2sW task)
//                                >= 8 : four lanes per game     >= 2 : eight lanes per game     else : a whole warp per game
// (prm.lk_thr; measured, tools/auto_probe.py: with many games per warp the lane utilisation decides, with few the per-piece latency)
// Tasks of different widths are the same device functions the fixed-width kernels run; results do not depend on the width.
// ---------------------------------------------------------------------------------------------
// width code of a bucket: 1 = one lane per game (K2sW task), 2 = one lane per game in each of TWO warps that share the slots
// (K2sP task), 4 / 8 / 32 = lanes per game of the all-columns task
// n_prev = games that entered the PREVIOUS stage of the same sequence (0: unknown, first stage).  A population that loses almost nobody
// per stage is STEADY: its games will still be there hundreds of pieces later, so the cheapest evaluation per placement (one lane per
// game, window-local) wins down to far fewer games per warp (lk_thr[4]) than for a population that is still thinning out, where the few
// long-lived games are better served by wide tasks (measured on one rank's 1 250 x 30 slice, tools/pair_probe.py: converged 11.3 ms with one
// lane per game against 12.6 with four; fresh 11.0 against 6.0).
TB_D uint32_t lock_auto_lanes(const PlayParams& prm, uint32_t n_in, uint32_t n_prev, uint32_t G, uint32_t warps)
{
    const uint64_t alive = (uint64_t)n_in * G;                            // estimate: the other sequences thin out alike
    const bool steady = n_prev != 0u && (uint64_t)n_in * 16u >= (uint64_t)n_prev * 15u;
    const uint32_t thr1 = steady && prm.lk_thr[4] < prm.lk_thr[0] ? prm.lk_thr[4] : prm.lk_thr[0];
    return alive >= (uint64_t)thr1 * warps ? 1u : (alive >= (uint64_t)prm.lk_thr[3] * warps ? 2u :
           (alive >= (uint64_t)prm.lk_thr[1] * warps ? 4u : (alive >= (uint64_t)prm.lk_thr[2] * warps ? 8u : 32u)));
}
TB_D uint32_t lock_auto_tasks(uint32_t n_in, uint32_t lanes) { return lanes == 2u ? (n_in + 31u) / 32u : (n_in * lanes + 31u) / 32u; }

template <int MINB, bool PRESS = false, bool PAIR = false>      // PAIR: with the K2sP task (host: lk_thr[3] < lk_thr[0] only for this instantiation)
__global__ void __launch_bounds__(BLOCK, MINB) play_games_w6lockauto_kernel(const PlayParams prm)
{
    __shared__ Tables s_tab;
    __shared__ alignas(16) uint32_t s_board[BLOCK / 32][LW_WORDS][32];
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    __shared__ alignas(8) unsigned long long s_bar[BLOCK / 64];      // K2sP: one mbarrier per pair of warps
    __shared__ uint32_t s_slot[BLOCK / 64][2];                        // K2sP: the pair's ticket, double-buffered
    const int lane = threadIdx.x & 31;
    const uint32_t wib = threadIdx.x >> 5;
    uint32_t (*sb)[32] = s_board[wib];
    const uint32_t C = prm.lk_C, S = prm.lk_S, G = (uint32_t)prm.G;
    const uint32_t warps = gridDim.x * (BLOCK / 32);
    PairLink pl;
    pl.bar = (uint32_t)__cvta_generic_to_shared(&s_bar[wib >> 1]); pl.phase = 0u; pl.part = (int)__reduce_max_sync(FULLMASK, wib & 1u); pl.peer = s_board[wib ^ 1u]; pl.par = 0u;
    if (PAIR && threadIdx.x < BLOCK / 64) asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_bar[threadIdx.x])) : "memory");
    __syncthreads();
    uint32_t sbuf = 0u;

    // A stage has ntasks x G tasks.  When that is less than the resident warps, only about as many warps as there are tasks take
    // tickets, and they are chosen so that they spread evenly over the SMs AND over the four warp schedulers of every SM — otherwise
    // the first warps to arrive pile the tasks onto a few SMs / schedulers while the others idle.  Block placement is not something
    // a kernel may assume (round 2 first ranked warps by blockIdx, taking blocks b, b + SMs, ... to share an SM: a 1 250 x 30 slice
    // played with one lane per game then took 20.4 ms against 11.3 ms for the exact-grid kernel), so every warp finds out where it
    // runs: k = its arrival ordinal on (its SM, its scheduler = %warpid mod 4), from per-launch counters behind the bucket table;
    // j = 4 k + scheduler orders the warp slots of an SM scheduler-minor, and a bucket is served by the slots j < floor(tasks / SMs).
    uint32_t slot_j;
    {
        uint32_t smid, wid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        const uint32_t sch = wid & 3u;
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(prm.lk_ctl + (size_t)(S + 1) * G * 4 + (size_t)(smid & 255u) * 4 + sch, 1u);
        k = __shfl_sync(FULLMASK, k, 0);
        slot_j = k * 4u + sch;
    }
    const uint32_t sms = gridDim.x / (uint32_t)MINB ? gridDim.x / (uint32_t)MINB : 1u;
    // floor, not ceil: one warp too many on an SM puts three warps on a scheduler that should hold two, and every sequence then waits for
    // such a warp at every stage boundary; one too few only makes the last tasks of a stage start a little late
    auto takes = [&](uint32_t tasks_in_stage) -> bool { return tasks_in_stage >= sms ? (uint64_t)(slot_j + 1u) * sms <= (uint64_t)tasks_in_stage : slot_j == 0u; };
    const uint32_t prank = (((wib >> 1) + blockIdx.x / sms) & 1u) * gridDim.x + blockIdx.x;      // K2sP (experiment): pairs of warps ranked by block index
    uint32_t st = 0, sq = 0;                       // current bucket
    uint32_t n_in = C;
    const uint32_t lanes0 = lock_auto_lanes(prm, C, 0u, G, warps);      // every bucket of the first stage
    uint32_t lanes = lanes0;
    uint32_t ntasks = lock_auto_tasks(n_in, lanes);
    bool mine = takes(ntasks * G);                 // this warp's slot serves the current bucket
    bool take = lanes == 2u ? (uint64_t)prank < (uint64_t)ntasks * G : mine;
    for (;;) {
        uint32_t* ctl = prm.lk_ctl + ((size_t)st * G + sq) * 4;
        uint32_t slot = 0xffffffffu;
        if (PAIR && lanes == 2u) {                 // a pair bucket: part 0 takes the ticket for both (both warps walk the same buckets and
            if (pl.part == 0) {                    // read the same final n_in, so both know this is a pair bucket)
                if (take && lane == 0) slot = atomicAdd(ctl, 1u);
                if (lane == 0) s_slot[wib >> 1][sbuf] = slot;
            }
            if (!pair_sync(pl.bar, pl.phase)) { if (lane == 0) atomicExch(prm.lk_ctl + (size_t)S * G * 4 + 3, 1u); return; }
            if (lane == 0) slot = s_slot[wib >> 1][sbuf];
            slot = __shfl_sync(FULLMASK, slot, 0);      // uniform in the compiler's eyes too (the tasks rely on the uniform datapath)
            sbuf ^= 1u;
        } else if (take) {
            if (lane == 0) slot = atomicAdd(ctl, 1u);
            slot = __shfl_sync(FULLMASK, slot, 0);
        }
        if (slot >= ntasks) {                      // bucket handed out (or not mine): open the next one
            if (++sq == G) { sq = 0; if (++st == S) break; }
            // One bucket = one 16-byte record {ticket, done, games in, -}: it is read with ONE 128-bit load, and the ticket word of the snapshot
            // tells a warp whether there is anything left to take before it spends an atomic on it.  (When a shard fills the machine only
            // once, every warp passes every bucket and most find it handed out: ncu of a 2 500 x 30 slice had 13.5 % of the warp time
            // waiting on these L2 round trips — 2 368 atomics queueing on one word per bucket — and 10.5 % asleep.)
            uint32_t n_prev = 0u, tk = 0u;
            if (st == 0) {
                if (lane == 0) tk = __ldcg(prm.lk_ctl + ((size_t)st * G + sq) * 4);
                tk = __shfl_sync(FULLMASK, tk, 0);
            } else {
                if (lane == 0) {
                    const uint32_t* prev = prm.lk_ctl + ((size_t)(st - 1) * G + sq) * 4;
                    const uint4 pr = __ldcg(reinterpret_cast<const uint4*>(prev));
                    n_prev = st == 1 ? C : pr.z;
                    // width the previous bucket of this sequence was played with: a function of its own count and of the count one stage earlier
                    // (both final since this warp passed there; the two loads are independent)
                    const uint32_t n_pp = st == 1 ? 0u : (st == 2 ? C : __ldcg(prev - (size_t)G * 4 + 2));
                    const uint32_t need = lock_auto_tasks(n_prev, lock_auto_lanes(prm, n_prev, n_pp, G, warps));      // tasks of the previous bucket of this sequence
                    // Warps that take no tickets at the moment (fewer tasks than warps) all wait at the first bucket that is not open yet: a thousand
                    // warps polling one L2 line every 200 ns starve the counters on that line that the finishing tasks must bump (measured: a
                    // 1 250 x 30 slice with one lane per game took 20.6 ms inside this kernel against 11.3 ms with an exact grid).  They poll 32 x slower.
                    const uint32_t nap = mine ? 200u : 6400u;
                    uint64_t waited = 0;
                    uint32_t done = pr.y;
                    while (done < need) {
                        __nanosleep(nap);
                        waited += nap;
                        if (waited > (1ull << 31)) { atomicExch(prm.lk_ctl + (size_t)S * G * 4 + 3, 1u); break; }   // safety valve (> 2 s of naps): never hang the device
                        done = *reinterpret_cast<volatile const uint32_t*>(prev + 1);
                    }
                    __threadfence();
                    const uint4 cu = __ldcg(reinterpret_cast<const uint4*>(prm.lk_ctl + ((size_t)st * G + sq) * 4));      // after the fence: `games in` is final
                    n_in = waited > (1ull << 31) ? 0u : cu.z;
                    tk = cu.x;
                }
                n_in = __shfl_sync(FULLMASK, n_in, 0);
                n_prev = __shfl_sync(FULLMASK, n_prev, 0);
                tk = __shfl_sync(FULLMASK, tk, 0);
            }
            lanes = lock_auto_lanes(prm, n_in, n_prev, G, warps);
            ntasks = lock_auto_tasks(n_in, lanes);
            mine = takes(ntasks * G);
            take = lanes == 2u ? (uint64_t)prank < (uint64_t)ntasks * G : (mine && tk < ntasks);      // tickets only grow: a snapshot >= ntasks = handed out
            continue;
        }
        if (lanes <= 2u) { if (!lockwin_task<PRESS, PAIR>(prm, s_tab, sb, st, sq, slot, n_in, ctl, PAIR && lanes == 2u, &pl)) return; }
        else if (lanes == 4u) lock_task<4, MODE_W6, PRESS>(prm, s_tab, st, sq, slot, n_in, ctl);
        else if (lanes == 8u) lock_task<8, MODE_W6, PRESS>(prm, s_tab, st, sq, slot, n_in, ctl);
        else lock_task<32, MODE_W6, PRESS>(prm, s_tab, st, sq, slot, n_in, ctl);
    }
}

// ---------------------------------------------------------------------------------------------
// K2b: the piece-binned throughput kernel for INDEPENDENT piece streams (BASELINE configs[4]: every game its own
// stream), w6, lookahead 1.  K2g keeps 18 of 32 threads busy there because the games of a warp are on different pieces
// (9 / 17 / 34 slots).  Here a game is not bound to a warp: seven device-wide queues, one per piece type, hold the games
// whose next piece has that type.  A warp pops up to 32 / LANES games from the fullest queue — all on the same piece —
// plays exactly one piece for each of them with every lane in step, and pushes each survivor onto the queue of its
// next piece.  Game state (96 bytes) travels through L2; with 5 warps per scheduler the pop / load latency hides behind
// the other warps' arithmetic.  Queues are rings of game ids with 64-bit reserved-tail / head counters: a pusher
// reserves slots by atomicAdd, writes the state, fences, then publishes the id into the slot; a popper claims by CAS
// on the head and reads each slot once it is published.  All waits are bounded (error flag instead of a hang).
// ---------------------------------------------------------------------------------------------
#define PB_EMPTY 0xffffffffu
constexpr uint32_t PB_SHARDS = 64;     // independent queue sets; a game always lives in shard game % PB_SHARDS, a warp's home shard is warp % PB_SHARDS

TB_D void pb_write_state(GameSave* dst, const uint32_t c[COLS], uint32_t t, uint32_t lines, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t h4,
                         uint32_t cnt, uint64_t gscore, const Heights& H, int cells)
{
    uint4* sv = reinterpret_cast<uint4*>(dst);
    __stcg(sv, make_uint4(c[0], c[1], c[2], c[3]));
    __stcg(sv + 1, make_uint4(c[4], c[5], c[6], c[7]));
    __stcg(sv + 2, make_uint4(c[8], c[9], t, lines));
    __stcg(sv + 3, make_uint4(h1, h2, h3, h4));
    __stcg(sv + 4, make_uint4(cnt, 0u, (uint32_t)gscore, (uint32_t)(gscore >> 32)));
    __stcg(sv + 5, make_uint4(H.H0, H.H1, H.H2, (uint32_t)cells));
}

// publish `game` at queue position pos of piece type p (the slot is empty unless a slow popper still holds it)
TB_D void pb_publish(const PlayParams& prm, uint32_t shard, int p, unsigned long long pos, uint32_t game)
{
    volatile uint32_t* slot = prm.pb_ring + ((size_t)shard * 7 + (size_t)p) * ((size_t)prm.pb_mask + 1) + (size_t)(pos & prm.pb_mask);
    uint32_t spins = 0;
    while (*slot != PB_EMPTY) { __nanosleep(100); if (++spins > (1u << 22)) { atomicExch(prm.pb_ctl + 16 * PB_SHARDS + 1, 1ull); break; } }
    *slot = game;
}

// every game enters the queue of its first piece (or finishes at once when there is nothing to play)
__global__ void pb_init_kernel(const PlayParams prm)
{
    const uint32_t game = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t T = prm.T;
    const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
    int p0 = 8;                                  // 8: no game in this thread
    if (game < prm.n_games) {
        const uint32_t sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[game] : game % (uint32_t)prm.G;
        uint32_t c[COLS];
#pragma unroll
        for (int k = 0; k < COLS; ++k) c[k] = 0;
        if (prm.init_rows) {
            uint16_t rows[ROWS];
            for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
            rows_to_cols(rows, c);
        }
        const uint32_t lines = prm.init_lines ? prm.init_lines[game] : 0u;
        p0 = stop == 0u ? 7 : (prm.pieces ? (int)prm.pieces[(size_t)sq * T] : hashed_piece(prm.seed, (uint64_t)sq, 0));
        if (p0 >= 7) {
            tb200_game_result* r = prm.results + game;
            r->pieces = 0; r->lines = lines; r->score = 0;
            if (prm.final_rows) { uint16_t rows[ROWS]; cols_to_rows(c, rows); for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2]; }
            atomicAdd(prm.pb_ctl + 16 * PB_SHARDS, 1ull);
            p0 = 8;
        } else {
            const GameStat g0 = game_stat(c);
            pb_write_state(prm.st_save + game, c, 0u, lines, 0u, 0u, 0u, 0u, 0u, 0ull, g0.H, g0.cells);
        }
    }
    if (p0 < 7) {
        const uint32_t shard = game % PB_SHARDS;
        const unsigned long long pos = atomicAdd(prm.pb_ctl + 16 * shard + p0, 1ull);
        prm.pb_ring[((size_t)shard * 7 + (size_t)p0) * ((size_t)prm.pb_mask + 1) + (size_t)(pos & prm.pb_mask)] = game;
    }
}

// (Measured and not kept: the window-local evaluation of K2sW inside this kernel — the games of a task are on the same piece, so the
// slot is warp-uniform here too — gave +1 %: a game is popped for ONE piece, so the window cache has to be rebuilt at every pop.)
template <int LANES>
__global__ void __launch_bounds__(BLOCK) play_games_w6bin_kernel(const PlayParams prm)
{
    constexpr uint32_t GPW = 32 / LANES;
    __shared__ Tables s_tab;
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const uint32_t T = prm.T;
    const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
    const size_t ring_stride = (size_t)prm.pb_mask + 1;
    const uint32_t warp_id = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    unsigned long long* const fin_ctr = prm.pb_ctl + 16 * PB_SHARDS;
    uint32_t idle = 0, probe = 0, hsum = 0, last_hsum = 0;
    unsigned long long last_fin = 0;

    for (;;) {
        // ---- pop up to GPW games of one piece type from the home shard (or, when that is empty, from the next shard that
        //      has work): lanes 0..6 look at one queue each, the fullest wins ----
        const uint32_t shard = (warp_id + probe) % PB_SHARDS;
        unsigned long long* const sctl = prm.pb_ctl + 16 * shard;
        unsigned long long hd = 0, avail = 0;
        if (lane < 7) { hd = *(volatile unsigned long long*)(sctl + 8 + lane); const unsigned long long tl = *(volatile unsigned long long*)(sctl + lane); avail = tl - hd; }
        const uint32_t a32 = avail > 0xffffffull ? 0xffffffu : (uint32_t)avail;
        // prefer full warps; among equally good queues rotate by warp so that the heads are not all hit at once
        const uint32_t score = (lane < 7 && a32 != 0u) ? (1u << 20) + ((a32 >= GPW ? GPW : a32) << 8 | ((uint32_t)(lane + warp_id) % 7u)) : 0u;
        const uint32_t best = __reduce_max_sync(FULLMASK, score);
        if (best == 0u) {
            hsum += __reduce_add_sync(FULLMASK, lane < 7 ? (uint32_t)hd : 0u);       // queue heads seen during this sweep: they advance whenever a piece is played
            if (++probe < PB_SHARDS) continue;                       // try the other shards before idling
            probe = 0;
            // nothing waiting anywhere: done when every game has finished, otherwise other warps are still playing
            const unsigned long long fin = *(volatile unsigned long long*)fin_ctr;
            if (fin >= (unsigned long long)prm.n_games) break;
            __nanosleep(1000);
            // progress = PIECES played anywhere (some queue head moved) or games finished, not games finished alone: a long
            // tail of a few surviving games with a large cap keeps every idle warp waiting instead of raising the flag
            if (hsum != last_hsum || fin != last_fin) { last_hsum = hsum; last_fin = fin; idle = 0; }
            hsum = 0;
            if (++idle > (1u << 16)) { if (lane == 0) atomicExch(fin_ctr + 1, 1ull); break; }    // seconds without a single piece played anywhere
            continue;
        }
        const int owner = __ffs((int)__ballot_sync(FULLMASK, score == best)) - 1;      // lane = piece type
        uint32_t n = 0;
        if (lane == owner) {
            n = a32 < GPW ? a32 : GPW;
            if (atomicCAS(sctl + 8 + lane, hd, hd + n) != hd) n = 0;                     // lost the race: look again
        }
        n = __shfl_sync(FULLMASK, n, owner);
        if (n == 0u) continue;
        idle = 0;
        const unsigned long long base = __shfl_sync(FULLMASK, hd, owner);
        const int p_cur = owner;
        // ---- fetch the games ----
        const uint32_t idx = (uint32_t)(lane / LANES);
        bool active = idx < n;
        uint32_t game = 0;
        if (active) {
            volatile uint32_t* slot = prm.pb_ring + ((size_t)shard * 7 + (size_t)p_cur) * ring_stride + (size_t)((base + idx) & prm.pb_mask);
            uint32_t spins = 0;
            while ((game = *slot) == PB_EMPTY) { __nanosleep(100); if (++spins > (1u << 22)) { atomicExch(fin_ctr + 1, 1ull); active = false; break; } }
        }
        __syncwarp(FULLMASK);
        if (active && gl == 0) *(volatile uint32_t*)(prm.pb_ring + ((size_t)shard * 7 + (size_t)p_cur) * ring_stride + (size_t)((base + idx) & prm.pb_mask)) = PB_EMPTY;
        __threadfence();
        uint32_t c[COLS];
        uint32_t t = 0, lines = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
        uint64_t gscore = 0;
#pragma unroll
        for (int k = 0; k < COLS; ++k) c[k] = 0;
        uint32_t sq = 0;
        GameStat g;
        g.H.H0 = g.H.H1 = g.H.H2 = 0u; g.cells = 0;
        if (active) {
            const uint4* sv = reinterpret_cast<const uint4*>(prm.st_save + game);
            const uint4 a0 = __ldcg(sv), a1 = __ldcg(sv + 1), a2 = __ldcg(sv + 2), a3 = __ldcg(sv + 3), a4 = __ldcg(sv + 4), a5 = __ldcg(sv + 5);
            g.H.H0 = a5.x; g.H.H1 = a5.y; g.H.H2 = a5.z; g.cells = (int)a5.w;
            c[0] = a0.x; c[1] = a0.y; c[2] = a0.z; c[3] = a0.w; c[4] = a1.x; c[5] = a1.y; c[6] = a1.z; c[7] = a1.w; c[8] = a2.x; c[9] = a2.y;
            t = a2.z; lines = a2.w; h1 = a3.x; h2 = a3.y; h3 = a3.z; h4 = a3.w;
            if (gl == 0) cnt = a4.x;
            gscore = ((uint64_t)a4.w << 32) | a4.z;
            sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[game] : game % (uint32_t)prm.G;
        }
        const uint8_t* seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
        unpack_heights(g.H, g.gh);
        double w[6];
        {
            const double* wsrc = prm.weights + (size_t)(active ? game / (uint32_t)prm.G : 0u) * 6;
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = __ldg(wsrc + i);
        }
        // ---- one piece, every lane in step ----
        const int ns = s_tab.nslots[p_cur];
        uint64_t key = 0; uint32_t tag = 0;
        for (int s2 = gl; s2 < ns; s2 += LANES) {
            const SlotDesc d = load_slot12(s_tab, p_cur, s2);
            W6Cand A;
            w6_pre(c, g, d, A);
            if (A.legal && A.full != 0u) w6_clear(d, A);
            Heights Hn;
            const uint64_t k = w6_post(A, w, Hn);
            cnt += (uint32_t)(A.legal && active);
            if (k > key) { key = k; tag = (uint32_t)s2; }
        }
        group_argmax<LANES>(key, tag);
#pragma unroll
        for (int off = LANES / 2; off > 0; off >>= 1) cnt += __shfl_xor_sync(FULLMASK, cnt, off);
        bool finish = false;
        int p_next = 7;
        if (active) {
            if (key == 0) {
                finish = true;                                           // no legal placement: game over (simulator.py:187-189)
            } else {
                const SlotDesc d = load_slot12(s_tab, p_cur, (int)tag);
                W6Cand A;
                w6_pre(c, g, d, A);
                if (A.full != 0u) w6_clear(d, A);
#pragma unroll
                for (int k = 0; k < COLS; ++k) c[k] = A.P.n[k];
                g.H.H0 = (uint32_t)A.h[0] | (uint32_t)A.h[1] << 8 | (uint32_t)A.h[2] << 16 | (uint32_t)A.h[3] << 24;
                g.H.H1 = (uint32_t)A.h[4] | (uint32_t)A.h[5] << 8 | (uint32_t)A.h[6] << 16 | (uint32_t)A.h[7] << 24;
                g.H.H2 = (uint32_t)A.h[8] | (uint32_t)A.h[9] << 8;
                g.cells = A.cells;
                const uint32_t wk = (uint32_t)A.P.k;
                if (wk != 0u) {
                    lines += wk;
                    h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                    const uint32_t pts = prm.pts[wk];
                    gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
                } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
                if (gl == 0) {
                    if (prm.trace) {
                        const uint32_t row = (uint32_t)(ROWS - A.P.y - A.P.ph);
                        reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = (uint32_t)A.P.rot | row << 8 | (uint32_t)A.P.c0 << 16 | wk << 24;
                    }
                    if (prm.score_trace) {
                        const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
                        prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                    }
                }
                ++t;
                if (t < stop) p_next = seq ? (int)__ldg(seq + t) : hashed_piece(prm.seed, (uint64_t)sq, t);
                if (p_next >= 7) finish = true;                          // cap reached or the sequence ends here
            }
        }
        // ---- finished games write their results; survivors go to the queue of their next piece ----
        const bool lead = active && gl == 0;
        if (lead && finish) {
            tb200_game_result* r = prm.results + game;
            r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4; r->score = gscore; r->placements = cnt;
            if (prm.final_rows) { uint16_t rows[ROWS]; cols_to_rows(c, rows); for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2]; }
        }
        if (lead && !finish) pb_write_state(prm.st_save + game, c, t, lines, h1, h2, h3, h4, cnt, gscore, g.H, g.cells);
        __threadfence();
        const uint32_t fin_mask = __ballot_sync(FULLMASK, lead && finish);
        if (fin_mask != 0u && lane == 0) atomicAdd(fin_ctr, (unsigned long long)__popc(fin_mask));
        const int my_next = (lead && !finish) ? p_next : 8;
        const uint32_t peers = __match_any_sync(FULLMASK, my_next);       // the survivors that go to the same queue
        if (my_next < 7) {
            unsigned long long pos = 0;
            const int leader = __ffs((int)peers) - 1;
            if (lane == leader) pos = atomicAdd(sctl + my_next, (unsigned long long)__popc(peers));      // the games of this task all live in `shard`
            pos = __shfl_sync(peers, pos, leader);
            pb_publish(prm, shard, my_next, pos + (unsigned long long)__popc(peers & ((1u << lane) - 1u)), game);
        }
    }
}

}  // namespace tb
