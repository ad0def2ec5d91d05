// play_latency.cuh — K2f and K2p: the w6 latency kernels (BASELINE config 2) with the re-spreading hand-over (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "play_common.cuh"

namespace tb {

// ---------------------------------------------------------------------------------------------
// K2f: one warp per game, lookahead 1, the w6 feature set — the BASELINE config-2 hot kernel.
// Per piece: [descriptors for this piece were prefetched] -> w6_pre (landing, legality, imprint,
// arithmetic heights) -> rare clear fix-up -> w6_post (transitions by popc, wells, fp64 score)
// -> REDUX argmax (+ ballot; the lo-word / tag REDUX only on a tie of the hi word) -> commit by
// broadcasting the winner's board, packed heights and move (14 SHFL, no clz, no popc).
// Slots 32 and 33 of the 34-slot pieces run as a second, interleaved instruction stream.
// ---------------------------------------------------------------------------------------------

template <bool STAGED>
__global__ void __launch_bounds__(BLOCK) play_games_w6fast_kernel(const PlayParams prm)
{
    constexpr bool staged = STAGED;
    uint32_t* const ticket = staged ? prm.st_ctl + prm.st_stage : prm.ticket;
    uint32_t n_in = prm.n_games;
    if (staged && prm.st_list_in) {
        n_in = prm.st_ctl[8 + prm.st_stage];
        if (n_in == 0u) return;                    // nothing was handed over: the earlier stage finished every game
    }
    const uint32_t* const finished = staged ? prm.st_ctl + 4 : nullptr;
    auto poll_finished = [&]() -> uint32_t {           // relaxed, L2-coherent; asm volatile so it is re-read every time
        uint32_t v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(finished) : "memory");
        return v;
    };
    __shared__ Tables s_tab;
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t T = prm.T;

    for (;;) {
        uint32_t game = 0;
        if (lane == 0) {
            game = atomicAdd(ticket, 1u);
            if (game < n_in && prm.st_list_in) game = prm.st_list_in[game]; else if (game >= n_in) game = 0xffffffffu;
        }
        game = __shfl_sync(FULLMASK, game, 0);
        if (game == 0xffffffffu) break;
        const bool resume = staged && prm.st_list_in != nullptr;
        const uint32_t sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[game] : (prm.per_game_weights != 0 ? game : game % (uint32_t)prm.G);
        const uint64_t stream = sq;
        const uint8_t* seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
        uint32_t c[COLS];
        uint32_t t = 0, lines = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
        uint64_t gscore = 0;
        if (resume) {
            const GameSave* sv = prm.st_save + game;
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = sv->c[k];
            t = sv->t; lines = sv->lines; h1 = sv->h1; h2 = sv->h2; h3 = sv->h3; h4 = sv->h4; gscore = sv->gscore;
            if (lane == 0) cnt = sv->cnt;
        } else if (prm.init_rows) {
            uint16_t rows[ROWS];
            for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
            rows_to_cols(rows, c);
        } else {
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = 0;
        }
        if (!resume && prm.init_lines) lines = prm.init_lines[game];
        GameStat g = game_stat(c);
        const int cells0 = g.cells - 4 * (int)t;       // g.cells = cells0 + 4 (pieces placed) - 10 (lines cleared since here)
        double w[6];
        {
            const size_t wrow = prm.per_game_weights > 0 ? (size_t)game : (prm.per_game_weights < 0 ? 0 : (size_t)(game / (uint32_t)prm.G));
            const double* wsrc = prm.weights + wrow * 6;
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = __ldg(wsrc + i);
        }
        // piece window: lane i holds pieces base+i (byte 0) and base+32+i (byte 1); handed out by shuffle
        auto load_piece = [&](uint32_t tt) -> uint32_t {
            if (tt >= T) return 7u;
            return seq ? (uint32_t)__ldg(seq + tt) : (uint32_t)hashed_piece(prm.seed, stream, tt);
        };
        uint32_t base = t;
        uint32_t win = load_piece(base + (uint32_t)lane) | load_piece(base + 32u + (uint32_t)lane) << 8;
        auto piece_at = [&](uint32_t tt) -> int {          // base <= tt < base + 64, tt warp-uniform
            const uint32_t idx = tt - base;
            const uint32_t v = __shfl_sync(FULLMASK, win, (int)(idx & 31u));
            return (int)((v >> ((idx >> 5) * 8u)) & 255u);
        };
        const uint32_t lines0 = lines;
        const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
        int p_cur = piece_at(t), p_next = piece_at(t + 1u);
        uint32_t fin_seen = staged ? poll_finished() : 0u;      // polled ahead of its use: the load never stalls the game
        bool handed = staged && prm.n_games - fin_seen <= prm.st_cut;
        // descriptors of the current piece (prefetched one piece ahead, they do not depend on the board);
        // word 7 of every descriptor of a piece (padding slots included) is its slot count
        const int slotB = lane < 2 ? lane + 32 : 0;
        SlotDesc dA = load_slot12(s_tab, p_cur < 7 ? p_cur : 0, lane);
        SlotDesc dB = load_slot12(s_tab, p_cur < 7 ? p_cur : 0, slotB);

#ifdef TB_PHASES
        long long ph[8] = {0,0,0,0,0,0,0,0}; long long last_ = clock64();
#endif
        bool over = false;                              // no legal placement
        while (!handed && t < stop && p_cur < 7) {
            uint32_t fin_next = fin_seen;
            if (staged && (t & 3u) == 0u) fin_next = poll_finished();
            PH(5)
            if (t + 2u - base >= 64u) {              // advance the piece window
                base += 32u;
                win = (win >> 8) | load_piece(base + 32u + (uint32_t)lane) << 8;
            }
            const int p_after = piece_at(t + 2u);
            // prefetch the next piece's descriptors
            const int pn = p_next < 7 ? p_next : 0;
            const SlotDesc dA_next = load_slot12(s_tab, pn, lane);
            const SlotDesc dB_next = load_slot12(s_tab, pn, slotB);
            const int ns = (int)dA.w[7];

            PH(0)
            W6Cand A;
            w6_pre(c, g, dA, A);
            A.legal = A.legal && lane < ns;
            uint64_t key; uint32_t tag = (uint32_t)lane;
            Heights Hn;
            if (ns > 32) {
                W6Cand B;
                w6_pre(c, g, dB, B);
                B.legal = B.legal && lane + 32 < ns;
                if ((A.legal && A.full != 0u) || (B.legal && B.full != 0u)) {
                    if (A.legal && A.full != 0u) w6_clear(dA, A);
                    if (B.legal && B.full != 0u) w6_clear(dB, B);
                }
                Heights HnB;
                const uint64_t kA = w6_post(A, w, Hn);
                const uint64_t kB = w6_post(B, w, HnB);
                cnt += (uint32_t)A.legal + (uint32_t)B.legal;
                key = kA;
                if (kB > kA) {                      // strict: the earlier slot keeps ties
                    key = kB; tag = (uint32_t)lane + 32u; Hn = HnB;
#pragma unroll
                    for (int k = 0; k < COLS; ++k) A.P.n[k] = B.P.n[k];
                    A.P.y = B.P.y; A.P.k = B.P.k; A.P.rot = B.P.rot; A.P.c0 = B.P.c0; A.P.ph = B.P.ph;
                }
            } else {
                if (A.legal && A.full != 0u) w6_clear(dA, A);
                key = w6_post(A, w, Hn);
                cnt += (uint32_t)A.legal;
            }
            PH(ns > 32 ? 2 : 1)
            // ---- argmax: hi word by REDUX; the lo word / tag only when two lanes share the hi word ----
            const uint32_t hi = (uint32_t)(key >> 32);
            const uint32_t mhi = __reduce_max_sync(FULLMASK, hi);
            if (mhi == 0u) { over = true; break; }  // no legal placement: game over (simulator.py:187-189)
            const uint32_t tie = __ballot_sync(FULLMASK, hi == mhi);
            int src = __ffs((int)tie) - 1;
            if (tie & (tie - 1u)) {
                const uint32_t lo = (uint32_t)key;
                const uint32_t mlo = __reduce_max_sync(FULLMASK, hi == mhi ? lo : 0u);
                const uint32_t mtag = __reduce_min_sync(FULLMASK, (hi == mhi && lo == mlo) ? tag : 0xffffffffu);
                src = (int)(mtag & 31u);
            }
            PH(3)
            // ---- commit: broadcast the winner's candidate board, heights and move ----
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = __shfl_sync(FULLMASK, A.P.n[k], src);
            g.H.H0 = __shfl_sync(FULLMASK, Hn.H0, src);
            g.H.H1 = __shfl_sync(FULLMASK, Hn.H1, src);
            g.H.H2 = __shfl_sync(FULLMASK, Hn.H2, src);
            const uint32_t info = __shfl_sync(FULLMASK, (uint32_t)A.P.y | (uint32_t)A.P.k << 8 | (uint32_t)A.P.rot << 12 | (uint32_t)A.P.c0 << 16 | (uint32_t)A.P.ph << 24, src);
            unpack_heights(g.H, g.gh);
            const uint32_t wk = (info >> 8) & 15u;
            if (wk != 0u) {
                lines += wk;
                h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                const uint32_t pts = prm.pts[wk];
                gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
            } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
            g.cells = cells0 + 4 * (int)(t + 1) - 10 * (int)(lines - lines0);
            if (prm.trace && lane == 0) {
                const uint32_t wy = info & 255u, wph = info >> 24;
                const uint32_t row = (uint32_t)ROWS - wy - wph;
                reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = ((info >> 12) & 15u) | row << 8 | ((info >> 16) & 255u) << 16 | wk << 24;
            }
            if (prm.score_trace) {
                const uint32_t khi = __shfl_sync(FULLMASK, (uint32_t)(key >> 32), src), klo = __shfl_sync(FULLMASK, (uint32_t)key, src);
                const uint64_t kk = ((uint64_t)khi << 32) | klo;
                const uint64_t b = (kk & 0x8000000000000000ull) ? (kk & 0x7FFFFFFFFFFFFFFFull) : ~kk;
                if (lane == 0) prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
            }
            ++t;
            p_cur = p_next; p_next = p_after;
            dA = dA_next; dB = dB_next;
            if (staged) { handed = prm.n_games - fin_seen <= prm.st_cut; fin_seen = fin_next; }
            PH(4)
        }
        handed = handed && !over && t < stop && p_cur < 7;      // a game that ended anyway is finished, not handed over
#ifdef TB_PHASES
        if (lane == 0) { for (int i = 0; i < 6; ++i) atomicAdd(&g_phase[i], (unsigned long long)ph[i]); atomicAdd(&g_phase[6], (unsigned long long)t); }
#endif
        cnt = __reduce_add_sync(FULLMASK, cnt);
        if (handed) {                                   // few enough games are left: hand this one to the next, sparser stage
            if (lane == 0) {
                GameSave* sv = prm.st_save + game;
#pragma unroll
                for (int k = 0; k < COLS; ++k) sv->c[k] = c[k];
                sv->t = t; sv->lines = lines; sv->h1 = h1; sv->h2 = h2; sv->h3 = h3; sv->h4 = h4; sv->cnt = cnt; sv->gscore = gscore;
                prm.st_list_out[atomicAdd(prm.st_ctl + 9 + prm.st_stage, 1u)] = game;
            }
            continue;
        }
        if (lane == 0) {
            if (staged) atomicAdd(prm.st_ctl + 4, 1u);
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

// ---------------------------------------------------------------------------------------------
// K2p: TWO warps per game (one 64-thread block per game), lookahead 1, w6 — for populations so
// small that warp schedulers would otherwise idle (BASELINE config 2: 1000 games on 592
// schedulers).  Slots are split even / odd between the two warps, so every piece — including the
// 34-slot T, L, J — is ONE pass of at most 17 lanes per warp.  Each warp finds its best slot
// (REDUX), its winning lane publishes {key, slot, move, board, heights} to shared memory, one
// block barrier, then both warps read both records, pick the global winner (higher key, then lower
// slot = first in enumeration order) and commit it.  Records are double-buffered by piece parity.
// Pieces are prefetched 32 at a time (one per lane) and handed out by shuffle.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) play_games_w6pair_kernel(const PlayParams prm)
{
    const bool staged = prm.st_ctl != nullptr;
    uint32_t* const ticket = staged ? prm.st_ctl + prm.st_stage : prm.ticket;
    uint32_t n_in = prm.n_games;
    if (staged && prm.st_list_in) {
        n_in = prm.st_ctl[8 + prm.st_stage];
        if (n_in == 0u) return;
    }
    __shared__ Tables s_tab;
    __shared__ uint4 s_x[2][2][5];
    __shared__ uint32_t s_game;
    load_tables<64>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = 2 * lane + warp;              // lanes 17.. have slot >= 34: never legal
    const int myslot = slot < TB_MAX_SLOTS ? slot : 0;
    const uint32_t T = prm.T;

    for (;;) {
        if (threadIdx.x == 0) {
            uint32_t gi = atomicAdd(ticket, 1u);
            if (gi < n_in && prm.st_list_in) gi = prm.st_list_in[gi]; else if (gi >= n_in) gi = 0xffffffffu;
            s_game = gi;
        }
        __syncthreads();
        const uint32_t game = s_game;
        __syncthreads();
        if (game == 0xffffffffu) break;
        const bool resume = staged && prm.st_list_in != nullptr;
        const uint32_t sq = prm.seq_of_game ? (uint32_t)prm.seq_of_game[game] : (prm.per_game_weights != 0 ? game : game % (uint32_t)prm.G);
        const uint64_t stream = sq;
        const uint8_t* seq = prm.pieces ? prm.pieces + (size_t)sq * T : nullptr;
        uint32_t c[COLS];
        uint32_t t = 0, lines = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, cnt = 0;
        uint64_t gscore = 0;
        if (resume) {
            const GameSave* sv = prm.st_save + game;
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = sv->c[k];
            t = sv->t; lines = sv->lines; h1 = sv->h1; h2 = sv->h2; h3 = sv->h3; h4 = sv->h4; gscore = sv->gscore;
            if (threadIdx.x == 0) cnt = sv->cnt;
        } else if (prm.init_rows) {
            uint16_t rows[ROWS];
            for (int r = 0; r < ROWS; ++r) rows[r] = prm.init_rows[(size_t)game * ROWS + r];
            rows_to_cols(rows, c);
        } else {
#pragma unroll
            for (int k = 0; k < COLS; ++k) c[k] = 0;
        }
        if (!resume && prm.init_lines) lines = prm.init_lines[game];
        GameStat g = game_stat(c);
        const int cells0 = g.cells - 4 * (int)t;
        double w[6];
        {
            const size_t wrow = prm.per_game_weights > 0 ? (size_t)game : (prm.per_game_weights < 0 ? 0 : (size_t)(game / (uint32_t)prm.G));
            const double* wsrc = prm.weights + wrow * 6;
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = __ldg(wsrc + i);
        }
        // piece window: lane i holds pieces base+i (byte 0) and base+32+i (byte 1)
        auto load_piece = [&](uint32_t tt) -> uint32_t {
            if (tt >= T) return 7u;
            return seq ? (uint32_t)__ldg(seq + tt) : (uint32_t)hashed_piece(prm.seed, stream, tt);
        };
        uint32_t base = t;
        uint32_t win = load_piece(base + (uint32_t)lane) | load_piece(base + 32u + (uint32_t)lane) << 8;
        auto piece_at = [&](uint32_t tt) -> int {          // base <= tt < base + 64, uniform tt
            const uint32_t idx = tt - base;
            const uint32_t v = __shfl_sync(FULLMASK, win, (int)(idx & 31u));
            return (int)((v >> ((idx >> 5) * 8u)) & 255u);
        };
        const uint32_t lines0 = lines;
        const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
        int p_cur = piece_at(t), p_next = piece_at(t + 1u);
        SlotDesc d = load_slot12(s_tab, p_cur < 7 ? p_cur : 0, myslot);
        uint32_t par = 0;

        while (t < stop && p_cur < 7) {
            // advance the piece window when the look-ahead index leaves it
            if (t + 2u - base >= 64u) {
                base += 32u;
                win = (win >> 8) | load_piece(base + 32u + (uint32_t)lane) << 8;
            }
            const int p_after = piece_at(t + 2u);
            const SlotDesc d_next = load_slot12(s_tab, p_next < 7 ? p_next : 0, myslot);

            W6Cand A;
            w6_pre(c, g, d, A);
            A.legal = A.legal && slot < (int)d.w[7];
            if (A.legal && A.full != 0u) w6_clear(d, A);
            Heights Hn;
            const uint64_t key = w6_post(A, w, Hn);
            cnt += (uint32_t)A.legal;
            // ---- warp-local argmax ----
            const uint32_t hi = (uint32_t)(key >> 32);
            const uint32_t mhi = __reduce_max_sync(FULLMASK, hi);
            const uint32_t tie = __ballot_sync(FULLMASK, hi == mhi);
            int src = __ffs((int)tie) - 1;
            if (mhi != 0u && (tie & (tie - 1u))) {
                const uint32_t lo = (uint32_t)key;
                const uint32_t mlo = __reduce_max_sync(FULLMASK, hi == mhi ? lo : 0u);
                const uint32_t mtag = __reduce_min_sync(FULLMASK, (hi == mhi && lo == mlo) ? (uint32_t)slot : 0xffffffffu);
                src = (int)(mtag >> 1);
            }
            if (lane == src) {
                uint4* rec = s_x[par][warp];
                const uint32_t info = (uint32_t)A.P.y | (uint32_t)A.P.k << 8 | (uint32_t)A.P.rot << 12 | (uint32_t)A.P.c0 << 16 | (uint32_t)A.P.ph << 24;
                rec[0] = make_uint4(hi, (uint32_t)key, (uint32_t)slot, info);
                rec[1] = make_uint4(A.P.n[0], A.P.n[1], A.P.n[2], A.P.n[3]);
                rec[2] = make_uint4(A.P.n[4], A.P.n[5], A.P.n[6], A.P.n[7]);
                rec[3] = make_uint4(A.P.n[8], A.P.n[9], Hn.H0, Hn.H1);
                rec[4] = make_uint4(Hn.H2, 0u, 0u, 0u);
            }
            __syncthreads();
            // ---- global winner of the two warps: higher key, then lower slot ----
            const uint4 ra = s_x[par][0][0], rb = s_x[par][1][0];
            const uint64_t ka = ((uint64_t)ra.x << 32) | ra.y, kb = ((uint64_t)rb.x << 32) | rb.y;
            if ((ra.x | rb.x) == 0u) break;         // no legal placement in either half: game over (simulator.py:187-189)
            const int wsel = (kb > ka || (kb == ka && rb.z < ra.z)) ? 1 : 0;
            const uint4* rec = s_x[par][wsel];
            const uint4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4];
            par ^= 1u;
            c[0] = r1.x; c[1] = r1.y; c[2] = r1.z; c[3] = r1.w; c[4] = r2.x; c[5] = r2.y; c[6] = r2.z; c[7] = r2.w; c[8] = r3.x; c[9] = r3.y;
            g.H.H0 = r3.z; g.H.H1 = r3.w; g.H.H2 = r4.x;
            unpack_heights(g.H, g.gh);
            const uint32_t info = r0.w;
            const uint32_t wk = (info >> 8) & 15u;
            if (wk != 0u) {
                lines += wk;
                h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                const uint32_t pts = prm.pts[wk];
                gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
            } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
            g.cells = cells0 + 4 * (int)(t + 1) - 10 * (int)(lines - lines0);
            if ((prm.trace != nullptr || prm.score_trace != nullptr) && threadIdx.x == 0) {     // uniform test first: no divergent region per piece when nothing is traced
                if (prm.trace) {
                    const uint32_t wy = info & 255u, wph = info >> 24;
                    const uint32_t row = (uint32_t)ROWS - wy - wph;
                    reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = ((info >> 12) & 15u) | row << 8 | ((info >> 16) & 255u) << 16 | wk << 24;
                }
                if (prm.score_trace) {
                    const uint64_t kk = wsel ? kb : ka;
                    const uint64_t b = (kk & 0x8000000000000000ull) ? (kk & 0x7FFFFFFFFFFFFFFFull) : ~kk;
                    prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                }
            }
            ++t;
            p_cur = p_next; p_next = p_after;
            d = d_next;
        }
        cnt = __reduce_add_sync(FULLMASK, cnt);
        if (lane == 0 && cnt) atomicAdd(reinterpret_cast<unsigned long long*>(&prm.results[game].placements), (unsigned long long)cnt);
        if (threadIdx.x == 0) {
            tb200_game_result* r = prm.results + game;
            r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4;
            r->score = gscore;
            if (prm.final_rows) {
                uint16_t rows[ROWS];
                cols_to_rows(c, rows);
                for (int r2 = 0; r2 < ROWS; ++r2) prm.final_rows[(size_t)game * ROWS + r2] = rows[r2];
            }
        }
    }
}

}  // namespace tb
