// play_lookahead2.cuh — K2L: lookahead 2, leaves flattened over the lanes (BASELINE configs[3]) (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "play_common.cuh"

namespace tb {

// ---------------------------------------------------------------------------------------------
// K2L: lookahead 2 with hard drops, one warp per game (BASELINE configs[3]).  The <= 34 legal layer-1 boards
// (post-clear columns, packed heights, cached statistics, the move) are first built into shared memory by the
// lanes; then the <= 34 x 34 leaves are FLATTENED over the 32 lanes (leaf -> (layer-1 rank, second slot)), so
// that lanes stay busy whatever the slot counts of the two pieces are (a per-first-move loop leaves half of the
// lanes idle: 9 / 17 / 34 second slots on 32 lanes).  Leaf-only scoring, tag = rank << 6 | second slot keeps the
// first-of-max order; when layer 2 is empty the layer-1 states are scored instead (simulator.py:170-174).
// ---------------------------------------------------------------------------------------------
struct L1Rec {                 // one legal first move
    uint32_t n[COLS];          // board after it
    uint32_t H0, H1, H2;       // packed heights
    int max_ht, all_ht, pits, cells, possum;
    uint32_t mv;               // slot | landing height << 8
    uint32_t pad;
};

template <int MODE>
__global__ void __launch_bounds__(BLOCK) play_games_l2flat_kernel(const PlayParams prm)
{
    constexpr int GROUPS = BLOCK / 32;
    __shared__ Tables s_tab;
    __shared__ L1Rec s_rec[GROUPS][TB_MAX_SLOTS];
    __shared__ double s_w[GROUPS][TB200_MAX_WEIGHTED];
    load_tables<BLOCK>(s_tab);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x >> 5;
    L1Rec* rec = s_rec[grp];
    const int F = prm.F;
    const uint32_t T = prm.T;
    const double* w = s_w[grp];

    auto score_leaf = [&](const uint32_t* c1, const Heights& H1, const PrevStat& pv1, const SlotDesc& d2, uint64_t& k, bool& legal) {
        if (MODE == MODE_W6) {
            GameStat g1;
            g1.H = H1; g1.cells = pv1.cells;
#pragma unroll
            for (int q = 0; q < COLS; ++q) g1.gh[q] = pv1.gh[q];
            W6Cand A;
            w6_pre(c1, g1, d2, A);
            if (A.legal && A.full != 0u) w6_clear(d2, A);
            Heights Hn;
            k = w6_post(A, w, Hn);
            legal = A.legal;
        } else {
            Placement P2;
            legal = place_slot(c1, H1, d2, P2);
            Feat f;
            eval_features<ModeTraits<MODE>::CMASK>(P2, d2, pv1, prm.rmask, f);
            const uint64_t kk = score_key(score_features<MODE>(f, prm.ids, w, F));
            k = legal ? kk : 0ull;
        }
    };

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
            const int ns1 = s_tab.nslots[p_cur];
            // ---- phase 1: legal first moves -> records in shared memory, in enumeration order ----
            int nrec = 0;
            __syncwarp(FULLMASK);
            for (int base = 0; base < ns1; base += 32) {
                const int s1 = base + lane;
                const SlotDesc d1 = load_slot12(s_tab, p_cur, s1 < ns1 ? s1 : 0);
                Placement P1;
                const bool legal1 = place_slot(c, H, d1, P1) && s1 < ns1;
                const uint32_t m = __ballot_sync(FULLMASK, legal1);
                if (legal1) {
                    const PrevStat pv1 = next_stat(pv, P1, d1);
                    const Heights H1 = pack_gh(pv1.gh);
                    L1Rec& r = rec[nrec + __popc(m & ((1u << lane) - 1u))];
#pragma unroll
                    for (int k = 0; k < COLS; ++k) r.n[k] = P1.n[k];
                    r.H0 = H1.H0; r.H1 = H1.H1; r.H2 = H1.H2;
                    r.max_ht = pv1.max_ht; r.all_ht = pv1.all_ht; r.pits = pv1.pits; r.cells = pv1.cells; r.possum = pv1.possum;
                    r.mv = (uint32_t)s1 | (uint32_t)P1.y << 8;
                }
                nrec += __popc(m);
            }
            __syncwarp(FULLMASK);
            uint64_t key = 0; uint32_t tag = 0;
            bool leaf_is_first = true;
            if (t + 1 < T && p_next < 7 && nrec > 0) {
                // ---- phase 2: leaves flattened over the lanes: leaf idx -> (record r, second slot s2) ----
                const int ns2 = s_tab.nslots[p_next];
                const int total = nrec * ns2;
                int r = 0, s2 = lane;
                while (s2 >= ns2) { s2 -= ns2; ++r; }
                for (int idx = lane; idx < total; idx += 32) {
                    const L1Rec& R = rec[r];
                    uint32_t c1[COLS];
#pragma unroll
                    for (int k = 0; k < COLS; ++k) c1[k] = R.n[k];
                    Heights H1; H1.H0 = R.H0; H1.H1 = R.H1; H1.H2 = R.H2;
                    PrevStat pv1;
                    pv1.max_ht = R.max_ht; pv1.all_ht = R.all_ht; pv1.pits = R.pits; pv1.cells = R.cells; pv1.possum = R.possum;
                    unpack_heights(H1, pv1.gh);
                    const SlotDesc d2 = load_slot12(s_tab, p_next, s2);
                    uint64_t k; bool legal2;
                    score_leaf(c1, H1, pv1, d2, k, legal2);
                    if (legal2) ++cnt;
                    if (k > key) { key = k; tag = (uint32_t)r << 6 | (uint32_t)s2; }
                    s2 += 32;
                    while (s2 >= ns2) { s2 -= ns2; ++r; }
                }
                warp_argmax32(key, tag);
                if (key != 0) leaf_is_first = false;
            }
            uint32_t mv;
            if (leaf_is_first) {
                // layer 2 empty (or the sequence ends): score the layer-1 states themselves
                key = 0; tag = 0;
                if (lane < nrec || lane + 32 < nrec) {
                    for (int r = lane; r < nrec; r += 32) {
                        const SlotDesc d1 = load_slot12(s_tab, p_cur, (int)(rec[r].mv & 255u));
                        uint64_t k; bool legal1;
                        score_leaf(c, H, pv, d1, k, legal1);
                        if (legal1) ++cnt;
                        if (k > key) { key = k; tag = (uint32_t)r; }
                    }
                }
                warp_argmax32(key, tag);
                if (key == 0) break;
                mv = rec[tag].mv;
            } else {
                mv = rec[tag >> 6].mv;
            }
            // ---- commit the first move ----
            const SlotDesc d = load_slot12(s_tab, p_cur, (int)(mv & 255u));
            Placement P; uint32_t full;
            place_at(c, d, (int)(mv >> 8), P, &full);
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
