// emu.cpp — host build of the CUDA core (tetris_ai_b200/csrc/tetris_core.cuh) for CPU-side tests.
//
// TEST INFRASTRUCTURE ONLY: it lets `pytest -m "not gpu"` run the very same placement / feature /
// scoring functions the kernels use against the oracles in a container without a GPU.  It is not
// part of the product library and libtetris_b200.so exports none of it.  The sequential loops here
// mirror the kernels' search order (slot-major, strict-greater, first-of-max) but not their
// shuffles or tickets — those are covered by the -m gpu tests.
#include <stdint.h>
#include <string.h>
#include "../../tetris_ai_b200/csrc/tetris_core.cuh"

using namespace tb;

static int pick_mode(const int* ids, int F)
{
    static const int w6[6] = {F_LANDING_HEIGHT, F_ERODED_CELLS, F_ROW_TRANS, F_COL_TRANS, F_PITS, F_CUML_WELLS};
    bool is_w6 = F == 6, is_all = F == NFEAT;
    for (int i = 0; i < F && is_w6; ++i) is_w6 = ids[i] == w6[i];
    for (int i = 0; i < F && is_all; ++i) is_all = ids[i] == i;
    return is_w6 ? MODE_W6 : (is_all ? MODE_ALL45 : MODE_GENERIC);
}

// move list of move_wiggle (from_top) / move_with_wiggle(move_drop): entries = slot | y << 8 in the reference's order
static int wiggle_list(const uint32_t c[COLS], const Heights& H, int piece, bool from_top, uint16_t* out, int cap)
{
    const Tables& tab = host_tables();
    int n = 0, s0 = 0;
    while (s0 < tab.nslots[piece]) {
        const int rot = (int)((tab.slot[piece][s0].w[0] >> 4) & 3u), w = (int)((tab.slot[piece][s0].w[0] >> 6) & 7u), ph = (int)((tab.slot[piece][s0].w[0] >> 9) & 7u);
        const int ncol = COLS - w + 1;
        uint32_t fm[COLS + 1]; uint16_t starts[COLS]; int ns = 0;
        for (int i = 0; i <= COLS; ++i) fm[i] = 0;
        for (int col = 0; col < ncol; ++col) {
            fm[col] = free_mask(c, tab.slot[piece][s0 + col]);
            if (from_top) starts[ns++] = (uint16_t)((ROWS - ph) | col << 8);
            else { Placement P; if (place_slot(c, H, tab.slot[piece][s0 + col], P)) starts[ns++] = (uint16_t)(P.y | col << 8); }
        }
        uint16_t stack[256]; uint32_t seen[COLS + 1]; uint16_t lst[256];
        const int m = wiggle_rot(fm, w, starts, ns, stack, seen, lst, 256);
        for (int i = 0; i < m && n < cap; ++i) out[n++] = (uint16_t)((uint32_t)(s0 + (lst[i] >> 8)) | (uint32_t)(lst[i] & 31u) << 8);
        (void)rot;
        s0 += ncol;
    }
    return n;
}

// self-check of wiggle_rot's stop set: for `trials` random subsets of each rotation's move list the early-out search must
// return the first listed move of the subset.  Returns the number of failures.
extern "C" int emu_wiggle_stop_check(const uint16_t* rows, int piece, int from_top, uint32_t rng, int trials)
{
    const Tables& tab = host_tables();
    uint32_t c[COLS];
    rows_to_cols(rows, c);
    const Heights H = pack_heights(c);
    int bad = 0, s0 = 0;
    while (s0 < tab.nslots[piece]) {
        const int w = (int)((tab.slot[piece][s0].w[0] >> 6) & 7u), ph = (int)((tab.slot[piece][s0].w[0] >> 9) & 7u);
        const int ncol = COLS - w + 1;
        uint32_t fm[COLS + 1]; uint16_t starts[COLS]; int ns = 0;
        for (int i = 0; i <= COLS; ++i) fm[i] = 0;
        for (int col = 0; col < ncol; ++col) {
            fm[col] = free_mask(c, tab.slot[piece][s0 + col]);
            if (from_top) starts[ns++] = (uint16_t)((ROWS - ph) | col << 8);
            else { Placement P; if (place_slot(c, H, tab.slot[piece][s0 + col], P)) starts[ns++] = (uint16_t)(P.y | col << 8); }
        }
        uint16_t stack[256]; uint32_t seen[COLS + 1]; uint16_t lst[256], one[256];
        const int m = wiggle_rot(fm, w, starts, ns, stack, seen, lst, 256);
        for (int tr = 0; tr < trials && m > 0; ++tr) {
            uint32_t stop[COLS + 1];
            for (int i = 0; i <= COLS; ++i) stop[i] = 0;
            int first = -1;
            for (int i = 0; i < m; ++i) {
                rng = rng * 1664525u + 1013904223u;
                if ((rng >> 29) == 0u || (first < 0 && i == m - 1)) { stop[lst[i] >> 8] |= 1u << (lst[i] & 31u); if (first < 0) first = i; }
            }
            const int r = wiggle_rot(fm, w, starts, ns, stack, seen, one, 256, stop);
            if (r != 1 || one[0] != lst[first]) ++bad;
        }
        s0 += ncol;
    }
    return bad;
}

// candidate bookkeeping: maximise key, ties -> smallest tag (enumeration order); mv = placement to commit
struct Best { uint64_t key; uint32_t tag; uint32_t mv; };

template <int MODE>
static void search_one(const uint32_t c[COLS], const Heights& H, const PrevStat& pv, int piece, const uint8_t* ids,
                       const double* w, int F, uint64_t rmask, int move_set, const Pressure& pr, int level, uint32_t tag_hi, uint32_t mv1,
                       bool leaf_is_first, Best& best, uint64_t& cnt)
{
    const Tables& tab = host_tables();
    auto consider = [&](const Placement& P, const SlotDesc& d, uint32_t tg, uint32_t mv) {
        if (pr.on && !is_time(d, (int)(mv >> 8), pr, level)) return;
        Feat f;
        memset(&f, 0, sizeof f);
        eval_features<ModeTraits<MODE>::CMASK>(P, d, pv, rmask, f);
        const uint64_t k = score_key(score_features<MODE>(f, ids, w, F));
        if (k > best.key || (k == best.key && tg < best.tag)) { best.key = k; best.tag = tg; best.mv = leaf_is_first ? mv : mv1; }
        ++cnt;
    };
    if (move_set & 12) {                 // wiggle move sets: the list order is the tag order
        uint16_t lst[640];
        const int m = wiggle_list(c, H, piece, (move_set & 4) != 0, lst, 640);
        for (int i = 0; i < m; ++i) {
            const SlotDesc& d2 = tab.slot[piece][lst[i] & 255u];
            Placement P2; uint32_t full;
            place_at(c, d2, (int)(lst[i] >> 8), P2, &full);
            if (full != 0u) place_clear(d2, P2, full);
            consider(P2, d2, tag_hi | (uint32_t)i, (uint32_t)lst[i]);
        }
        return;
    }
    for (int s = 0; s < tab.nslots[piece]; ++s) {
        Placement P;
        if (!place_slot(c, H, tab.slot[piece][s], P)) continue;
        consider(P, tab.slot[piece][s], tag_hi | (uint32_t)s << 5, (uint32_t)s | (uint32_t)P.y << 8);
        if (move_set & 1) {
            for_each_slide(c, H, tab.slot[piece][s], s, P.y, [&](int s2) { return tab.slot[piece][s2]; },
                [&](const SlotDesc& d2, int s2, int yy, int j) {
                    Placement P2; uint32_t full;
                    place_at(c, d2, yy, P2, &full);
                    if (full != 0u) place_clear(d2, P2, full);
                    consider(P2, d2, tag_hi | (uint32_t)s << 5 | (uint32_t)j, (uint32_t)s2 | (uint32_t)yy << 8);
                });
        }
    }
}

template <int MODE>
static void play(const int* ids_in, int F, const double* w, const uint8_t* pieces, uint32_t T, uint64_t seed, uint64_t stream,
                 int lookahead, uint32_t max_moves, const uint16_t* init_rows, uint64_t* res8, uint8_t* trace,
                 double* score_trace, uint16_t* final_rows, int move_set, const double* press)
{
    const Tables& tab = host_tables();
    Pressure pr = {0, 0, 0.0, 0.0, 0.0};
    if ((move_set & 2) && press) { pr.on = 1; pr.avg_lat = press[0]; pr.resp_time = press[1]; pr.move_eff = press[2]; pr.two_but_rot = press[3] != 0.0; }
    uint8_t ids[64]; uint64_t rmask = 0;
    for (int i = 0; i < F; ++i) { ids[i] = (uint8_t)ids_in[i]; rmask |= 1ull << ids_in[i]; }
    uint32_t c[COLS];
    if (init_rows) rows_to_cols(init_rows, c); else memset(c, 0, sizeof c);
    PrevStat pv = prev_stat(c); Heights H = pack_gh(pv.gh);
    uint32_t t = 0, lines = 0, hist[4] = {0, 0, 0, 0}; uint64_t score = 0, cnt = 0;
    if (!max_moves) max_moves = T;
    auto piece_at = [&](uint32_t tt) { return pieces ? (int)pieces[tt] : hashed_piece(seed, stream, tt); };
    while (t < T && t < max_moves) {
        const int p1 = piece_at(t);
        if (p1 >= 7) break;
        Best best = {0, 0, 0};
        const int p2 = (lookahead == 2 && t + 1 < T) ? piece_at(t + 1) : 255;
        if (p2 < 7) {
            // layer 1 in enumeration order (base move, its right slides, its left slides), each expanded by piece 2
            auto expand = [&](const Placement& P1, const SlotDesc& d1, uint32_t tag1, uint32_t mv1) {
                if (pr.on && !is_time(d1, (int)(mv1 >> 8), pr, (int)(lines / 10))) return;
                const PrevStat pv1 = next_stat(pv, P1, d1); const Heights H1 = pack_gh(pv1.gh);
                search_one<MODE>(P1.n, H1, pv1, p2, ids, w, F, rmask, move_set, pr, (int)((lines + (uint32_t)P1.k) / 10), tag1 << 11, mv1, false, best, cnt);
            };
            if (move_set & 12) {             // wiggle move sets: layer 1 = the flood-fill list, in its order
                uint16_t lst[640];
                const int m = wiggle_list(c, H, p1, (move_set & 4) != 0, lst, 640);
                for (int i = 0; i < m; ++i) {
                    const SlotDesc& d2 = tab.slot[p1][lst[i] & 255u];
                    Placement P2; uint32_t full;
                    place_at(c, d2, (int)(lst[i] >> 8), P2, &full);
                    if (full != 0u) place_clear(d2, P2, full);
                    expand(P2, d2, (uint32_t)i, (uint32_t)lst[i]);
                }
            } else
            for (int s1 = 0; s1 < tab.nslots[p1]; ++s1) {
                Placement P1;
                if (!place_slot(c, H, tab.slot[p1][s1], P1)) continue;
                expand(P1, tab.slot[p1][s1], (uint32_t)s1 << 5, (uint32_t)s1 | (uint32_t)P1.y << 8);
                if (move_set & 1) {
                    for_each_slide(c, H, tab.slot[p1][s1], s1, P1.y, [&](int s2) { return tab.slot[p1][s2]; },
                        [&](const SlotDesc& d2, int s2, int yy, int j) {
                            Placement P2; uint32_t full;
                            place_at(c, d2, yy, P2, &full);
                            if (full != 0u) place_clear(d2, P2, full);
                            expand(P2, d2, (uint32_t)s1 << 5 | (uint32_t)j, (uint32_t)s2 | (uint32_t)yy << 8);
                        });
                }
            }
        }
        if (best.key == 0) {
            Best b1 = {0, 0, 0};
            search_one<MODE>(c, H, pv, p1, ids, w, F, rmask, move_set, pr, (int)(lines / 10), 0u, 0u, true, b1, cnt);
            best = b1;
        }
        if (best.key == 0) break;
        const SlotDesc& dw = tab.slot[p1][best.mv & 255u];
        Placement P; uint32_t full;
        place_at(c, dw, (int)(best.mv >> 8), P, &full);
        if (full != 0u) place_clear(dw, P, full);
        memcpy(c, P.n, sizeof c);
        pv = next_stat(pv, P, dw); H = pack_gh(pv.gh);
        lines += (uint32_t)P.k; if (P.k) hist[P.k - 1] += 1;
        static const uint32_t PTS[5] = {0, 40, 100, 300, 1200};
        score += (uint64_t)PTS[P.k] * (lines / 10 + 1);
        if (trace) { trace[4 * t] = (uint8_t)P.rot; trace[4 * t + 1] = (uint8_t)(ROWS - P.y - P.ph); trace[4 * t + 2] = (uint8_t)P.c0; trace[4 * t + 3] = (uint8_t)P.k; }
        if (score_trace) {
            const uint64_t b = (best.key & 0x8000000000000000ull) ? (best.key & 0x7FFFFFFFFFFFFFFFull) : ~best.key;
            memcpy(&score_trace[t], &b, 8);
        }
        ++t;
    }
    res8[0] = t; res8[1] = lines; res8[2] = hist[0]; res8[3] = hist[1]; res8[4] = hist[2]; res8[5] = hist[3]; res8[6] = score; res8[7] = cnt;
    if (final_rows) cols_to_rows(c, final_rows);
}

// the w32 kernel's fast path for the w6 set (w6_pre / w6_clear / w6_post), lookahead 1
static void play_w6_fast(const double* w, const uint8_t* pieces, uint32_t T, uint64_t seed, uint64_t stream, uint32_t max_moves,
                         const uint16_t* init_rows, uint64_t* res8, uint8_t* trace, double* score_trace, uint16_t* final_rows)
{
    const Tables& tab = host_tables();
    uint32_t c[COLS];
    if (init_rows) rows_to_cols(init_rows, c); else memset(c, 0, sizeof c);
    GameStat g = game_stat(c);
    const int cells0 = g.cells;
    uint32_t t = 0, lines = 0, hist[4] = {0, 0, 0, 0}; uint64_t score = 0, cnt = 0;
    if (!max_moves) max_moves = T;
    while (t < T && t < max_moves) {
        const int p1 = pieces ? (int)pieces[t] : hashed_piece(seed, stream, t);
        if (p1 >= 7) break;
        uint64_t key = 0; int best = -1; W6Cand bestc; Heights bestH;
        memset(&bestc, 0, sizeof bestc); memset(&bestH, 0, sizeof bestH);
        for (int s = 0; s < tab.nslots[p1]; ++s) {
            W6Cand a;
            w6_pre(c, g, tab.slot[p1][s], a);
            if (a.legal && a.full != 0u) w6_clear(tab.slot[p1][s], a);
            Heights Hn;
            const uint64_t k = w6_post(a, w, Hn);
            cnt += a.legal ? 1 : 0;
            if (k > key) { key = k; best = s; bestc = a; bestH = Hn; }
        }
        if (best < 0) break;
        memcpy(c, bestc.P.n, sizeof c);
        g.H = bestH; unpack_heights(g.H, g.gh);
        const Placement& P = bestc.P;
        lines += (uint32_t)P.k; if (P.k) hist[P.k - 1] += 1;
        g.cells = cells0 + 4 * (int)(t + 1) - 10 * (int)lines;
        static const uint32_t PTS[5] = {0, 40, 100, 300, 1200};
        score += (uint64_t)PTS[P.k] * (lines / 10 + 1);
        if (trace) { trace[4 * t] = (uint8_t)P.rot; trace[4 * t + 1] = (uint8_t)(ROWS - P.y - P.ph); trace[4 * t + 2] = (uint8_t)P.c0; trace[4 * t + 3] = (uint8_t)P.k; }
        if (score_trace) {
            const uint64_t b = (key & 0x8000000000000000ull) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
            memcpy(&score_trace[t], &b, 8);
        }
        ++t;
    }
    res8[0] = t; res8[1] = lines; res8[2] = hist[0]; res8[3] = hist[1]; res8[4] = hist[2]; res8[5] = hist[3]; res8[6] = score; res8[7] = cnt;
    if (final_rows) cols_to_rows(c, final_rows);
}

extern "C" {

// mode: -1 = as tb200_set_features would choose, 0 = force the generic (run-time feature list) path
int emu_play_one(const int* ids, int F, const double* w, const uint8_t* pieces, uint32_t T, uint64_t seed, uint64_t stream,
                 int lookahead, uint32_t max_moves, int mode, const uint16_t* init_rows, uint64_t* res8, uint8_t* trace,
                 double* score_trace, uint16_t* final_rows, int move_set, const double* press)
{
    const int m = mode < 0 ? pick_mode(ids, F) : mode;
    if (m == 3) {           // fast w6 path of the one-warp-per-game kernel (caller guarantees the w6 feature list, lookahead 1)
        play_w6_fast(w, pieces, T, seed, stream, max_moves, init_rows, res8, trace, score_trace, final_rows);
        return m;
    }
    if (m == MODE_W6) play<MODE_W6>(ids, F, w, pieces, T, seed, stream, lookahead, max_moves, init_rows, res8, trace, score_trace, final_rows, move_set, press);
    else if (m == MODE_ALL45) play<MODE_ALL45>(ids, F, w, pieces, T, seed, stream, lookahead, max_moves, init_rows, res8, trace, score_trace, final_rows, move_set, press);
    else play<MODE_GENERIC>(ids, F, w, pieces, T, seed, stream, lookahead, max_moves, init_rows, res8, trace, score_trace, final_rows, move_set, press);
    return m;
}

// all 45 features of prev.future(piece, rot, <drop row>, col); info = {legal (-1 = no such slot), row, k}
void emu_eval_placement(const uint16_t* prev_rows, int piece, int rot, int col, double* feats, uint16_t* post_rows, int* info)
{
    const Tables& tab = host_tables();
    uint32_t c[COLS];
    rows_to_cols(prev_rows, c);
    const Heights H = pack_heights(c); const PrevStat pv = prev_stat(c);
    info[0] = -1; info[1] = -1; info[2] = 0;
    for (int s = 0; s < tab.nslots[piece]; ++s) {
        const SlotDesc& d = tab.slot[piece][s];
        if ((int)((d.w[0] >> 4) & 3u) != rot || (int)(d.w[0] & 15u) != col) continue;
        Placement P;
        const bool legal = place_slot(c, H, d, P);
        info[0] = legal ? 1 : 0;
        if (legal) {
            info[1] = ROWS - P.y - P.ph; info[2] = P.k;
            Feat f; memset(&f, 0, sizeof f);
            eval_features<0>(P, d, pv, ALL45_MASK, f);
            for (int j = 0; j < NFEAT; ++j) feats[j] = feat_value(f, j);
            if (post_rows) cols_to_rows(P.n, post_rows);
        }
        return;
    }
}

int emu_hashed_piece(uint64_t seed, uint64_t stream, uint64_t t) { return hashed_piece(seed, stream, t); }

// the wiggle move SET (bit-parallel flood fill, no order): (rot, row, col) triples in slot order
int emu_wiggle_set(const uint16_t* rows, int piece, int from_top, int* out)
{
    const Tables& tab = host_tables();
    uint32_t c[COLS];
    rows_to_cols(rows, c);
    const Heights H = pack_heights(c);
    int n = 0, s0 = 0;
    while (s0 < tab.nslots[piece]) {
        const int rot = (int)((tab.slot[piece][s0].w[0] >> 4) & 3u), w = (int)((tab.slot[piece][s0].w[0] >> 6) & 7u), ph = (int)((tab.slot[piece][s0].w[0] >> 9) & 7u);
        const int ncol = COLS - w + 1;
        uint32_t fm[COLS + 1], reach[COLS + 1]; uint16_t starts[COLS]; int ns = 0;
        for (int i = 0; i <= COLS; ++i) fm[i] = 0;
        for (int col = 0; col < ncol; ++col) {
            fm[col] = free_mask(c, tab.slot[piece][s0 + col]);
            if (from_top) starts[ns++] = (uint16_t)((ROWS - ph) | col << 8);
            else { Placement P; if (place_slot(c, H, tab.slot[piece][s0 + col], P)) starts[ns++] = (uint16_t)(P.y | col << 8); }
        }
        wiggle_reach(fm, w, starts, ns, reach);
        for (int col = 0; col < ncol; ++col)
            for (int y = 0; y < ROWS; ++y)
                if ((reach[col] >> y) & 1u) { out[3 * n] = rot; out[3 * n + 1] = ROWS - y - ph; out[3 * n + 2] = col; ++n; }
        s0 += ncol;
    }
    return n;
}

// move list (rot, row, col) of move_set 0 (move_drop) / 1 (move_with_overhang_slide(move_drop)); returns the count
int emu_moves(const uint16_t* rows, int piece, int move_set, int* out)
{
    const Tables& tab = host_tables();
    uint32_t c[COLS];
    rows_to_cols(rows, c);
    const Heights H = pack_heights(c);
    int n = 0;
    if (move_set & 12) {
        uint16_t lst[640];
        const int m = wiggle_list(c, H, piece, (move_set & 4) != 0, lst, 640);
        for (int i = 0; i < m; ++i) {
            const SlotDesc& d2 = tab.slot[piece][lst[i] & 255u];
            out[3 * n] = (int)((d2.w[0] >> 4) & 3u); out[3 * n + 1] = ROWS - (int)(lst[i] >> 8) - (int)((d2.w[0] >> 9) & 7u); out[3 * n + 2] = (int)(d2.w[0] & 15u); ++n;
        }
        return n;
    }
    for (int s = 0; s < tab.nslots[piece]; ++s) {
        Placement P;
        if (!place_slot(c, H, tab.slot[piece][s], P)) continue;
        out[3 * n] = P.rot; out[3 * n + 1] = ROWS - P.y - P.ph; out[3 * n + 2] = P.c0; ++n;
        if (move_set == 1)
            for_each_slide(c, H, tab.slot[piece][s], s, P.y, [&](int s2) { return tab.slot[piece][s2]; },
                [&](const SlotDesc& d2, int, int yy, int) {
                    out[3 * n] = (int)((d2.w[0] >> 4) & 3u); out[3 * n + 1] = ROWS - yy - (int)((d2.w[0] >> 9) & 7u); out[3 * n + 2] = (int)(d2.w[0] & 15u); ++n;
                });
    }
    return n;
}

}  // extern "C"


// Window-local w6 evaluation (win_eval) against the general w6 evaluation (w6_pre / w6_post) for EVERY slot of every piece
// on one board.  Returns the number of slots compared in the high 16 bits... no: returns mismatches; *compared receives the
// number of legal, non-clearing slots whose keys were compared, *clearing the number of slots that complete a row.
extern "C" int emu_win_check(const uint16_t* rows, const double* w6w, int* compared, int* clearing)
{
    const Tables& tab = host_tables();
    uint32_t c[COLS];
    rows_to_cols(rows, c);
    const GameStat g = game_stat(c);
    WinCache wc;
    win_cache_init(c, wc);
    int bad = 0, ncmp = 0, nclr = 0;
    for (int piece = 0; piece < 7; ++piece) {
        WinSlide ws;
        win_reset(wc, ws);
        for (int s = 0; s < tab.nslots[piece]; ++s) {
            const SlotDesc& d = tab.slot[piece][s];
            if ((d.w[0] & 15u) == 0u) win_reset(wc, ws);                 // new rotation: the windows slide from column 0 again
            W6Cand A;
            w6_pre(c, g, d, A);
            const int c0 = (int)(d.w[0] & 15u), wd = (int)((d.w[0] >> 6) & 7u);
            uint32_t and_out = 0xffffffffu;
            for (int k = 0; k < COLS; ++k) if (k < c0 || k >= c0 + wd) and_out &= c[k];
            uint32_t full = 0; bool legal = false;
            const uint64_t kw = win_eval(wc, d, w6w, [&](int k) { return (k >= 0 && k < COLS) ? c[k] : 0xdeadbeefu; }, and_out, &full, &legal);
            {   // the sliding windows the device uses must give the same key / full mask / legality
                uint32_t full2 = 0; bool legal2 = false;
                const uint64_t ks = win_eval_at(wc, ws, d, win_desc(d), w6w, [&](int k) { return (k >= 0 && k < COLS) ? c[k] : 0xdeadbeefu; }, and_out, &full2, &legal2);
                if (ks != kw || full2 != full || legal2 != legal) ++bad;
                {   // K2sP starts the windows in the middle of a rotation: win_skip(c0) == c0 slides
                    WinSlide wk;
                    win_reset(wc, wk);
                    win_skip(wk, c0);
                    for (int i = 0; i < 4; ++i) if (wk.x[i] != ws.x[i]) ++bad;
                    for (int i = 0; i < 3; ++i) if (wk.r[i] != ws.r[i] || wk.t[i] != ws.t[i] || wk.d[i] != ws.d[i]) ++bad;
                }
                win_advance(ws);
            }
            if (legal != A.legal) { ++bad; continue; }
            if (!legal) { if (kw != 0ull) ++bad; continue; }
            if (full != A.full) { ++bad; continue; }
            if (full != 0u) { ++nclr; continue; }
            Heights Hn;
            const uint64_t k0 = w6_post(A, w6w, Hn);
            ++ncmp;
            if (k0 != kw) ++bad;
        }
    }
    *compared = ncmp; *clearing = nclr;
    return bad;
}
