/* coracle.c — plain-C CPU restatement of the CogWorks/Tetris-AI placement hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  The product (tetris_ai_b200/, libtetris_b200.so) never links, loads
 * or calls this file; tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it as
 * the checker / timed baseline.
 *
 * Parity status: PINNED BY REFERENCE OUTPUTS — checked against the golden vectors that
 * tests/golden/make_golden.py produced by running the unmodified reference (feature vectors,
 * best-move steps, whole-game traces, cross-entropy generations); see tests/test_oracle_golden.py.
 *
 * Deliberately naive: the board is a byte per cell (cells[r][c], r = 0 is the TOP row, exactly
 * the reference's numpy layout), moves are found with the reference's collide-and-lower loop and
 * every feature is computed from its definition with plain loops.  It shares no code, tables or
 * bit tricks with the CUDA kernels, so agreement between the two is meaningful.
 *
 * Citations are to /root/reference/cogworks/...
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -shared -fPIC -pthread)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ROWS 20
#define COLS 10
#define NFEAT 45

typedef struct { uint8_t c[ROWS][COLS]; } board_t;

/* ---- pieces: tetris/game.py:6-34 (rotation i = np.rot90(shape, k=-i), order I,T,L,J,O,Z,S) ---- */
typedef struct { int h, w; uint8_t m[4][4]; } shape_t;
static const char* BASE_ROWS[7][2] = {
    {"####", 0}, {"###", ".#."}, {"###", "#.."}, {"###", "..#"}, {"##", "##"}, {"##.", ".##"}, {".##", "##."}};
static const int NROT[7] = {2, 4, 4, 4, 1, 2, 2};
static shape_t SHAPES[7][4];
static int shapes_ready = 0;

static void rotate_cw(const shape_t* a, shape_t* out)
{   /* new[i][j] = old[h-1-j][i] */
    memset(out, 0, sizeof *out);
    out->h = a->w; out->w = a->h;
    for (int i = 0; i < out->h; ++i)
        for (int j = 0; j < out->w; ++j)
            out->m[i][j] = a->m[a->h - 1 - j][i];
}

static void init_shapes(void)
{
    if (shapes_ready) return;
    for (int p = 0; p < 7; ++p) {
        shape_t s; memset(&s, 0, sizeof s);
        s.h = BASE_ROWS[p][1] ? 2 : 1;
        s.w = (int)strlen(BASE_ROWS[p][0]);
        for (int r = 0; r < s.h; ++r)
            for (int c = 0; c < s.w; ++c) s.m[r][c] = BASE_ROWS[p][r][c] == '#';
        for (int r = 0; r < NROT[p]; ++r) {
            SHAPES[p][r] = s;
            shape_t t; rotate_cw(&s, &t); s = t;
        }
    }
    shapes_ready = 1;
}

/* ---- board: tetris/game.py:38-108 ---- */
static int col_height(const board_t* b, int c)          /* game.py:82-87 */
{
    for (int r = 0; r < ROWS; ++r) if (b->c[r][c]) return ROWS - r;
    return 0;
}

static int collides(const board_t* b, const shape_t* s, int row, int col)   /* game.py:69-75 */
{
    if (row < 0 || col < 0) return 1;
    if (row + s->h > ROWS || col + s->w > COLS) return 1;
    for (int i = 0; i < s->h; ++i)
        for (int j = 0; j < s->w; ++j)
            if (s->m[i][j] && b->c[row + i][col + j]) return 1;
    return 0;
}

/* simulator.py:3-13 move_drop: rot-major, col-minor; returns number of moves */
static int drop_moves(const board_t* b, int piece, int mv[34][3])
{
    int n = 0;
    for (int rot = 0; rot < NROT[piece]; ++rot) {
        const shape_t* s = &SHAPES[piece][rot];
        for (int col = 0; col <= COLS - s->w; ++col) {
            int highest = 0;
            for (int j = 0; j < s->w; ++j) { int h = col_height(b, col + j); if (h > highest) highest = h; }
            int row = ROWS - highest - s->h;
            if (!collides(b, s, row, col)) {
                while (!collides(b, s, row + 1, col)) ++row;
                mv[n][0] = rot; mv[n][1] = row; mv[n][2] = col; ++n;
            }
        }
    }
    return n;
}

#define MAX_MOVES 640

/* simulator.py:15-47 move_with_overhang_slide(move_drop): every hard-drop move, followed by the placements reached
 * by sliding sideways at the landed row under an overhang (right slides in ascending column order, then left slides
 * in descending column order), each lowered as far as it goes; a slide direction stops at the first blocked column. */
static int slide_moves(const board_t* b, int piece, int mv[MAX_MOVES][3])
{
    int base[34][3]; int nb = drop_moves(b, piece, base);
    int n = 0;
    for (int i = 0; i < nb; ++i) {
        int rot = base[i][0], row = base[i][1], col = base[i][2];
        const shape_t* s = &SHAPES[piece][rot];
        mv[n][0] = rot; mv[n][1] = row; mv[n][2] = col; ++n;
        if (col + s->w < COLS && row > ROWS - col_height(b, col + s->w)) {            /* :25 */
            for (int c = col + 1; c <= COLS - s->w; ++c) {                           /* :27 */
                if (collides(b, s, row, c)) break;
                int r = row;
                while (!collides(b, s, r + 1, c)) ++r;
                mv[n][0] = rot; mv[n][1] = r; mv[n][2] = c; ++n;
            }
        }
        if (col > 0 && row > ROWS - col_height(b, col - 1)) {                        /* :36 */
            for (int c = col - 1; c >= 0; --c) {                                     /* :38 */
                if (collides(b, s, row, c)) break;
                int r = row;
                while (!collides(b, s, r + 1, c)) ++r;
                mv[n][0] = rot; mv[n][1] = r; mv[n][2] = c; ++n;
            }
        }
    }
    return n;
}

/* simulator.py:49-85 move_with_wiggle / move_wiggle: depth-first flood fill of the positions the piece can reach by moving
 * down, left and right (in that order) from each start; a position where the piece cannot move down is a move.  `seen` is
 * shared by all starts of one piece, so the ORDER of the move list is the order of the recursion. */
typedef struct { uint8_t seen[4][ROWS + 2][COLS + 2]; int (*mv)[3]; int n; const board_t* b; int piece; } wiggle_t;

static void wiggle_rec(wiggle_t* wg, int rot, int row, int col)
{
    if (wg->seen[rot][row][col]) return;                                  /* :52 */
    wg->seen[rot][row][col] = 1;
    const shape_t* s = &SHAPES[wg->piece][rot];
    if (collides(wg->b, s, row, col)) return;                             /* :54 */
    if (collides(wg->b, s, row + 1, col)) {                               /* :55-57 resting: a move */
        wg->mv[wg->n][0] = rot; wg->mv[wg->n][1] = row; wg->mv[wg->n][2] = col; ++wg->n;
    } else {
        wiggle_rec(wg, rot, row + 1, col);                                /* :58-60 */
    }
    if (col > 0) wiggle_rec(wg, rot, row, col - 1);                       /* :63-64 */
    if (col + s->w <= COLS) wiggle_rec(wg, rot, row, col + 1);            /* :65-66 */
}

/* from_top != 0: move_wiggle (starts (rot, 0, col) for every rot, col); else move_with_wiggle(move_drop) */
static int wiggle_moves(const board_t* b, int piece, int from_top, int mv[MAX_MOVES][3])
{
    wiggle_t wg; memset(&wg, 0, sizeof wg);
    wg.mv = mv; wg.b = b; wg.piece = piece;
    if (from_top) {
        for (int rot = 0; rot < NROT[piece]; ++rot)
            for (int col = 0; col <= COLS - SHAPES[piece][rot].w; ++col) wiggle_rec(&wg, rot, 0, col);
    } else {
        int base[34][3]; int nb = drop_moves(b, piece, base);
        for (int i = 0; i < nb; ++i) wiggle_rec(&wg, base[i][0], base[i][1], base[i][2]);
    }
    return wg.n;
}

/* simulator.py:88-120 move_with_pressure: keep a move iff the time to key it in fits before the piece locks.
 * press = {avg_lat, resp_time, move_eff, two_but_rot}; level = state.level() of the state the move is generated from. */
static const int START_COL[7][4] = {{3, 5, 0, 0}, {4, 4, 4, 5}, {4, 4, 4, 5}, {4, 4, 4, 5}, {4, 0, 0, 0}, {4, 5, 0, 0}, {4, 5, 0, 0}};   /* I,T,L,J,O,Z,S :92-100 */
static int is_time(int piece, int rot, int row, int col, const double* press, int level)
{
    static const int BR[15][2] = {{29, 20}, {19, 30}, {16, 50}, {13, 70}, {10, 80}, {9, 100}, {8, 130}, {7, 220}, {6, 300},
                                  {5, 380}, {4, 470}, {3, 550}, {2, 630}, {1, 720}, {0, 800}};                          /* :110-112 */
    int d = START_COL[piece][rot] - col; if (d < 0) d = -d;
    int clicks = d + ((press[3] != 0.0 && rot == 3) ? 1 : rot);                                                        /* :104 */
    double t = (double)clicks * press[2];                                                                                /* :107, left to right */
    t = t * press[0];
    t = press[1] + t;
    int speed = 800;
    for (int i = 0; i < 15; ++i) if (BR[i][0] <= level) { speed = BR[i][1]; break; }                                    /* :113-114 */
    int limit = (ROWS - row) * speed;                                                                                    /* :116 */
    return t <= (double)limit;
}

/* move_set bit 0: overhang slides; bit 2: move_wiggle; bit 3: move_with_wiggle(move_drop); bit 1: pressure filter applied to
 * whatever the inner generator yields */
static int gen_moves(const board_t* b, int piece, int move_set, const double* press, int level, int mv[MAX_MOVES][3])
{
    int n;
    if (move_set & 4) n = wiggle_moves(b, piece, 1, mv);
    else if (move_set & 8) n = wiggle_moves(b, piece, 0, mv);
    else if (move_set & 1) n = slide_moves(b, piece, mv);
    else {
        int base[34][3]; n = drop_moves(b, piece, base);
        for (int i = 0; i < n; ++i) { mv[i][0] = base[i][0]; mv[i][1] = base[i][1]; mv[i][2] = base[i][2]; }
    }
    if ((move_set & 2) && press) {
        int m = 0;
        for (int i = 0; i < n; ++i)
            if (is_time(piece, mv[i][0], mv[i][1], mv[i][2], press, level)) { mv[m][0] = mv[i][0]; mv[m][1] = mv[i][1]; mv[m][2] = mv[i][2]; ++m; }
        n = m;
    }
    return n;
}

typedef struct {
    board_t board;          /* post-clear */
    int piece, rot, row, col;
    int k;                  /* rows cleared by this move */
    uint8_t gone[ROWS];     /* gone[r] = 1 iff pre-clear row r was full */
} placed_t;

/* game.py:166-180 State.future = copy, imprint (:62-67), full (:77-78), clear (:80-87) */
static void place(const board_t* b, int piece, int rot, int row, int col, placed_t* out)
{
    const shape_t* s = &SHAPES[piece][rot];
    board_t t = *b;
    for (int i = 0; i < s->h; ++i)
        for (int j = 0; j < s->w; ++j)
            if (s->m[i][j]) t.c[row + i][col + j] = 1;
    out->piece = piece; out->rot = rot; out->row = row; out->col = col; out->k = 0;
    for (int r = 0; r < ROWS; ++r) {
        int full = 1;
        for (int c = 0; c < COLS; ++c) if (!t.c[r][c]) { full = 0; break; }
        out->gone[r] = (uint8_t)full;
        out->k += full;
    }
    if (out->k) {                  /* delete full rows, insert empty rows at the top */
        board_t u; memset(&u, 0, sizeof u);
        int dst = ROWS - 1;
        for (int r = ROWS - 1; r >= 0; --r)
            if (!out->gone[r]) { memcpy(u.c[dst], t.c[r], COLS); --dst; }
        out->board = u;
    } else {
        out->board = t;
    }
}

/* ---- features: tetris/features.py:18-347; ids = source order ---- */
static void heights_of(const board_t* b, int h[COLS]) { for (int c = 0; c < COLS; ++c) h[c] = col_height(b, c); }

static int pits_of(const board_t* b, const int h[COLS])       /* features.py:146-160 */
{
    int n = 0;
    for (int c = 0; c < COLS; ++c)
        for (int r = ROWS - h[c] + 1; r < ROWS; ++r) if (!b->c[r][c]) ++n;
    return n;
}

static void features45(const board_t* prev, const placed_t* st, double f[NFEAT])
{
    const board_t* b = &st->board;
    int h[COLS], ph[COLS];
    heights_of(b, h); heights_of(prev, ph);
    int sum = 0, mx = h[0], mn = h[0];
    for (int c = 0; c < COLS; ++c) { sum += h[c]; if (h[c] > mx) mx = h[c]; if (h[c] < mn) mn = h[c]; }
    int psum = 0, pmx = ph[0];
    for (int c = 0; c < COLS; ++c) { psum += ph[c]; if (ph[c] > pmx) pmx = ph[c]; }
    double mean = (double)sum / (double)COLS;                  /* :24-26 */
    double pmean = (double)psum / (double)COLS;
    f[0] = mean; f[1] = mx; f[2] = mn; f[3] = sum;
    f[4] = (double)mx - mean;                                  /* :48-50 */
    f[5] = mean - (double)mn;                                  /* :54-56 */
    int k = st->k;
    f[6] = k; f[7] = (double)(k * (k + 1)) / 2.0; f[8] = k / 4;     /* :62-77 */
    int ct = 0, rt = 0;
    for (int r = 0; r + 1 < ROWS; ++r) for (int c = 0; c < COLS; ++c) ct += b->c[r][c] != b->c[r + 1][c];   /* :83-85 */
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c + 1 < COLS; ++c) rt += b->c[r][c] != b->c[r][c + 1];   /* :89-91 */
    f[9] = ct; f[10] = rt; f[11] = ct + rt;
    /* wells :106-137 */
    int wsum = 0, deep = 0, wmax = 0; double cuml = 0.0;
    for (int c = 0; c < COLS; ++c) {
        int l = c == 0 ? ROWS : h[c - 1], r = c == COLS - 1 ? ROWS : h[c + 1];
        int w = (l < r ? l : r) - h[c]; if (w < 0) w = 0;
        wsum += w; cuml += (double)(w * (w + 1)) / 2.0; if (w > 1) deep += w; if (w > wmax) wmax = w;
    }
    f[12] = wsum; f[13] = cuml; f[14] = deep; f[15] = wmax;
    /* pits :146-194 */
    int pits = 0, pit_rows = 0, lumped = 0, depth = 0;
    uint8_t row_has[ROWS]; memset(row_has, 0, sizeof row_has);
    for (int c = 0; c < COLS; ++c) {
        int top = ROWS - h[c];
        for (int r = top + 1; r < ROWS; ++r) {
            if (b->c[r][c]) continue;
            ++pits; row_has[r] = 1;
            int above_is_pit = (r - 1 >= top + 1) && !b->c[r - 1][c];
            if (!above_is_pit) ++lumped;
            for (int q = top; q < r; ++q) if (b->c[q][c]) ++depth;     /* non-pit cells above it */
        }
    }
    for (int r = 0; r < ROWS; ++r) pit_rows += row_has[r];
    f[16] = pits; f[17] = pit_rows; f[18] = lumped; f[19] = depth;
    f[20] = pits != 0 ? (double)depth / (double)pits : 0.0;
    /* column differences :201-233 */
    int all = 0, dmax = h[1] - h[0], jag = 0;
    for (int n = 1; n < COLS; ++n) {
        int d = h[n] - h[n - 1];
        f[20 + n] = d; all += d; if (d > dmax) dmax = d; jag += d < 0 ? -d : d;
    }
    f[30] = all; f[31] = dmax; f[32] = jag;
    /* deltas vs previous state :239-257 */
    f[33] = mx - pmx; f[34] = sum - psum; f[35] = mean - pmean; f[36] = pits - pits_of(prev, ph);
    /* landing_height :265-267 (pre-clear coordinates) */
    f[37] = ROWS - (st->row + SHAPES[st->piece][st->rot].h);
    /* pattern_div :271-273: distinct difference vectors between adjacent columns */
    int distinct = 0;
    for (int j = 0; j + 1 < COLS; ++j) {
        int seen = 0;
        for (int i = 0; i < j && !seen; ++i) {
            int same = 1;
            for (int r = 0; r < ROWS; ++r)
                if ((int)b->c[r][j + 1] - (int)b->c[r][j] != (int)b->c[r][i + 1] - (int)b->c[r][i]) { same = 0; break; }
            seen = same;
        }
        distinct += !seen;
    }
    f[38] = distinct;
    /* nine_filled :279-292, tetris_progress :297-319 */
    int nine = 0, hole_of[ROWS];
    for (int r = 0; r < ROWS; ++r) {
        int empties = 0, e = -1;
        for (int c = 0; c < COLS; ++c) if (!b->c[r][c]) { ++empties; e = c; }
        hole_of[r] = empties == 1 ? e : -1;
        nine += empties == 1;
    }
    f[39] = nine;
    int progress = 0, ncov = 0; uint8_t covered[COLS]; memset(covered, 0, sizeof covered);
    for (int r = 0; r < ROWS; ++r) {
        for (int c = 0; c < COLS; ++c) if (b->c[r][c] && !covered[c]) { covered[c] = 1; ++ncov; }
        if (hole_of[r] >= 0) {
            if (!covered[hole_of[r]]) ++progress; else break;
        } else if (ncov < COLS) {
            progress = 0;
        } else break;
    }
    f[40] = progress;
    int cells = 0, weighted = 0;
    for (int r = 0; r < ROWS; ++r) { int n = 0; for (int c = 0; c < COLS; ++c) n += b->c[r][c]; cells += n; weighted += n * (ROWS - r); }
    f[41] = cells; f[42] = weighted;                           /* :323-331 */
    /* eroded_cells :335-340: piece cells that sat in cleared rows (pre-clear coordinates) */
    const shape_t* s = &SHAPES[st->piece][st->rot];
    int er = 0;
    for (int r = 0; r < ROWS; ++r)
        if (st->gone[r]) { int i = r - st->row; if (i >= 0 && i < s->h) for (int j = 0; j < s->w; ++j) er += s->m[i][j]; }
    f[43] = er; f[44] = (double)(er * (er + 1)) / 2.0;         /* :345-347 */
}

/* canonical scorer: acc = 0.0; acc += float(f) * w, in the caller's order */
static double linear_score(const double f[NFEAT], const int* fids, const double* w, int F)
{
    double acc = 0.0;
    for (int i = 0; i < F; ++i) { double p = f[fids[i]] * w[i]; acc = acc + p; }
    return acc;
}

/* ---- public API (all pointers caller-owned) ------------------------------------------------ */
static void rows_to_board(const uint16_t rows[ROWS], board_t* b)
{
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < COLS; ++c) b->c[r][c] = (rows[r] >> c) & 1;
}
static void board_to_rows(const board_t* b, uint16_t rows[ROWS])
{
    for (int r = 0; r < ROWS; ++r) { uint16_t v = 0; for (int c = 0; c < COLS; ++c) v |= (uint16_t)(b->c[r][c] << c); rows[r] = v; }
}

/* legal moves of `piece` on a board given as 20 row masks; moves[i] = {rot,row,col}; returns count */
int co_drop_moves(const uint16_t rows[ROWS], int piece, int* moves)
{
    init_shapes();
    board_t b; rows_to_board(rows, &b);
    int mv[34][3]; int n = drop_moves(&b, piece, mv);
    for (int i = 0; i < n; ++i) { moves[3 * i] = mv[i][0]; moves[3 * i + 1] = mv[i][1]; moves[3 * i + 2] = mv[i][2]; }
    return n;
}

/* move list of move_set (0 = move_drop, 1 = move_with_overhang_slide(move_drop)); moves[i] = {rot,row,col} */
int co_moves(const uint16_t rows[ROWS], int piece, int move_set, int* moves)
{
    init_shapes();
    board_t b; rows_to_board(rows, &b);
    static __thread int mv[MAX_MOVES][3];
    int n = gen_moves(&b, piece, move_set & 13, 0, 0, mv);
    for (int i = 0; i < n; ++i) { moves[3 * i] = mv[i][0]; moves[3 * i + 1] = mv[i][1]; moves[3 * i + 2] = mv[i][2]; }
    return n;
}

/* features of prev.future(piece, rot, <drop row>, col); returns 0 if that slot is illegal, else 1 */
int co_eval_placement(const uint16_t prev_rows[ROWS], int piece, int rot, int col,
                      double* feats45, uint16_t* post_rows, int* row_out, int* k_out)
{
    init_shapes();
    board_t b; rows_to_board(prev_rows, &b);
    int mv[34][3]; int n = drop_moves(&b, piece, mv);
    for (int i = 0; i < n; ++i)
        if (mv[i][0] == rot && mv[i][2] == col) {
            placed_t st; place(&b, piece, rot, mv[i][1], col, &st);
            if (feats45) features45(&b, &st, feats45);
            if (post_rows) board_to_rows(&st.board, post_rows);
            if (row_out) *row_out = mv[i][1];
            if (k_out) *k_out = st.k;
            return 1;
        }
    return 0;
}

typedef struct { int rot, row, col, k; double score; int n_scored; placed_t st; } choice_t;

/* One simulate() iteration (simulator.py:156-180) with move_drop and policy_best(linear, first).
 * next_piece < 0 => only one visible piece.  Returns 0 when there is no move (game over). */
static int choose(const board_t* b, int lines, int piece, int next_piece, int lookahead, int move_set, const double* press,
                  const int* fids, const double* w, int F, choice_t* out)
{
    static __thread int mv1[MAX_MOVES][3];
    int n1 = gen_moves(b, piece, move_set, press, lines / 10, mv1);
    if (n1 == 0) return 0;
    static __thread placed_t l1[MAX_MOVES];
    for (int i = 0; i < n1; ++i) place(b, piece, mv1[i][0], mv1[i][1], mv1[i][2], &l1[i]);
    double f[NFEAT];
    int have = 0; double best = 0.0; int best_i = -1; int scored = 0;
    if (lookahead >= 2 && next_piece >= 0) {
        for (int i = 0; i < n1; ++i) {                       /* layer 2 in enumeration order :162-167 */
            static __thread int mv2[MAX_MOVES][3];
            int n2 = gen_moves(&l1[i].board, next_piece, move_set, press, (lines + l1[i].k) / 10, mv2);
            for (int j = 0; j < n2; ++j) {
                placed_t leaf; place(&l1[i].board, next_piece, mv2[j][0], mv2[j][1], mv2[j][2], &leaf);
                features45(&l1[i].board, &leaf, f);
                double s = linear_score(f, fids, w, F); ++scored;
                if (!have || best < s) { have = 1; best = s; best_i = i; }   /* all_max :125-140, first kept */
            }
        }
    }
    if (!have) {                                             /* farthest non-empty pool is layer 1 :170-174 */
        scored = 0;
        for (int i = 0; i < n1; ++i) {
            features45(b, &l1[i], f);
            double s = linear_score(f, fids, w, F); ++scored;
            if (!have || best < s) { have = 1; best = s; best_i = i; }
        }
    }
    out->rot = mv1[best_i][0]; out->row = mv1[best_i][1]; out->col = mv1[best_i][2];
    out->k = l1[best_i].k; out->score = best; out->n_scored = scored; out->st = l1[best_i];
    return 1;
}

/* best move for one board.  out_move = {rot,row,col,k}; returns scored-leaf count (0 = no move) */
int co_best_move(const uint16_t rows[ROWS], int piece, int next_piece, int lookahead,
                 const int* fids, const double* w, int F, int* out_move, double* out_score, int move_set)
{
    init_shapes();
    board_t b; rows_to_board(rows, &b);
    choice_t ch;
    if (!choose(&b, 0, piece, next_piece, lookahead, move_set & 1, 0, fids, w, F, &ch)) return 0;
    out_move[0] = ch.rot; out_move[1] = ch.row; out_move[2] = ch.col; out_move[3] = ch.k;
    if (out_score) *out_score = ch.score;
    return ch.n_scored;
}

/* K4 twin: counter-based piece stream, identical to tb_hashed_piece() / pyoracle.hashed_piece() */
static int hashed_piece(uint64_t seed, uint64_t game, uint64_t t)
{
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + game * 0xD1B54A32D192ED03ull + t * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (int)(((x >> 32) * 7ull) >> 32);
}
int co_hashed_piece(uint64_t seed, uint64_t game, uint64_t t) { return hashed_piece(seed, game, t); }

typedef struct {
    uint32_t pieces, lines, hist[4];
    uint64_t score, placements;
} game_result_t;

/* One whole game (simulator.py:152-189).  pieces == NULL => hashed stream (seed, stream_id).
 * trace (optional) receives 4 bytes {rot,row,col,k} per move. */
static void play_game(const int* fids, const double* w, int F, const uint8_t* pieces, uint32_t T,
                      uint64_t seed, uint64_t stream_id, int lookahead, int move_set, const double* press, const uint16_t* init_rows,
                      uint32_t init_lines, game_result_t* res, uint8_t* trace, uint16_t* final_rows)
{
    static const uint32_t POINTS[5] = {0, 40, 100, 300, 1200};         /* game.py:159 */
    board_t b; memset(&b, 0, sizeof b);
    if (init_rows) rows_to_board(init_rows, &b);
    memset(res, 0, sizeof *res);
    res->lines = init_lines;              /* a continuing game: State.lines_cleared() of the states before it (game.py:150-157) */
    for (uint32_t t = 0; t < T; ++t) {
        int p  = pieces ? pieces[t] : hashed_piece(seed, stream_id, t);
        int nx = (lookahead >= 2 && t + 1 < T) ? (pieces ? pieces[t + 1] : hashed_piece(seed, stream_id, t + 1)) : -1;
        choice_t ch;
        if (p < 0 || p > 6) break;            /* not a piece: the sequence ends here (harness convention, matches the library) */
        if (nx > 6) nx = -1;
        if (!choose(&b, (int)res->lines, p, nx, lookahead, move_set, press, fids, w, F, &ch)) break;
        b = ch.st.board;
        res->placements += (uint64_t)ch.n_scored;
        res->pieces += 1;
        res->lines += (uint32_t)ch.k;
        if (ch.k) res->hist[ch.k - 1] += 1;
        res->score += (uint64_t)POINTS[ch.k] * (uint64_t)(res->lines / 10 + 1);   /* level after the move */
        if (trace) { trace[4 * t] = (uint8_t)ch.rot; trace[4 * t + 1] = (uint8_t)ch.row; trace[4 * t + 2] = (uint8_t)ch.col; trace[4 * t + 3] = (uint8_t)ch.k; }
    }
    if (final_rows) board_to_rows(&b, final_rows);
}

typedef struct {
    const int* fids; int F; const double* weights; int C; int G;
    const uint8_t* pieces; int S; uint32_t T; const int32_t* seq_of_game; uint64_t seed; int lookahead;
    game_result_t* results; uint8_t* trace; volatile int* next; int total; int move_set; const double* press;
} job_t;

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    for (;;) {
        int g = __sync_fetch_and_add(j->next, 1);
        if (g >= j->total) break;
        int cand = g / j->G, gi = g % j->G;
        int seq = j->seq_of_game ? j->seq_of_game[g] : gi;
        const uint8_t* ps = j->pieces ? j->pieces + (size_t)seq * j->T : 0;
        play_game(j->fids, j->weights + (size_t)cand * j->F, j->F, ps, j->T, j->seed, (uint64_t)seq, j->lookahead, j->move_set, j->press, 0,
                  0, &j->results[g], j->trace ? j->trace + (size_t)g * j->T * 4 : 0, 0);
    }
    return 0;
}

/* C candidates x G games.  weights f64[C][F]; pieces u8[S][T] or NULL (hashed stream, stream id =
 * seq_of_game[g] or g % G); results: 8 x u64 per game {pieces, lines, h1, h2, h3, h4, score, placements}.
 * threads <= 0 => 1. */
int co_play_games(const int* fids, int F, const double* weights, int C, int G,
                  const uint8_t* pieces, int S, uint32_t T, const int32_t* seq_of_game, uint64_t seed,
                  int lookahead, int threads, uint64_t* results, uint8_t* trace, int move_set, const double* press)
{
    init_shapes();
    (void)S;
    int total = C * G;
    game_result_t* tmp = (game_result_t*)calloc((size_t)total, sizeof(game_result_t));
    if (!tmp) return -1;
    volatile int next = 0;
    job_t j = {fids, F, weights, C, G, pieces, S, T, seq_of_game, seed, lookahead, tmp, trace, &next, total, move_set, press};
    if (threads <= 1) worker(&j);
    else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int i = 0; i < threads; ++i) pthread_create(&th[i], 0, worker, &j);
        for (int i = 0; i < threads; ++i) pthread_join(th[i], 0);
        free(th);
    }
    for (int g = 0; g < total; ++g) {
        uint64_t* r = results + (size_t)g * 8;
        r[0] = tmp[g].pieces; r[1] = tmp[g].lines;
        r[2] = tmp[g].hist[0]; r[3] = tmp[g].hist[1]; r[4] = tmp[g].hist[2]; r[5] = tmp[g].hist[3];
        r[6] = tmp[g].score; r[7] = tmp[g].placements;
    }
    free(tmp);
    return 0;
}

/* single game with optional initial board / final board (used by the chunked simulate() tests) */
int co_play_one(const int* fids, int F, const double* w, const uint8_t* pieces, uint32_t T, int lookahead,
                const uint16_t* init_rows, uint64_t* result8, uint8_t* trace, uint16_t* final_rows, int move_set, const double* press)
{
    init_shapes();
    game_result_t r;
    play_game(fids, w, F, pieces, T, 0, 0, lookahead, move_set, press, init_rows, 0, &r, trace, final_rows);
    result8[0] = r.pieces; result8[1] = r.lines; result8[2] = r.hist[0]; result8[3] = r.hist[1];
    result8[4] = r.hist[2]; result8[5] = r.hist[3]; result8[6] = r.score; result8[7] = r.placements;
    return 0;
}

/* same, for a game that already cleared init_lines lines (level of the score and of the pressure filter) */
int co_play_one_ex(const int* fids, int F, const double* w, const uint8_t* pieces, uint32_t T, int lookahead,
                   const uint16_t* init_rows, uint32_t init_lines, uint64_t* result8, uint8_t* trace, uint16_t* final_rows,
                   int move_set, const double* press)
{
    init_shapes();
    game_result_t r;
    play_game(fids, w, F, pieces, T, 0, 0, lookahead, move_set, press, init_rows, init_lines, &r, trace, final_rows);
    result8[0] = r.pieces; result8[1] = r.lines; result8[2] = r.hist[0]; result8[3] = r.hist[1];
    result8[4] = r.hist[2]; result8[5] = r.hist[3]; result8[6] = r.score; result8[7] = r.placements;
    return 0;
}
