// tetris_core.cuh — column-bitmask Tetris placement core (sm_100a), shared by every kernel.
//
// What it replaces in the reference (paths under /root/reference/cogworks/):
//   tetris/simulator.py:3-13    move_drop        -> tb::place_slot (closed-form landing row, legality)
//   tetris/game.py:62-87,166-180 imprint/full/clear/future -> tb::place_slot (imprint, full-row AND, compress)
//   tetris/features.py:18-347   the 45 features  -> tb::eval_features (popc / clz bit formulas)
//   feature.py:50-55 + the canonical linear scorer -> tb::score_features
//
// Data layout: a board is 10 column words; bit i of col[k] is the cell of column k at height i
// above the floor (i = 0 is the reference's row 19, i = 19 its row 0).  Heights are 32 - clz.
// Everything is __host__ __device__ so that tests/emu can run the identical code on the CPU
// against the oracles; the product only ever runs it on the GPU.
#pragma once
#if defined(__CUDACC_RTC__)      // run-time compilation (NVRTC) has no host headers
typedef unsigned char uint8_t; typedef unsigned short uint16_t; typedef unsigned int uint32_t; typedef int int32_t;
typedef unsigned long long uint64_t; typedef long long int64_t;
#else
#include <stdint.h>
#endif
#include "tetris_tables.h"

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#define TB_D  __device__ __forceinline__
#define TB_CX __host__ __device__ constexpr
#else
#define TB_HD inline
#define TB_CX constexpr
#endif

namespace tb {

constexpr int ROWS = 20;
constexpr int COLS = 10;
constexpr int NFEAT = 45;
constexpr uint32_t COLMASK = 0xFFFFFu;

// feature ids = source order of tetris/features.py
enum : int {
    F_MEAN_HT = 0, F_MAX_HT, F_MIN_HT, F_ALL_HT, F_MAX_HT_DIFF, F_MIN_HT_DIFF, F_CLEARED, F_CUML_CLEARED,
    F_TETRIS, F_COL_TRANS, F_ROW_TRANS, F_ALL_TRANS, F_WELLS, F_CUML_WELLS, F_DEEP_WELLS, F_MAX_WELL,
    F_PITS, F_PIT_ROWS, F_LUMPED_PITS, F_PIT_DEPTH, F_MEAN_PIT_DEPTH,
    F_CD_1, F_CD_2, F_CD_3, F_CD_4, F_CD_5, F_CD_6, F_CD_7, F_CD_8, F_CD_9,
    F_ALL_DIFFS, F_MAX_DIFFS, F_JAGGEDNESS, F_D_MAX_HT, F_D_ALL_HT, F_D_MEAN_HT, F_D_PITS,
    F_LANDING_HEIGHT, F_PATTERN_DIV, F_NINE_FILLED, F_TETRIS_PROGRESS, F_FULL_CELLS, F_WEIGHTED_CELLS,
    F_ERODED_CELLS, F_CUML_ERODED
};
TB_CX uint64_t bit(int id) { return 1ull << id; }
constexpr uint64_t ALL45_MASK = (1ull << NFEAT) - 1;
// the "w6" set of SURVEY.md §8c, in its summation order
constexpr uint64_t W6_MASK = bit(F_LANDING_HEIGHT) | bit(F_ERODED_CELLS) | bit(F_ROW_TRANS) | bit(F_COL_TRANS) |
                             bit(F_PITS) | bit(F_CUML_WELLS);

// ---------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------
struct SlotDesc { uint32_t w[12]; };         // layout: tools/gen_tables.py
struct alignas(16) Tables {                  // 16-byte aligned and padded: copied into shared memory with 128-bit loads
    SlotDesc slot[7][TB_MAX_SLOTS];
    int nslots[7];
};
#if !defined(__CUDA_ARCH__)
static const Tables& host_tables()
{
    static const Tables t = { TB_SLOT_TABLE_INIT, TB_NSLOTS_INIT };
    return t;
}
#endif

// ---------------------------------------------------------------------------------------------
// small intrinsics with host twins
// ---------------------------------------------------------------------------------------------
TB_HD int popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
TB_HD int fls(uint32_t x)                    // 32 - clz(x); 0 for x == 0  (column height)
{
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int)x);
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}
TB_HD int ctz(uint32_t x)                    // x != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
TB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)      // (hi:lo) >> sh, sh in 0..31
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
TB_HD uint32_t byte_of(uint32_t v, int j)    // j compile-time
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x4440 | j);
#else
    return (v >> (8 * j)) & 0xFFu;
#endif
}
// fp64 with explicit rounding per operation: never contracted into FMA
TB_HD double dmul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
TB_HD double dadd(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
TB_HD double dsub(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    volatile double r = a - b; return r;
#endif
}
TB_HD double ddiv(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}
TB_HD double i2d(int x)                      // exact int -> double without the quarter-rate I2F pipe
{
#if defined(__CUDA_ARCH__)
    // 2^52 + 2^31 + x is exactly representable; subtracting the bias is exact.
    return __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)), 4503601774854144.0);
#else
    return (double)x;
#endif
}

TB_HD double u2d(int x)                      // exact conversion of a NON-NEGATIVE int (one DADD, no integer op)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(__hiloint2double(0x43300000, x), 4503599627370496.0);      // (2^52 + x) - 2^52
#else
    return (double)x;
#endif
}

// ---------------------------------------------------------------------------------------------
// K4: counter-based piece stream (bit-identical twins: oracle/pyoracle.py hashed_piece,
// oracle/coracle.c hashed_piece)
// ---------------------------------------------------------------------------------------------
TB_HD int hashed_piece(uint64_t seed, uint64_t stream, uint64_t t)
{
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull + t * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (int)(((x >> 32) * 7ull) >> 32);
}

// ---------------------------------------------------------------------------------------------
// board helpers
// ---------------------------------------------------------------------------------------------
struct Heights { uint32_t H0, H1, H2; };     // byte k of (H0,H1,H2) = height of column k; bytes 10,11 = 0

TB_HD Heights pack_heights(const uint32_t c[COLS])
{
    Heights h;
    h.H0 = (uint32_t)fls(c[0]) | (uint32_t)fls(c[1]) << 8 | (uint32_t)fls(c[2]) << 16 | (uint32_t)fls(c[3]) << 24;
    h.H1 = (uint32_t)fls(c[4]) | (uint32_t)fls(c[5]) << 8 | (uint32_t)fls(c[6]) << 16 | (uint32_t)fls(c[7]) << 24;
    h.H2 = (uint32_t)fls(c[8]) | (uint32_t)fls(c[9]) << 8;
    return h;
}

// Cached statistics of a board: what the d_* features need from the previous state
// (features.py:239-257) plus what lets a placement WITHOUT a line clear be evaluated without
// clz / popc over the whole board (column heights, filled cells, sum of the height indices of
// the filled cells).
struct PrevStat { int max_ht, all_ht, pits, cells, possum; int gh[COLS]; };

TB_HD int possum_of(const uint32_t c[COLS])          // sum over filled cells of their height index
{
    int p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) {
        p0 += popc(c[k] & 0xAAAAAu); p1 += popc(c[k] & 0xCCCCCu); p2 += popc(c[k] & 0x0F0F0u);
        p3 += popc(c[k] & 0x0FF00u); p4 += popc(c[k] & 0xF0000u);
    }
    return p0 + 2 * p1 + 4 * p2 + 8 * p3 + 16 * p4;
}

TB_HD PrevStat prev_stat(const uint32_t c[COLS])
{
    PrevStat s; s.max_ht = 0; s.all_ht = 0; s.cells = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) { const int h = fls(c[k]); s.gh[k] = h; s.all_ht += h; s.max_ht = h > s.max_ht ? h : s.max_ht; s.cells += popc(c[k]); }
    s.pits = s.all_ht - s.cells;
    s.possum = possum_of(c);
    return s;
}

TB_HD Heights pack_gh(const int gh[COLS])
{
    Heights h;
    h.H0 = (uint32_t)gh[0] | (uint32_t)gh[1] << 8 | (uint32_t)gh[2] << 16 | (uint32_t)gh[3] << 24;
    h.H1 = (uint32_t)gh[4] | (uint32_t)gh[5] << 8 | (uint32_t)gh[6] << 16 | (uint32_t)gh[7] << 24;
    h.H2 = (uint32_t)gh[8] | (uint32_t)gh[9] << 8;
    return h;
}

// reference row-mask layout (u16[20], row 0 = top, bit c = column c) <-> column words
TB_HD void rows_to_cols(const uint16_t rows[ROWS], uint32_t c[COLS])
{
#pragma unroll
    for (int k = 0; k < COLS; ++k) c[k] = 0;
    for (int r = 0; r < ROWS; ++r) {
        uint32_t v = rows[r];
#pragma unroll
        for (int k = 0; k < COLS; ++k) c[k] |= ((v >> k) & 1u) << (ROWS - 1 - r);
    }
}
TB_HD void cols_to_rows(const uint32_t c[COLS], uint16_t rows[ROWS])
{
    for (int r = 0; r < ROWS; ++r) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < COLS; ++k) v |= ((c[k] >> (ROWS - 1 - r)) & 1u) << k;
        rows[r] = (uint16_t)v;
    }
}

// ---------------------------------------------------------------------------------------------
// one placement: drop + imprint + full rows + clear
// ---------------------------------------------------------------------------------------------
struct Placement {
    uint32_t n[COLS];      // post-clear board
    int y;                 // height of the piece's bottom bbox row above the floor (pre-clear) = landing_height
    int k;                 // rows cleared
    int eroded;            // piece cells that sat in cleared rows
    int rot, c0, ph;
};

// Straight-line part of a placement: landing row, legality, imprint, full-row mask.  Returns the
// legality (move_drop's hover row >= 0, simulator.py:8-10); *full receives the mask of full rows
// (bit i = height index i), still present in p.n.
TB_HD bool place_pre(const uint32_t c[COLS], const Heights& H, const SlotDesc& d, Placement& p, uint32_t* full)
{
    const uint32_t misc = d.w[0];
    const uint32_t sel = d.w[3];
    // 4-byte window of heights starting at column c0
    const uint32_t pair = sel >> 8;
    const uint32_t lo = pair == 0 ? H.H0 : (pair == 1 ? H.H1 : H.H2);
    const uint32_t hi = pair == 0 ? H.H1 : (pair == 1 ? H.H2 : 0u);
    const uint32_t hw = funnel_r(lo, hi, sel & 31u);
    // legal iff max_j h[c0+j] + ph <= ROWS
    const bool legal = ((hw + d.w[2]) & 0x80808080u) == 0u;
    // landing: y = max_j (h[c0+j] - bottom_j); biased by 32 so that the byte arithmetic stays unsigned
    const uint32_t v = hw + d.w[1];
    uint32_t m01 = byte_of(v, 0), b1 = byte_of(v, 1), b2 = byte_of(v, 2), b3 = v >> 24;
    m01 = m01 > b1 ? m01 : b1; b2 = b2 > b3 ? b2 : b3; m01 = m01 > b2 ? m01 : b2;
    const int y = (int)m01 - 32;
    p.y = y; p.c0 = (int)(misc & 15u); p.rot = (int)((misc >> 4) & 3u); p.ph = (int)((misc >> 9) & 7u);
    // imprint: piece cells never overlap board cells, so OR == ADD == one multiply-add per column
    const uint32_t one_y = 1u << (y & 31);
    p.n[0] = byte_of(d.w[4], 0) * one_y + c[0];
    p.n[1] = byte_of(d.w[4], 1) * one_y + c[1];
    p.n[2] = byte_of(d.w[4], 2) * one_y + c[2];
    p.n[3] = (d.w[4] >> 24)     * one_y + c[3];
    p.n[4] = byte_of(d.w[5], 0) * one_y + c[4];
    p.n[5] = byte_of(d.w[5], 1) * one_y + c[5];
    p.n[6] = byte_of(d.w[5], 2) * one_y + c[6];
    p.n[7] = (d.w[5] >> 24)     * one_y + c[7];
    p.n[8] = byte_of(d.w[6], 0) * one_y + c[8];
    p.n[9] = byte_of(d.w[6], 1) * one_y + c[9];
    // full rows (game.py:77-78)
    *full = (p.n[0] & p.n[1] & p.n[2]) & (p.n[3] & p.n[4] & p.n[5]) & (p.n[6] & p.n[7] & p.n[8]) & p.n[9];
    p.k = 0; p.eroded = 0;
    return legal;
}

// Rare part: remove the full rows `f` (non-zero) from p.n, count them and the eroded piece cells.
TB_HD void place_clear(const SlotDesc& d, Placement& p, uint32_t f)
{
    p.k = popc(f);
    // eroded_cells (features.py:335-340): piece cells in the cleared rows, pre-clear coordinates
    const uint32_t fr = f >> p.y, rc = d.w[0] >> 12;
    p.eroded = (int)(((fr & 1u) ? (rc & 7u) : 0u) + ((fr & 2u) ? ((rc >> 3) & 7u) : 0u) +
                     ((fr & 4u) ? ((rc >> 6) & 7u) : 0u) + ((fr & 8u) ? ((rc >> 9) & 7u) : 0u));
    // clear (game.py:80-87): drop each full row, everything above moves down one
    do {
        const uint32_t low = (1u << ctz(f)) - 1u;
#pragma unroll
        for (int k = 0; k < COLS; ++k) p.n[k] = (p.n[k] & low) | ((p.n[k] >> 1) & ~low);
        f = (f & low) | ((f >> 1) & ~low);
    } while (f != 0u);
}

// ---------------------------------------------------------------------------------------------
// SURVEY.md §8f rank 2 — move_with_overhang_slide (simulator.py:15-47): placements reached by sliding the
// piece sideways at its landed row under an overhang.  These are not hard drops, so they need a real
// collision test against the board columns and an imprint at a given height.
// ---------------------------------------------------------------------------------------------
TB_HD int height_at(const Heights& H, int k)                 // k in 0..9 (run-time index into the packed heights)
{
    const uint32_t wsel = (uint32_t)k >> 2;
    const uint32_t word = wsel == 0 ? H.H0 : (wsel == 1 ? H.H1 : H.H2);
    return (int)((word >> (8u * ((uint32_t)k & 3u))) & 255u);
}

// does the piece of slot d, with its bbox bottom at height y, lie inside the board and on empty cells?
TB_HD bool fits_at(const uint32_t c[COLS], const SlotDesc& d, int y)
{
    const int ph = (int)((d.w[0] >> 9) & 7u);
    if (y < 0 || y + ph > ROWS) return false;
    uint32_t o = 0;
    o |= (byte_of(d.w[4], 0) << y) & c[0]; o |= (byte_of(d.w[4], 1) << y) & c[1];
    o |= (byte_of(d.w[4], 2) << y) & c[2]; o |= ((d.w[4] >> 24) << y) & c[3];
    o |= (byte_of(d.w[5], 0) << y) & c[4]; o |= (byte_of(d.w[5], 1) << y) & c[5];
    o |= (byte_of(d.w[5], 2) << y) & c[6]; o |= ((d.w[5] >> 24) << y) & c[7];
    o |= (byte_of(d.w[6], 0) << y) & c[8]; o |= (byte_of(d.w[6], 1) << y) & c[9];
    return o == 0u;
}

// imprint the piece of slot d with its bbox bottom at height y (caller guarantees fits_at); *full = full-row mask
TB_HD void place_at(const uint32_t c[COLS], const SlotDesc& d, int y, Placement& p, uint32_t* full)
{
    const uint32_t misc = d.w[0];
    p.y = y; p.c0 = (int)(misc & 15u); p.rot = (int)((misc >> 4) & 3u); p.ph = (int)((misc >> 9) & 7u);
    const uint32_t one_y = 1u << (y & 31);
    p.n[0] = byte_of(d.w[4], 0) * one_y + c[0]; p.n[1] = byte_of(d.w[4], 1) * one_y + c[1];
    p.n[2] = byte_of(d.w[4], 2) * one_y + c[2]; p.n[3] = (d.w[4] >> 24)     * one_y + c[3];
    p.n[4] = byte_of(d.w[5], 0) * one_y + c[4]; p.n[5] = byte_of(d.w[5], 1) * one_y + c[5];
    p.n[6] = byte_of(d.w[5], 2) * one_y + c[6]; p.n[7] = (d.w[5] >> 24)     * one_y + c[7];
    p.n[8] = byte_of(d.w[6], 0) * one_y + c[8]; p.n[9] = byte_of(d.w[6], 1) * one_y + c[9];
    *full = (p.n[0] & p.n[1] & p.n[2]) & (p.n[3] & p.n[4] & p.n[5]) & (p.n[6] & p.n[7] & p.n[8]) & p.n[9];
    p.k = 0; p.eroded = 0;
}

// Enumerate the slide placements derived from the hard-drop move of slot s0 (descriptor d0, landed at height
// y0) in the reference's order: right slides by ascending column (:25-33), then left slides by descending
// column (:36-45); a direction stops at the first column where the piece does not fit at the landed row.
// fetch(s) returns the descriptor of slot s (slots of one rotation are consecutive, column-minor);
// emit(d, s, y, j) receives each slide placement with its 1-based index j within this base move.
template <class Fetch, class Emit>
TB_HD void for_each_slide(const uint32_t c[COLS], const Heights& H, const SlotDesc& d0, int s0, int y0, Fetch fetch, Emit emit)
{
    const int c0 = (int)(d0.w[0] & 15u), w = (int)((d0.w[0] >> 6) & 7u), ph = (int)((d0.w[0] >> 9) & 7u);
    int j = 1;
    if (c0 + w < COLS && height_at(H, c0 + w) > y0 + ph) {
        for (int cc = c0 + 1; cc <= COLS - w; ++cc) {
            const SlotDesc d = fetch(s0 + (cc - c0));
            if (!fits_at(c, d, y0)) break;
            int yy = y0;
            while (fits_at(c, d, yy - 1)) --yy;
            emit(d, s0 + (cc - c0), yy, j++);
        }
    }
    if (c0 > 0 && height_at(H, c0 - 1) > y0 + ph) {
        for (int cc = c0 - 1; cc >= 0; --cc) {
            const SlotDesc d = fetch(s0 - (c0 - cc));
            if (!fits_at(c, d, y0)) break;
            int yy = y0;
            while (fits_at(c, d, yy - 1)) --yy;
            emit(d, s0 - (c0 - cc), yy, j++);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SURVEY.md §8f rank 3 — move_with_pressure (simulator.py:88-120): a move survives iff the time to key it in
// (response time + key presses x efficiency x latency) fits before the piece locks at the current level's speed.
// ---------------------------------------------------------------------------------------------
struct Pressure { int on; int two_but_rot; double avg_lat, resp_time, move_eff; };

TB_HD int speed_of_level(int level)                           // the descending bracket table of simulator.py:110-114
{
    return level >= 29 ? 20 : level >= 19 ? 30 : level >= 16 ? 50 : level >= 13 ? 70 : level >= 10 ? 80 :
           level == 9 ? 100 : level == 8 ? 130 : level == 7 ? 220 : level == 6 ? 300 : level == 5 ? 380 :
           level == 4 ? 470 : level == 3 ? 550 : level == 2 ? 630 : level == 1 ? 720 : 800;
}

// d = descriptor of the placement's slot, y = height of its bbox bottom, speed = speed_of_level(level of the state it is generated from)
// (the level is the same for every move of a piece: callers that loop over the slots look the bracket up once)
TB_HD bool is_time_at_speed(const SlotDesc& d, int y, const Pressure& pr, int speed)
{
    const int rot = (int)((d.w[0] >> 4) & 3u), ph = (int)((d.w[0] >> 9) & 7u);
    const int clicks = (int)d.w[11] - ((pr.two_but_rot && rot == 3) ? 2 : 0);            // rot 3 costs 1 press instead of 3
    const double t = dadd(pr.resp_time, dmul(dmul(i2d(clicks), pr.move_eff), pr.avg_lat));   // :107, left to right
    const int limit = (y + ph) * speed;                                                  // (rows - row) * speed, :116
    return t <= i2d(limit);
}
TB_HD bool is_time(const SlotDesc& d, int y, const Pressure& pr, int level) { return is_time_at_speed(d, y, pr, speed_of_level(level)); }

// ---------------------------------------------------------------------------------------------
// SURVEY.md §8f rank 4 — move_with_wiggle / move_wiggle (simulator.py:49-85): depth-first flood fill of the
// positions a piece can reach by moving down, left, right (in that order); positions where it cannot move
// down are the moves, and the ORDER of the list (= the tie-break order) is the order of the recursion.
// ---------------------------------------------------------------------------------------------
// Free map of one slot (rotation, column): bit y = the piece fits with its bbox bottom at height y.
TB_HD uint32_t free_mask(const uint32_t c[COLS], const SlotDesc& d)
{
    const int ph = (int)((d.w[0] >> 9) & 7u);
    uint32_t blocked = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) {
        const uint32_t q = (d.w[4 + (k >> 2)] >> (8 * (k & 3))) & 255u;       // piece cells in board column k (bit b = bbox row b)
        blocked |= ((q & 1u) ? c[k] : 0u) | ((q & 2u) ? c[k] >> 1 : 0u) | ((q & 4u) ? c[k] >> 2 : 0u) | ((q & 8u) ? c[k] >> 3 : 0u);
    }
    return ~blocked & ((2u << (ROWS - ph)) - 1u);                             // y in 0 .. ROWS - ph
}

// The reference's recursion for ONE rotation, with an explicit stack.  fm[col] = free map of (rot, col) for col in
// 0..COLS-w and 0 beyond; starts[i] = y | col << 8 in the order the reference visits them; out[i] = y | col << 8 of the
// i-th move.  stack needs 256 entries, seen COLS+1 words (scratch).  Returns the number of moves (<= cap).  With a stop
// set (stop[col] bit y) the search ends at the first visited move that belongs to it: out[0] = that move, returns 1.
TB_HD int wiggle_rot(const uint32_t* fm, int w, const uint16_t* starts, int nstarts, uint16_t* stack, uint32_t* seen, uint16_t* out, int cap,
                      const uint32_t* stop = nullptr)
{
    // seen[col] holds the cells that are free AND not yet visited (a blocked cell returns at once whether or not it was
    // seen before, so only free cells need the visited bit)
    for (int i = 0; i <= COLS; ++i) seen[i] = fm[i];
    int n = 0;
    for (int si = 0; si < nstarts; ++si) {
        int sp = 0;
        stack[sp++] = (uint16_t)(starts[si] & 0x0FFFu);                       // frame = y (5 bits) | col << 8 (4 bits) | stage << 12
        while (sp > 0) {
            const uint32_t fr = stack[sp - 1];
            const int y = (int)(fr & 31u), col = (int)((fr >> 8) & 15u), stage = (int)(fr >> 12);
            if (stage == 0) {
                if (!((seen[col] >> y) & 1u)) { --sp; continue; }                 // :52 seen before, or :54 overlaps here
                seen[col] &= ~(1u << y);
                stack[sp - 1] = (uint16_t)(fr | 0x1000u);
                if (y > 0 && ((fm[col] >> (y - 1)) & 1u)) stack[sp++] = (uint16_t)((uint32_t)(y - 1) | (uint32_t)col << 8);   // :58-60 keep falling
                else {                                                                                                         // :55-57 resting: a move
                    if (n < cap) out[n++] = (uint16_t)((uint32_t)y | (uint32_t)col << 8);
                    if (stop != nullptr && ((stop[col] >> y) & 1u)) {                     // tie break: the first visited move of the stop set
                        out[0] = (uint16_t)((uint32_t)y | (uint32_t)col << 8);
                        return 1;
                    }
                }
            } else if (stage == 1) {
                stack[sp - 1] = (uint16_t)((fr & 0x0FFFu) | 0x2000u);
                if (col > 0) stack[sp++] = (uint16_t)((uint32_t)y | (uint32_t)(col - 1) << 8);          // :63-64
            } else {
                // :65-66 the right neighbour is the last thing this frame does: replace the frame instead of pushing
                if (col + w <= COLS) stack[sp - 1] = (uint16_t)((uint32_t)y | (uint32_t)(col + 1) << 8);
                else --sp;
            }
        }
    }
    return n;
}

// Set-only companion of wiggle_rot: the cells reachable from the starts by down / left / right moves through free
// cells, by bit-parallel flood fill (no visiting order).  The resting cells among them are the moves of the wiggle move
// sets as a SET; the kernels use it when the best move is unique and fall back to wiggle_rot only to break ties.
TB_HD uint32_t fill_down(uint32_t seed, uint32_t free)      // cells reachable from seed (subset of free) by moving down through free cells
{
    uint32_t r = seed, f = free;
    r |= (r >> 1) & f; f &= f >> 1;
    r |= (r >> 2) & f; f &= f >> 2;
    r |= (r >> 4) & f; f &= f >> 4;
    r |= (r >> 8) & f; f &= f >> 8;
    r |= (r >> 16) & f;
    return r;
}

TB_HD void wiggle_reach(const uint32_t* fm, int w, const uint16_t* starts, int nstarts, uint32_t* reach)
{
    const int ncol = COLS - w + 1;
    for (int i = 0; i <= COLS; ++i) reach[i] = 0u;
    for (int i = 0; i < nstarts; ++i) {
        const int y = (int)(starts[i] & 31u), col = (int)(starts[i] >> 8);
        reach[col] |= fm[col] & (1u << y);
    }
    bool changed = true;
    while (changed) {
        changed = false;
        for (int col = 0; col < ncol; ++col) {
            const uint32_t seed = (reach[col] | (col > 0 ? reach[col - 1] : 0u) | (col + 1 < ncol ? reach[col + 1] : 0u)) & fm[col];
            const uint32_t r = fill_down(seed, fm[col]);
            if (r != reach[col]) { reach[col] = r; changed = true; }
        }
        for (int col = ncol - 1; col >= 0; --col) {
            const uint32_t seed = (reach[col] | (col > 0 ? reach[col - 1] : 0u) | (col + 1 < ncol ? reach[col + 1] : 0u)) & fm[col];
            const uint32_t r = fill_down(seed, fm[col]);
            if (r != reach[col]) { reach[col] = r; changed = true; }
        }
    }
    for (int col = 0; col <= COLS; ++col) reach[col] &= ~(fm[col] << 1);     // keep the resting cells: free, and the cell below is not
}

TB_HD bool place_slot(const uint32_t c[COLS], const Heights& H, const SlotDesc& d, Placement& p)
{
    uint32_t f;
    const bool legal = place_pre(c, H, d, p, &f);
    if (legal && f != 0u) place_clear(d, p, f);
    return legal;
}

// statistics of the board after placement P (made on the board whose statistics are pv)
TB_HD PrevStat next_stat(const PrevStat& pv, const Placement& P, const SlotDesc& d)
{
    if (P.k != 0) return prev_stat(P.n);
    PrevStat s;
    const int yb = P.y - 32;
    int t;
    t = (int)byte_of(d.w[8], 0) + yb;  s.gh[0] = t > pv.gh[0] ? t : pv.gh[0];
    t = (int)byte_of(d.w[8], 1) + yb;  s.gh[1] = t > pv.gh[1] ? t : pv.gh[1];
    t = (int)byte_of(d.w[8], 2) + yb;  s.gh[2] = t > pv.gh[2] ? t : pv.gh[2];
    t = (int)(d.w[8] >> 24) + yb;      s.gh[3] = t > pv.gh[3] ? t : pv.gh[3];
    t = (int)byte_of(d.w[9], 0) + yb;  s.gh[4] = t > pv.gh[4] ? t : pv.gh[4];
    t = (int)byte_of(d.w[9], 1) + yb;  s.gh[5] = t > pv.gh[5] ? t : pv.gh[5];
    t = (int)byte_of(d.w[9], 2) + yb;  s.gh[6] = t > pv.gh[6] ? t : pv.gh[6];
    t = (int)(d.w[9] >> 24) + yb;      s.gh[7] = t > pv.gh[7] ? t : pv.gh[7];
    t = (int)byte_of(d.w[10], 0) + yb; s.gh[8] = t > pv.gh[8] ? t : pv.gh[8];
    t = (int)byte_of(d.w[10], 1) + yb; s.gh[9] = t > pv.gh[9] ? t : pv.gh[9];
    s.all_ht = 0; s.max_ht = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) { s.all_ht += s.gh[k]; s.max_ht = s.gh[k] > s.max_ht ? s.gh[k] : s.max_ht; }
    s.cells = pv.cells + 4;
    s.pits = s.all_ht - s.cells;
    s.possum = pv.possum + 4 * P.y + (int)((d.w[0] >> 24) & 15u);
    return s;
}

// ---------------------------------------------------------------------------------------------
// features (features.py:18-347).  Ints stay ints, the five inexact ones are computed in fp64
// with exactly the reference's operation order.
// ---------------------------------------------------------------------------------------------
struct Feat {
    int    iv[NFEAT];      // integer-valued features (also holds exact "x.0" ones)
    double mean_ht, max_ht_diff, min_ht_diff, mean_pit_depth, d_mean_ht;
};

constexpr uint64_t M_HEIGHTS = bit(F_MEAN_HT) | bit(F_MAX_HT) | bit(F_MIN_HT) | bit(F_ALL_HT) | bit(F_MAX_HT_DIFF) |
    bit(F_MIN_HT_DIFF) | bit(F_WELLS) | bit(F_CUML_WELLS) | bit(F_DEEP_WELLS) | bit(F_MAX_WELL) | bit(F_PITS) |
    bit(F_PIT_ROWS) | bit(F_LUMPED_PITS) | bit(F_MEAN_PIT_DEPTH) | (0x1FFull << F_CD_1) | bit(F_ALL_DIFFS) |
    bit(F_MAX_DIFFS) | bit(F_JAGGEDNESS) | bit(F_D_MAX_HT) | bit(F_D_ALL_HT) | bit(F_D_MEAN_HT) | bit(F_D_PITS) |
    bit(F_TETRIS_PROGRESS);
constexpr uint64_t M_WELLS   = bit(F_WELLS) | bit(F_CUML_WELLS) | bit(F_DEEP_WELLS) | bit(F_MAX_WELL);
constexpr uint64_t M_CELLS   = bit(F_FULL_CELLS) | bit(F_PITS) | bit(F_D_PITS) | bit(F_PIT_DEPTH) | bit(F_MEAN_PIT_DEPTH) |
    bit(F_WEIGHTED_CELLS);
constexpr uint64_t M_WEIGHTED = bit(F_WEIGHTED_CELLS) | bit(F_PIT_DEPTH) | bit(F_MEAN_PIT_DEPTH);
constexpr uint64_t M_PITMASK = bit(F_PIT_ROWS) | bit(F_LUMPED_PITS);
constexpr uint64_t M_DIFFS   = (0x1FFull << F_CD_1) | bit(F_ALL_DIFFS) | bit(F_MAX_DIFFS) | bit(F_JAGGEDNESS);
constexpr uint64_t M_NINE    = bit(F_NINE_FILLED) | bit(F_TETRIS_PROGRESS);
constexpr uint64_t M_MEAN    = bit(F_MEAN_HT) | bit(F_MAX_HT_DIFF) | bit(F_MIN_HT_DIFF) | bit(F_D_MEAN_HT);

// CMASK != 0: compile-time feature set (dead code removed); CMASK == 0: `rmask` decides at run time.
// `pv` are the cached statistics of the board the piece was placed on and `d` the slot's descriptor:
// when the move clears no row, heights / cells / height-index sums follow from them by arithmetic.
template <uint64_t CMASK>
TB_HD void eval_features(const Placement& p, const SlotDesc& d, const PrevStat& pv, uint64_t rmask, Feat& f)
{
    const uint64_t need = CMASK ? CMASK : rmask;
    const uint32_t* n = p.n;
    int h[COLS];
    int all_ht = 0, max_ht = 0, min_ht = 0;
    if (need & M_HEIGHTS) {
        if (p.k == 0) {                 // new top of a covered column = landing row + top of the piece column
            const int yb = p.y - 32;
            int t;
            t = (int)byte_of(d.w[8], 0) + yb;  h[0] = t > pv.gh[0] ? t : pv.gh[0];
            t = (int)byte_of(d.w[8], 1) + yb;  h[1] = t > pv.gh[1] ? t : pv.gh[1];
            t = (int)byte_of(d.w[8], 2) + yb;  h[2] = t > pv.gh[2] ? t : pv.gh[2];
            t = (int)(d.w[8] >> 24) + yb;      h[3] = t > pv.gh[3] ? t : pv.gh[3];
            t = (int)byte_of(d.w[9], 0) + yb;  h[4] = t > pv.gh[4] ? t : pv.gh[4];
            t = (int)byte_of(d.w[9], 1) + yb;  h[5] = t > pv.gh[5] ? t : pv.gh[5];
            t = (int)byte_of(d.w[9], 2) + yb;  h[6] = t > pv.gh[6] ? t : pv.gh[6];
            t = (int)(d.w[9] >> 24) + yb;      h[7] = t > pv.gh[7] ? t : pv.gh[7];
            t = (int)byte_of(d.w[10], 0) + yb; h[8] = t > pv.gh[8] ? t : pv.gh[8];
            t = (int)byte_of(d.w[10], 1) + yb; h[9] = t > pv.gh[9] ? t : pv.gh[9];
        } else {
#pragma unroll
            for (int k = 0; k < COLS; ++k) h[k] = fls(n[k]);
        }
        all_ht = h[0]; max_ht = h[0]; min_ht = h[0];
#pragma unroll
        for (int k = 1; k < COLS; ++k) { all_ht += h[k]; max_ht = h[k] > max_ht ? h[k] : max_ht; min_ht = h[k] < min_ht ? h[k] : min_ht; }
        f.iv[F_MAX_HT] = max_ht; f.iv[F_MIN_HT] = min_ht; f.iv[F_ALL_HT] = all_ht;
    }
    if (need & M_MEAN) {
        f.mean_ht = ddiv(i2d(all_ht), 10.0);                          // features.py:24-26
        f.max_ht_diff = dsub(i2d(max_ht), f.mean_ht);                 // :48-50
        f.min_ht_diff = dsub(f.mean_ht, i2d(min_ht));                 // :54-56
        f.d_mean_ht = dsub(f.mean_ht, ddiv(i2d(pv.all_ht), 10.0));    // :251-253
    }
    f.iv[F_CLEARED] = p.k;                                            // :62-77
    f.iv[F_CUML_CLEARED] = (p.k * (p.k + 1)) >> 1;
    f.iv[F_TETRIS] = p.k >> 2;
    if (need & (bit(F_COL_TRANS) | bit(F_ALL_TRANS))) {               // :83-85 (no floor / ceiling)
        int s = 0;
#pragma unroll
        for (int k = 0; k < COLS; ++k) s += popc((n[k] ^ (n[k] >> 1)) & 0x7FFFFu);
        f.iv[F_COL_TRANS] = s;
    }
    if (need & (bit(F_ROW_TRANS) | bit(F_ALL_TRANS))) {               // :89-91 (no walls)
        int s = 0;
#pragma unroll
        for (int k = 0; k + 1 < COLS; ++k) s += popc(n[k] ^ n[k + 1]);
        f.iv[F_ROW_TRANS] = s;
    }
    if (need & bit(F_ALL_TRANS)) f.iv[F_ALL_TRANS] = f.iv[F_COL_TRANS] + f.iv[F_ROW_TRANS];
    if (need & M_WELLS) {                                             // :106-137, walls count as height ROWS
        int ws = 0, cw = 0, dw = 0, mw = 0;
#pragma unroll
        for (int k = 0; k < COLS; ++k) {
            const int l = k == 0 ? ROWS : h[k - 1], r = k == COLS - 1 ? ROWS : h[k + 1];
            int w = (l < r ? l : r) - h[k]; w = w > 0 ? w : 0;
            ws += w; cw += w * (w + 1); dw += w > 1 ? w : 0; mw = w > mw ? w : mw;
        }
        f.iv[F_WELLS] = ws; f.iv[F_CUML_WELLS] = cw >> 1; f.iv[F_DEEP_WELLS] = dw; f.iv[F_MAX_WELL] = mw;
    }
    // filled cells: every placement adds 4 cells and every cleared row removes 10 (exact, no popc)
    const int cells = pv.cells + 4 - 10 * p.k;
    if (need & M_CELLS) {
        f.iv[F_FULL_CELLS] = cells;                                   // :323-325
        f.iv[F_PITS] = all_ht - cells;                                // :146-160  (empty cells under a column top)
        f.iv[F_D_PITS] = f.iv[F_PITS] - pv.pits;                      // :257-259
    }
    if (need & M_WEIGHTED) {                                          // :329-331  sum over cells of (height index + 1)
        // sum of height indices: without a clear it grows by the piece's own cells (4 y + a per-orientation constant)
        const int possum = p.k == 0 ? pv.possum + 4 * p.y + (int)((d.w[0] >> 24) & 15u) : possum_of(n);
        f.iv[F_WEIGHTED_CELLS] = possum + cells;
        // pit_depth (:182-188) = sum over filled cells of the empties below them
        //                      = sum of height indices of filled cells - sum_k m_k(m_k-1)/2, m_k = cells of column k
        int tri = 0;
        if (need & (bit(F_PIT_DEPTH) | bit(F_MEAN_PIT_DEPTH))) {
#pragma unroll
            for (int k = 0; k < COLS; ++k) { const int m = popc(n[k]); tri += m * (m - 1); }
            tri >>= 1;
        }
        f.iv[F_PIT_DEPTH] = possum - tri;
    }
    if (need & bit(F_MEAN_PIT_DEPTH))                                 // :192-194
        f.mean_pit_depth = f.iv[F_PITS] != 0 ? ddiv(i2d(f.iv[F_PIT_DEPTH]), i2d(f.iv[F_PITS])) : 0.0;
    if (need & M_PITMASK) {                                           // :164-178
        uint32_t any = 0; int lumped = 0;
#pragma unroll
        for (int k = 0; k < COLS; ++k) {
            const uint32_t pm = ((1u << h[k]) - 1u) & ~n[k];
            any |= pm; lumped += popc(pm & ~(pm >> 1));
        }
        f.iv[F_PIT_ROWS] = popc(any); f.iv[F_LUMPED_PITS] = lumped;
    }
    if (need & M_DIFFS) {                                             // :201-233
        int mx = h[1] - h[0], jag = 0;
#pragma unroll
        for (int k = 1; k < COLS; ++k) {
            const int d = h[k] - h[k - 1];
            f.iv[F_CD_1 + k - 1] = d; mx = d > mx ? d : mx; jag += d < 0 ? -d : d;
        }
        f.iv[F_ALL_DIFFS] = h[COLS - 1] - h[0]; f.iv[F_MAX_DIFFS] = mx; f.iv[F_JAGGEDNESS] = jag;
    }
    f.iv[F_D_MAX_HT] = max_ht - pv.max_ht;                            // :239-247
    f.iv[F_D_ALL_HT] = all_ht - pv.all_ht;
    f.iv[F_LANDING_HEIGHT] = p.y;                                     // :265-267 (pre-clear)
    if (need & bit(F_PATTERN_DIV)) {                                  // :271-273 distinct column-to-column difference vectors
        uint32_t a[COLS - 1], b[COLS - 1];
#pragma unroll
        for (int j = 0; j + 1 < COLS; ++j) { a[j] = n[j + 1] & ~n[j]; b[j] = n[j] & ~n[j + 1]; }
        int distinct = 0;
#pragma unroll
        for (int j = 0; j + 1 < COLS; ++j) {
            uint32_t dup = 0;
#pragma unroll
            for (int i = 0; i < j; ++i) dup |= (uint32_t)(((a[i] ^ a[j]) | (b[i] ^ b[j])) == 0u);
            distinct += dup ? 0 : 1;
        }
        f.iv[F_PATTERN_DIV] = distinct;
    }
    if (need & M_NINE) {                                              // :279-319
        uint32_t one = 0, two = 0;     // rows with >=1 / >=2 empty cells
#pragma unroll
        for (int k = 0; k < COLS; ++k) { const uint32_t e = ~n[k]; two |= one & e; one |= e; }
        const uint32_t nine = one & ~two & COLMASK;
        f.iv[F_NINE_FILLED] = popc(nine);
        // tetris_progress: the scan from the top can only end just below the lowest column top, so the
        // result is the run of nine-filled rows starting at height index min_ht
        const uint32_t run = ~(nine >> min_ht);
        f.iv[F_TETRIS_PROGRESS] = ctz(run);
    }
    f.iv[F_ERODED_CELLS] = p.eroded;                                  // :335-347
    f.iv[F_CUML_ERODED] = (p.eroded * (p.eroded + 1)) >> 1;
}

// value of feature `id` as the double the canonical scorer multiplies (float(f))
TB_HD double feat_value(const Feat& f, int id)
{
    switch (id) {
    case F_MEAN_HT: return f.mean_ht;
    case F_MAX_HT_DIFF: return f.max_ht_diff;
    case F_MIN_HT_DIFF: return f.min_ht_diff;
    case F_MEAN_PIT_DEPTH: return f.mean_pit_depth;
    case F_D_MEAN_HT: return f.d_mean_ht;
    case F_MAX_HT: return i2d(f.iv[F_MAX_HT]);
    case F_MIN_HT: return i2d(f.iv[F_MIN_HT]);
    case F_ALL_HT: return i2d(f.iv[F_ALL_HT]);
    case F_CLEARED: return i2d(f.iv[F_CLEARED]);
    case F_CUML_CLEARED: return i2d(f.iv[F_CUML_CLEARED]);
    case F_TETRIS: return i2d(f.iv[F_TETRIS]);
    case F_COL_TRANS: return i2d(f.iv[F_COL_TRANS]);
    case F_ROW_TRANS: return i2d(f.iv[F_ROW_TRANS]);
    case F_ALL_TRANS: return i2d(f.iv[F_ALL_TRANS]);
    case F_WELLS: return i2d(f.iv[F_WELLS]);
    case F_CUML_WELLS: return i2d(f.iv[F_CUML_WELLS]);
    case F_DEEP_WELLS: return i2d(f.iv[F_DEEP_WELLS]);
    case F_MAX_WELL: return i2d(f.iv[F_MAX_WELL]);
    case F_PITS: return i2d(f.iv[F_PITS]);
    case F_PIT_ROWS: return i2d(f.iv[F_PIT_ROWS]);
    case F_LUMPED_PITS: return i2d(f.iv[F_LUMPED_PITS]);
    case F_PIT_DEPTH: return i2d(f.iv[F_PIT_DEPTH]);
    case F_CD_1: return i2d(f.iv[F_CD_1]);
    case F_CD_2: return i2d(f.iv[F_CD_2]);
    case F_CD_3: return i2d(f.iv[F_CD_3]);
    case F_CD_4: return i2d(f.iv[F_CD_4]);
    case F_CD_5: return i2d(f.iv[F_CD_5]);
    case F_CD_6: return i2d(f.iv[F_CD_6]);
    case F_CD_7: return i2d(f.iv[F_CD_7]);
    case F_CD_8: return i2d(f.iv[F_CD_8]);
    case F_CD_9: return i2d(f.iv[F_CD_9]);
    case F_ALL_DIFFS: return i2d(f.iv[F_ALL_DIFFS]);
    case F_MAX_DIFFS: return i2d(f.iv[F_MAX_DIFFS]);
    case F_JAGGEDNESS: return i2d(f.iv[F_JAGGEDNESS]);
    case F_D_MAX_HT: return i2d(f.iv[F_D_MAX_HT]);
    case F_D_ALL_HT: return i2d(f.iv[F_D_ALL_HT]);
    case F_D_PITS: return i2d(f.iv[F_D_PITS]);
    case F_LANDING_HEIGHT: return i2d(f.iv[F_LANDING_HEIGHT]);
    case F_PATTERN_DIV: return i2d(f.iv[F_PATTERN_DIV]);
    case F_NINE_FILLED: return i2d(f.iv[F_NINE_FILLED]);
    case F_TETRIS_PROGRESS: return i2d(f.iv[F_TETRIS_PROGRESS]);
    case F_FULL_CELLS: return i2d(f.iv[F_FULL_CELLS]);
    case F_WEIGHTED_CELLS: return i2d(f.iv[F_WEIGHTED_CELLS]);
    case F_ERODED_CELLS: return i2d(f.iv[F_ERODED_CELLS]);
    case F_CUML_ERODED: return i2d(f.iv[F_CUML_ERODED]);
    default: return 0.0;
    }
}

// Scoring modes: the feature list and its order are compile-time for the two headline sets.
// MODE_JIT: a feature list fixed when the translation unit is compiled — TB_JIT_F features with ids TB_JIT_IDS (in summation
// order) and mask TB_JIT_MASK, handed in as macros by the run-time compilation of csrc/jit.cuh (tb200_set_features): the
// same code as MODE_GENERIC with the list known to the compiler, so unused features vanish and the scorer is unrolled.
enum : int { MODE_GENERIC = 0, MODE_W6 = 1, MODE_ALL45 = 2, MODE_JIT = 3 };
template <int MODE> struct ModeTraits;
template <> struct ModeTraits<MODE_GENERIC> { static constexpr uint64_t CMASK = 0; };
template <> struct ModeTraits<MODE_W6>      { static constexpr uint64_t CMASK = W6_MASK; };
template <> struct ModeTraits<MODE_ALL45>   { static constexpr uint64_t CMASK = ALL45_MASK; };
#if defined(TB_JIT_MASK)
template <> struct ModeTraits<MODE_JIT>     { static constexpr uint64_t CMASK = TB_JIT_MASK; };
#endif

// canonical scorer: acc = 0.0; acc += float(f_i) * w_i in the caller's order,
// one rounded multiply and one rounded add per feature.
template <int MODE>
TB_HD double score_features(const Feat& f, const uint8_t* ids, const double* w, int F)
{
    double acc = 0.0;
    if (MODE == MODE_W6) {
        acc = dadd(acc, dmul(feat_value(f, F_LANDING_HEIGHT), w[0]));
        acc = dadd(acc, dmul(feat_value(f, F_ERODED_CELLS), w[1]));
        acc = dadd(acc, dmul(feat_value(f, F_ROW_TRANS), w[2]));
        acc = dadd(acc, dmul(feat_value(f, F_COL_TRANS), w[3]));
        acc = dadd(acc, dmul(feat_value(f, F_PITS), w[4]));
        acc = dadd(acc, dmul(feat_value(f, F_CUML_WELLS), w[5]));
    } else if (MODE == MODE_ALL45) {
#pragma unroll
        for (int i = 0; i < NFEAT; ++i) acc = dadd(acc, dmul(feat_value(f, i), w[i]));
#if defined(TB_JIT_MASK)
    } else if (MODE == MODE_JIT) {
        constexpr int jit_ids[TB_JIT_F] = TB_JIT_IDS;
#pragma unroll
        for (int i = 0; i < TB_JIT_F; ++i) acc = dadd(acc, dmul(feat_value(f, jit_ids[i]), w[i]));
#endif
    } else {
        for (int i = 0; i < F; ++i) acc = dadd(acc, dmul(feat_value(f, ids[i]), w[i]));
    }
    return acc;
}

// order-preserving map double -> u64 (a < b  <=>  key(a) < key(b)) for every score the canonical scorer can produce.
// The one pair of doubles that compares equal with different bits is (-0.0, +0.0) (simulator.py:130,134 use `<`), and the
// scorer cannot return -0.0: it starts from acc = +0.0 and in round-to-nearest a sum is -0.0 only when BOTH operands are
// -0.0 ((+0.0) + (-0.0) = +0.0, x + (-x) = +0.0), so by induction acc is never -0.0.  Scores are therefore reported with
// their exact sign and need no canonicalising add here.  NaN cannot occur either: weights are finite and bounded
// (include/tetris_b200.h, Limits) and features are small integers.
TB_HD uint64_t score_key(double s)
{
#if defined(__CUDA_ARCH__)
    const uint64_t b = (uint64_t)__double_as_longlong(s);
#else
    uint64_t b; __builtin_memcpy(&b, &s, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------
// Fast evaluation of the "w6" feature set (landing_height, eroded_cells, row_trans, col_trans,
// pits, cuml_wells — SURVEY.md §8c).  Same results as eval_features<W6_MASK> + score_features, but
// organised for a short dependency chain: when the placement clears no row (the common case)
// the new column heights come from arithmetic on the landing row (no clz), the filled-cell count
// is the game's count + 4 (no popc), and only the two transition counts need popc.
// ---------------------------------------------------------------------------------------------
struct GameStat {
    int gh[COLS];          // column heights of the game's current board
    Heights H;             // the same, packed (byte k = height of column k)
    int cells;             // filled cells on the board
};

TB_HD void unpack_heights(const Heights& H, int gh[COLS])
{
    gh[0] = (int)byte_of(H.H0, 0); gh[1] = (int)byte_of(H.H0, 1); gh[2] = (int)byte_of(H.H0, 2); gh[3] = (int)(H.H0 >> 24);
    gh[4] = (int)byte_of(H.H1, 0); gh[5] = (int)byte_of(H.H1, 1); gh[6] = (int)byte_of(H.H1, 2); gh[7] = (int)(H.H1 >> 24);
    gh[8] = (int)byte_of(H.H2, 0); gh[9] = (int)byte_of(H.H2, 1);
}

TB_HD GameStat game_stat(const uint32_t c[COLS])
{
    GameStat g;
    g.H = pack_heights(c);
    unpack_heights(g.H, g.gh);
    g.cells = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) g.cells += popc(c[k]);
    return g;
}

struct W6Cand {            // one candidate placement in flight
    Placement P;
    int h[COLS];           // post-move column heights
    int cells;             // post-move filled cells
    uint32_t full;         // full-row mask before clearing
    bool legal;
};

// straight-line part: landing, legality, imprint, full-row mask, no-clear heights and cell count
TB_HD void w6_pre(const uint32_t c[COLS], const GameStat& g, const SlotDesc& d, W6Cand& a)
{
    a.legal = place_pre(c, g.H, d, a.P, &a.full);
    const int yb = a.P.y - 32;
    int t;
    t = (int)byte_of(d.w[8], 0) + yb;  a.h[0] = t > g.gh[0] ? t : g.gh[0];
    t = (int)byte_of(d.w[8], 1) + yb;  a.h[1] = t > g.gh[1] ? t : g.gh[1];
    t = (int)byte_of(d.w[8], 2) + yb;  a.h[2] = t > g.gh[2] ? t : g.gh[2];
    t = (int)(d.w[8] >> 24) + yb;      a.h[3] = t > g.gh[3] ? t : g.gh[3];
    t = (int)byte_of(d.w[9], 0) + yb;  a.h[4] = t > g.gh[4] ? t : g.gh[4];
    t = (int)byte_of(d.w[9], 1) + yb;  a.h[5] = t > g.gh[5] ? t : g.gh[5];
    t = (int)byte_of(d.w[9], 2) + yb;  a.h[6] = t > g.gh[6] ? t : g.gh[6];
    t = (int)(d.w[9] >> 24) + yb;      a.h[7] = t > g.gh[7] ? t : g.gh[7];
    t = (int)byte_of(d.w[10], 0) + yb; a.h[8] = t > g.gh[8] ? t : g.gh[8];
    t = (int)byte_of(d.w[10], 1) + yb; a.h[9] = t > g.gh[9] ? t : g.gh[9];
    a.cells = g.cells + 4;
}

// rare part (call only when a.legal && a.full != 0): clear rows, then heights by clz, cells - 10 per row
TB_HD void w6_clear(const SlotDesc& d, W6Cand& a)
{
    place_clear(d, a.P, a.full);
#pragma unroll
    for (int k = 0; k < COLS; ++k) a.h[k] = fls(a.P.n[k]);
    a.cells -= 10 * a.P.k;
}

// features + canonical fp64 score; returns the comparison key (0 when the slot is illegal) and the
// packed post-move heights
TB_HD uint64_t w6_post(const W6Cand& a, const double* w, Heights& Hn)
{
    const uint32_t* n = a.P.n;
    const int* h = a.h;
    int all_ht = ((h[0] + h[1]) + (h[2] + h[3])) + ((h[4] + h[5]) + (h[6] + h[7])) + (h[8] + h[9]);
    const int pits = all_ht - a.cells;                                       // features.py:146-160
    int w1 = 0, w2 = 0;                                                      // features.py:106-125: sum w(w+1)/2
#pragma unroll
    for (int k = 0; k < COLS; ++k) {
        const int l = k == 0 ? ROWS : h[k - 1], r = k == COLS - 1 ? ROWS : h[k + 1];
        int wl = (l < r ? l : r) - h[k]; wl = wl > 0 ? wl : 0;
        w1 += wl; w2 += wl * wl;
    }
    const int cw = (w1 + w2) >> 1;
    int rt = 0, ct = 0;
#pragma unroll
    for (int k = 0; k + 1 < COLS; ++k) rt += popc(n[k] ^ n[k + 1]);          // features.py:89-91
    // features.py:83-85: bit i of n ^ 2n (i = 1..19) compares heights i-1 and i; the doubling is an integer
    // multiply-add, which runs on the FMA pipe instead of the (busier) ALU pipe
#pragma unroll
    for (int k = 0; k < COLS; ++k) ct += popc((n[k] ^ (n[k] * 2u)) & 0xFFFFEu);
    Hn.H0 = (uint32_t)h[0] | (uint32_t)h[1] << 8 | (uint32_t)h[2] << 16 | (uint32_t)h[3] << 24;
    Hn.H1 = (uint32_t)h[4] | (uint32_t)h[5] << 8 | (uint32_t)h[6] << 16 | (uint32_t)h[7] << 24;
    Hn.H2 = (uint32_t)h[8] | (uint32_t)h[9] << 8;
    double acc = 0.0;                                                        // canonical order
    acc = dadd(acc, dmul(u2d(a.P.y), w[0]));                                 // landing_height (all six are >= 0)
    acc = dadd(acc, dmul(u2d(a.P.eroded), w[1]));                            // eroded_cells
    acc = dadd(acc, dmul(u2d(rt), w[2]));                                    // row_trans
    acc = dadd(acc, dmul(u2d(ct), w[3]));                                    // col_trans
    acc = dadd(acc, dmul(u2d(pits), w[4]));                                  // pits
    acc = dadd(acc, dmul(u2d(cw), w[5]));                                    // cuml_wells
    const uint64_t key = score_key(acc);
    return a.legal ? key : 0ull;
}

// ---------------------------------------------------------------------------------------------
// Window-local evaluation of the w6 set (lockstep kernel K2s, one lane per game).
// A hard drop changes at most four columns, so of the six features only the contributions of a small window move:
//   row_trans  : the column pairs that touch the piece's columns            (<= 5 of 9 pairs)
//   col_trans  : the piece's columns                                         (<= 4 of 10 columns)
//   pits       : sum of heights - cells, heights change only under the piece
//   cuml_wells : the wells of the piece's columns and of their two neighbours (<= 6 of 10 columns)
// WinCache keeps the per-pair / per-column contributions of the CURRENT board (packed bytes) and their totals; a
// placement that completes no row is then scored from the window alone: (old total) - (old window part) + (new window part).
// Placements that complete a row are NOT handled here (the caller defers them to the general evaluation): every column
// moves when a row disappears.  Same values as w6_pre + w6_post, bit for bit (tests/test_core_emu.py checks every slot of
// thousands of boards on the CPU; the GPU parity tests check whole games).
// ---------------------------------------------------------------------------------------------
struct WinCache {
    uint32_t X[4];       // extended packed heights: byte i = height of column i - 2; bytes 0, 1 and 12, 13 (outside the board) = ROWS: walls
    uint32_t RTp[3];     // byte k + 1 = popc(c_k ^ c_{k+1}) for the pair (k, k+1), k = 0..8; byte 0 = 0
    uint32_t CTc[3];     // byte k = popc((c_k ^ 2 c_k) & 0xFFFFE)
    uint32_t Wd[3];      // byte k + 1 = well depth of column k; byte 0 = 0
    int RT, CT, CW2, SH, cells;      // totals: row_trans, col_trans, sum of d (d + 1) over the wells, sum of heights, filled cells
};

TB_HD int bytesum_sel(uint32_t v, uint32_t sel01)        // sum of the bytes of v selected by sel01 (0x01 in the chosen bytes)
{
#if defined(__CUDA_ARCH__)
    return (int)__dp4a(v, sel01, 0u);
#else
    int s = 0;
    for (int i = 0; i < 4; ++i) s += (int)((v >> (8 * i)) & 255u) * (int)((sel01 >> (8 * i)) & 255u);
    return s;
#endif
}
TB_HD int bytesq_sel(uint32_t v, uint32_t selff)         // sum of b * b + b over the bytes of v selected by selff (0xFF in the chosen bytes)
{
    const uint32_t m = v & selff;
#if defined(__CUDA_ARCH__)
    return (int)__dp4a(m, m, __dp4a(m, 0x01010101u, 0u));
#else
    int s = 0;
    for (int i = 0; i < 4; ++i) { const int b = (int)((m >> (8 * i)) & 255u); s += b * b + b; }
    return s;
#endif
}
// bytes i .. i+3 of the byte array held in (a0, a1, a2, a3); i in 0..12 (bytes beyond the array read as 0)
TB_HD uint32_t win_bytes4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, int i)
{
    const int wi = i >> 2;
    const uint32_t sh = 8u * (uint32_t)(i & 3);
    const uint32_t lo = wi == 0 ? a0 : (wi == 1 ? a1 : (wi == 2 ? a2 : a3));
    const uint32_t hi = wi == 0 ? a1 : (wi == 1 ? a2 : (wi == 2 ? a3 : 0u));
    return funnel_r(lo, hi, sh);
}
TB_CX uint32_t low_bytes01(int n) { return n <= 0 ? 0u : (n >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - n)))); }   // 0x01 in the n low bytes

TB_HD void win_cache_init(const uint32_t c[COLS], WinCache& wc)
{
    int h[COLS];
    wc.SH = 0; wc.cells = 0; wc.RT = 0; wc.CT = 0; wc.CW2 = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) { h[k] = fls(c[k]); wc.SH += h[k]; wc.cells += popc(c[k]); }
    wc.X[0] = (uint32_t)ROWS | (uint32_t)ROWS << 8 | (uint32_t)h[0] << 16 | (uint32_t)h[1] << 24;
    wc.X[1] = (uint32_t)h[2] | (uint32_t)h[3] << 8 | (uint32_t)h[4] << 16 | (uint32_t)h[5] << 24;
    wc.X[2] = (uint32_t)h[6] | (uint32_t)h[7] << 8 | (uint32_t)h[8] << 16 | (uint32_t)h[9] << 24;
    wc.X[3] = (uint32_t)ROWS | (uint32_t)ROWS << 8;
    uint32_t rt[3] = {0u, 0u, 0u}, ct[3] = {0u, 0u, 0u}, wd[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < COLS; ++k) {
        if (k + 1 < COLS) { const int v = popc(c[k] ^ c[k + 1]); wc.RT += v; rt[(k + 1) >> 2] |= (uint32_t)v << (8 * ((k + 1) & 3)); }
        const int t = popc((c[k] ^ (c[k] * 2u)) & 0xFFFFEu); wc.CT += t; ct[k >> 2] |= (uint32_t)t << (8 * (k & 3));
        const int l = k == 0 ? ROWS : h[k - 1], r = k == COLS - 1 ? ROWS : h[k + 1];
        int d = (l < r ? l : r) - h[k]; d = d > 0 ? d : 0;
        wc.CW2 += d * d + d; wd[(k + 1) >> 2] |= (uint32_t)d << (8 * ((k + 1) & 3));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { wc.RTp[i] = rt[i]; wc.CTc[i] = ct[i]; wc.Wd[i] = wd[i]; }
}

// The windows a placement at first column c0 needs, cut out of the cache.  With slots enumerated column-minor (c0 = 0, 1,
// 2, ... within a rotation) they SLIDE: every array moves down by one byte per slot (win_advance), so no run-time byte
// indexing is needed on the device — a reset at c0 == 0 and three or four constant shifts per slot.
struct WinSlide {
    uint32_t x[4];       // byte 0 = height of column c0 - 2 (the 8-byte heights window is x[0], x[1])
    uint32_t r[3];       // byte 0 = pair (c0 - 1, c0)   (0 when c0 == 0; pairs beyond the board read 0)
    uint32_t t[3];       // byte 0 = col_trans of column c0
    uint32_t d[3];       // byte 0 = well depth of column c0 - 1 (0 outside the board)
};
TB_HD void win_reset(const WinCache& wc, WinSlide& ws)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) ws.x[i] = wc.X[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ws.r[i] = wc.RTp[i]; ws.t[i] = wc.CTc[i]; ws.d[i] = wc.Wd[i]; }
}
TB_HD void win_advance(WinSlide& ws)
{
    ws.x[0] = funnel_r(ws.x[0], ws.x[1], 8u); ws.x[1] = funnel_r(ws.x[1], ws.x[2], 8u); ws.x[2] = funnel_r(ws.x[2], ws.x[3], 8u); ws.x[3] >>= 8;
    ws.r[0] = funnel_r(ws.r[0], ws.r[1], 8u); ws.r[1] = funnel_r(ws.r[1], ws.r[2], 8u); ws.r[2] >>= 8;
    ws.t[0] = funnel_r(ws.t[0], ws.t[1], 8u); ws.t[1] = funnel_r(ws.t[1], ws.t[2], 8u); ws.t[2] >>= 8;
    ws.d[0] = funnel_r(ws.d[0], ws.d[1], 8u); ws.d[1] = funnel_r(ws.d[1], ws.d[2], 8u); ws.d[2] >>= 8;
}
// start the windows at first column c0 (0 .. 9, warp-uniform on the device) instead of 0: whole words move, the rest slides
TB_HD void win_skip(WinSlide& ws, int c0)
{
    for (int i = c0 >> 2; i > 0; --i) {
        ws.x[0] = ws.x[1]; ws.x[1] = ws.x[2]; ws.x[2] = ws.x[3]; ws.x[3] = 0u;
        ws.r[0] = ws.r[1]; ws.r[1] = ws.r[2]; ws.r[2] = 0u;
        ws.t[0] = ws.t[1]; ws.t[1] = ws.t[2]; ws.t[2] = 0u;
        ws.d[0] = ws.d[1]; ws.d[1] = ws.d[2]; ws.d[2] = 0u;
    }
    for (int i = c0 & 3; i > 0; --i) win_advance(ws);
}
// per-slot constants that do not depend on the board (uniform on the device): piece cells / tops of the piece's own four
// columns, cut out of the slot descriptor
struct WinDesc {
    uint32_t pbw, tpw;               // piece cells / tops of the piece's columns 0..3
    uint32_t mr0, mr1, mt, md0, md1; // byte selectors of the old contributions: first wdt + 1 pairs, wdt columns, wdt + 2 wells of the windows
    uint32_t pad;
};
TB_HD WinDesc win_desc(const SlotDesc& d)
{
    const int c0 = (int)(d.w[0] & 15u), wdt = (int)((d.w[0] >> 6) & 7u);
    WinDesc wd;
    wd.pbw = win_bytes4(d.w[4], d.w[5], d.w[6], 0u, c0);       // piece cells of the piece's columns 0..3 (bit q = bbox row q)
    wd.tpw = win_bytes4(d.w[8], d.w[9], d.w[10], 0u, c0);      // 32 + top of the piece's columns (0 beyond its width)
    wd.mr0 = low_bytes01(wdt + 1); wd.mr1 = low_bytes01(wdt - 3);
    wd.mt = low_bytes01(wdt);
    wd.md0 = low_bytes01(wdt + 2) * 255u; wd.md1 = low_bytes01(wdt - 2) * 255u;
    wd.pad = 0u;
    return wd;
}

// One hard drop scored from its window.  ws: the windows at the slot's first column; col(k): column word k of the current
// board (k = 0..9; only the piece's columns and their two neighbours are asked for); and_out: AND of every column OUTSIDE
// the piece's columns (for the full-row test).  Returns the comparison key (0 = illegal); *full = rows the placement
// completes: when it is not 0 the returned key is meaningless and the caller must evaluate the placement with
// w6_pre / w6_clear / w6_post.
// WDT = the piece's width in this rotation (1..4), a compile-time constant: the lockstep kernel switches on it once per
// rotation (it is warp-uniform), so that the slot loop carries no width tests, only the WDT piece columns, WDT height updates,
// WDT + 2 wells and the byte selectors of exactly that width as immediates (about a tenth fewer instructions per slot than the
// run-time-width body, and straight-line code between the two neighbour tests).
template <int WDT, class ColFn>
TB_HD uint64_t win_eval_w(const WinCache& wc, const WinSlide& ws, const SlotDesc& d, const WinDesc& wdsc, const double* w, ColFn col,
                          uint32_t and_out, uint32_t* full, bool* legal_out, int* y_out = nullptr)
{
    static_assert(WDT >= 1 && WDT <= 4, "piece width");
    const int c0 = (int)(d.w[0] & 15u);
    const uint32_t W0 = ws.x[0], W1 = ws.x[1];                            // heights of columns c0-2 .. c0+5
    const uint32_t hw = funnel_r(W0, W1, 16u);                            // columns c0 .. c0+3
    const bool legal = ((hw + d.w[2]) & 0x80808080u) == 0u;
    const uint32_t v = hw + d.w[1];
    uint32_t m01 = byte_of(v, 0), b1 = byte_of(v, 1), b2 = byte_of(v, 2), b3 = v >> 24;
    m01 = m01 > b1 ? m01 : b1; b2 = b2 > b3 ? b2 : b3; m01 = m01 > b2 ? m01 : b2;
    const int y = (int)m01 - 32;
    *legal_out = legal;
    if (y_out) *y_out = y;
    const uint32_t pbw = wdsc.pbw, tpw = wdsc.tpw;
    const uint32_t one_y = 1u << (y & 31);
    // imprint + transitions of the window
    const uint32_t n0 = byte_of(pbw, 0) * one_y + col(c0);
    uint32_t nl = n0;                                                     // last (right-most) piece column
    uint32_t f = and_out & n0;
    int rtn = 0, ctn = popc((n0 ^ (n0 * 2u)) & 0xFFFFEu);
    if (c0 > 0) rtn += popc(col(c0 - 1) ^ n0);
    if constexpr (WDT > 1) { const uint32_t n1 = byte_of(pbw, 1) * one_y + col(c0 + 1); f &= n1; rtn += popc(nl ^ n1); ctn += popc((n1 ^ (n1 * 2u)) & 0xFFFFEu); nl = n1; }
    if constexpr (WDT > 2) { const uint32_t n2 = byte_of(pbw, 2) * one_y + col(c0 + 2); f &= n2; rtn += popc(nl ^ n2); ctn += popc((n2 ^ (n2 * 2u)) & 0xFFFFEu); nl = n2; }
    if constexpr (WDT > 3) { const uint32_t n3 = (pbw >> 24) * one_y + col(c0 + 3); f &= n3; rtn += popc(nl ^ n3); ctn += popc((n3 ^ (n3 * 2u)) & 0xFFFFEu); nl = n3; }
    if (c0 + WDT < COLS) rtn += popc(nl ^ col(c0 + WDT));
    *full = f;
    // old contributions of the same pairs / columns: the first WDT + 1 pairs, WDT columns, WDT + 2 wells of the windows
    // (entries outside the board are 0 in the arrays, so the edges need no special case)
    int rto = bytesum_sel(ws.r[0], low_bytes01(WDT + 1));
    if constexpr (WDT > 3) rto += bytesum_sel(ws.r[1], low_bytes01(WDT - 3));
    const int cto = bytesum_sel(ws.t[0], low_bytes01(WDT));
    int cwo = bytesq_sel(ws.d[0], low_bytes01(WDT + 2) * 255u);
    if constexpr (WDT > 2) cwo += bytesq_sel(ws.d[1], low_bytes01(WDT - 2) * 255u);
    // heights under the piece
    const int yb = y - 32;
    int q[8];
    q[0] = (int)byte_of(W0, 0); q[1] = (int)byte_of(W0, 1); q[2] = (int)byte_of(W0, 2); q[3] = (int)(W0 >> 24);
    q[4] = (int)byte_of(W1, 0); q[5] = (int)byte_of(W1, 1); q[6] = (int)byte_of(W1, 2); q[7] = (int)(W1 >> 24);
    int dsh = 0;
    { const int t = (int)byte_of(tpw, 0) + yb; const int hn = t > q[2] ? t : q[2]; dsh += hn - q[2]; q[2] = hn; }
    if constexpr (WDT > 1) { const int t = (int)byte_of(tpw, 1) + yb; const int hn = t > q[3] ? t : q[3]; dsh += hn - q[3]; q[3] = hn; }
    if constexpr (WDT > 2) { const int t = (int)byte_of(tpw, 2) + yb; const int hn = t > q[4] ? t : q[4]; dsh += hn - q[4]; q[4] = hn; }
    if constexpr (WDT > 3) { const int t = (int)(tpw >> 24) + yb;     const int hn = t > q[5] ? t : q[5]; dsh += hn - q[5]; q[5] = hn; }
    // wells of columns c0-1 .. c0+WDT (window positions 1 .. WDT+2); positions outside the board are walls of height ROWS in the
    // window, so their "depth" is 0 by itself
    int cwn = 0;
#pragma unroll
    for (int p = 1; p <= WDT + 2; ++p) {
        int dd = (q[p - 1] < q[p + 1] ? q[p - 1] : q[p + 1]) - q[p]; dd = dd > 0 ? dd : 0;
        cwn += dd * dd + dd;
    }
    const int rt = wc.RT - rto + rtn, ct = wc.CT - cto + ctn;
    const int pits = wc.SH + dsh - (wc.cells + 4);
    const int cw = (wc.CW2 - cwo + cwn) >> 1;
    double acc = 0.0;                                                              // canonical order
    acc = dadd(acc, dmul(u2d(y), w[0]));                                           // landing_height
    // eroded_cells = 0 here (no row completed): the reference adds 0.0 * w[1] = +-0.0, which never changes the bits of acc
    // (x + (+-0) = x, and acc is never -0.0: see score_key), so the term is skipped
    acc = dadd(acc, dmul(u2d(rt), w[2]));
    acc = dadd(acc, dmul(u2d(ct), w[3]));
    acc = dadd(acc, dmul(u2d(pits), w[4]));
    acc = dadd(acc, dmul(u2d(cw), w[5]));
    const uint64_t key = score_key(acc);
    return legal ? key : 0ull;
}

// the same with the width read from the descriptor (host tests, and callers outside the per-rotation switch)
template <class ColFn>
TB_HD uint64_t win_eval_at(const WinCache& wc, const WinSlide& ws, const SlotDesc& d, const WinDesc& wdsc, const double* w, ColFn col,
                           uint32_t and_out, uint32_t* full, bool* legal_out, int* y_out = nullptr)
{
    switch ((int)((d.w[0] >> 6) & 7u)) {
    case 1: return win_eval_w<1>(wc, ws, d, wdsc, w, col, and_out, full, legal_out, y_out);
    case 2: return win_eval_w<2>(wc, ws, d, wdsc, w, col, and_out, full, legal_out, y_out);
    case 3: return win_eval_w<3>(wc, ws, d, wdsc, w, col, and_out, full, legal_out, y_out);
    default: return win_eval_w<4>(wc, ws, d, wdsc, w, col, and_out, full, legal_out, y_out);
    }
}

// the same for a stand-alone slot (host tests): windows cut out by byte index
template <class ColFn>
TB_HD uint64_t win_eval(const WinCache& wc, const SlotDesc& d, const double* w, ColFn col, uint32_t and_out, uint32_t* full, bool* legal_out)
{
    const int c0 = (int)(d.w[0] & 15u);
    WinSlide ws;
    ws.x[0] = win_bytes4(wc.X[0], wc.X[1], wc.X[2], wc.X[3], c0); ws.x[1] = win_bytes4(wc.X[0], wc.X[1], wc.X[2], wc.X[3], c0 + 4); ws.x[2] = ws.x[3] = 0u;
    ws.r[0] = win_bytes4(wc.RTp[0], wc.RTp[1], wc.RTp[2], 0u, c0); ws.r[1] = win_bytes4(wc.RTp[0], wc.RTp[1], wc.RTp[2], 0u, c0 + 4); ws.r[2] = 0u;
    ws.t[0] = win_bytes4(wc.CTc[0], wc.CTc[1], wc.CTc[2], 0u, c0); ws.t[1] = ws.t[2] = 0u;
    ws.d[0] = win_bytes4(wc.Wd[0], wc.Wd[1], wc.Wd[2], 0u, c0); ws.d[1] = win_bytes4(wc.Wd[0], wc.Wd[1], wc.Wd[2], 0u, c0 + 4); ws.d[2] = 0u;
    return win_eval_at(wc, ws, d, win_desc(d), w, col, and_out, full, legal_out);
}

}  // namespace tb
