// play_quad.cuh — K2q: FOUR warps per game with sub-placement parallelism, the last stages of the latency regime
// (BASELINE config 2) (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "play_common.cuh"

namespace tb {

// ---------------------------------------------------------------------------------------------
// K2q: one 128-thread block per game, lookahead 1, w6.  When at most one or two games per SM are still alive the run
// time is the serial chain  pieces x per-piece latency  of those games, and a lone warp cannot issue faster than one
// ALU-pipe instruction every two cycles, so what matters is the number of instructions EACH WARP executes per piece.
// K2f / K2p give every placement one lane, which evaluates all ten columns (~260 instructions).  Here a placement is
// evaluated by THREE lanes, each owning a column range of the board:
//     part 0: columns 0-3     part 1: columns 4-6     part 2: columns 7-9
// A lane holds only its own columns (+ the right neighbour's first column for the row-transition pair that straddles
// the boundary) and the six column heights around them; it imprints / clears / counts transitions and wells for its
// range, and the three lanes meet in ONE round of shuffles (packed partial sums) before the canonical fp64 score,
// which all three compute identically.  Ten placements per warp, forty per block: the 34-slot pieces (T, L, J) are one
// pass; warps whose slot range is empty for the current piece skip the evaluation.
// Per piece:  y / legality (every lane, from the packed heights) -> imprint own columns -> full-row mask (AND over the
// three lanes, one shuffle round) -> [rows completed: clear own columns, heights by clz, left neighbour's height by
// shuffle] -> heights / wells / transitions of the own range -> partial sums by shuffle -> score -> REDUX argmax inside
// the warp -> the four warp winners meet in shared memory (one block barrier, double-buffered records) -> every lane
// re-applies the winning placement to ITS columns (commit), the packed heights are rebuilt by three shuffles.
// Same results as every other w6 kernel (first-of-max: slots ascend with the warp index and, inside a warp, with the lane).
// ---------------------------------------------------------------------------------------------
struct QDesc { uint32_t e[8]; };     // per (piece, slot, part): see build_qdesc
constexpr int QPARTS = 3;
constexpr int QGROUPS = 10;          // placements per warp
constexpr int QBLOCK = 128;

struct alignas(16) QShared {
    QDesc qd[7][TB_MAX_SLOTS][QPARTS];
    uint4 rec[2][4];                 // per parity, per warp: {key hi, key lo, slot | y << 8, full-row mask}
    uint32_t fin[2];                 // finished-games counter as seen by thread 0 (hand-over decisions must be block-uniform)
    uint32_t game;
};

// e0 misc (as SlotDesc w0)   e1 landing bias   e2 legality bias   e3 window select | nslots << 16
// e4 piece column masks of board columns b .. b+3 (b = first own column: 0 / 4 / 7)   e5 mask of column b+4
// e6 top bytes (32 + top, 0 = not covered) of the view columns b-1 .. b+2   e7 of b+3, b+4
__device__ __forceinline__ QDesc build_qdesc(const SlotDesc& d, int part)
{
    const int b = part == 0 ? 0 : (part == 1 ? 4 : 7);
    auto cm = [&](int k) -> uint32_t { return (k < 0 || k > 11) ? 0u : (d.w[4 + (k >> 2)] >> (8 * (k & 3))) & 255u; };
    auto tp = [&](int k) -> uint32_t { return (k < 0 || k > 11) ? 0u : (d.w[8 + (k >> 2)] >> (8 * (k & 3))) & 255u; };
    QDesc q;
    q.e[0] = d.w[0]; q.e[1] = d.w[1]; q.e[2] = d.w[2]; q.e[3] = d.w[3] | d.w[7] << 16;
    q.e[4] = cm(b) | cm(b + 1) << 8 | cm(b + 2) << 16 | cm(b + 3) << 24;
    q.e[5] = cm(b + 4);
    q.e[6] = tp(b - 1) | tp(b) << 8 | tp(b + 1) << 16 | tp(b + 2) << 24;
    q.e[7] = tp(b + 3) | tp(b + 4) << 8;
    return q;
}

__global__ void __launch_bounds__(QBLOCK) play_games_w6quad_kernel(const PlayParams prm)
{
    const bool staged = prm.st_ctl != nullptr;
    uint32_t* const ticket = staged ? prm.st_ctl + prm.st_stage : prm.ticket;
    uint32_t n_in = prm.n_games;
    if (staged && prm.st_list_in) {
        n_in = prm.st_ctl[8 + prm.st_stage];
        if (n_in == 0u) return;                      // nothing was handed over: the earlier stages finished every game
    }
    const uint32_t* const finished = staged ? prm.st_ctl + 4 : nullptr;
    extern __shared__ uint4 q_smem_raw[];
    QShared& S = *reinterpret_cast<QShared*>(q_smem_raw);
    for (int i = threadIdx.x; i < 7 * TB_MAX_SLOTS * QPARTS; i += QBLOCK) {
        const int p = i / (TB_MAX_SLOTS * QPARTS), s = (i / QPARTS) % TB_MAX_SLOTS, part = i % QPARTS;
        S.qd[p][s][part] = build_qdesc(g_tables.slot[p][s], part);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // lanes 30, 31 shadow group 9 (they keep a valid copy of the state, never hold a legal placement)
    const int gq = lane < 30 ? lane / 3 : 9;
    const int part = lane < 30 ? lane - 3 * gq : lane - 30;
    const bool shadow = lane >= 30;
    const int g3 = 3 * gq;                             // first lane of the group
    const int slot = warp * QGROUPS + gq;
    const int myslot = slot < TB_MAX_SLOTS ? slot : 0;
    const int src1 = g3 + (part + 1) % 3, src2 = g3 + (part + 2) % 3;
    const uint32_t T = prm.T;
    // per-part constants: which of the lane's four column registers / pairs belong to it
    const uint32_t own3 = part == 0 ? 0xffffffffu : 0u;          // q3 is an own column (part 0 only)
    const uint32_t pm2 = part == 2 ? 0u : 0xffffffffu;           // pair (q2, q3) is mine
    const int wall4 = part == 2 ? ROWS : 0;                      // view column 4 is the right wall (part 2)

    for (;;) {
        if (threadIdx.x == 0) {
            uint32_t gi = atomicAdd(ticket, 1u);
            if (gi < n_in && prm.st_list_in) gi = prm.st_list_in[gi]; else if (gi >= n_in) gi = 0xffffffffu;
            S.game = gi;
        }
        __syncthreads();
        const uint32_t game = S.game;
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
        if (staged && threadIdx.x == 0) prm.st_save[game].cnt = 0u;      // the four warps add their placement counts here at a hand-over (barriers follow before any of them does)
        // ---- distribute the board: own columns q0..q4 = board columns b .. b+4, view heights hv0..hv5 = columns b-1 .. b+4
        uint32_t q0, q1, q2, q3, q4;
        int hv0, hv1, hv2, hv3, hv4, hv5;
        Heights H;
        int cells0;
        {
            const GameStat g = game_stat(c);
            H = g.H;
            cells0 = g.cells - 4 * (int)t;
            if (part == 0) { q0 = c[0]; q1 = c[1]; q2 = c[2]; q3 = c[3]; q4 = c[4]; hv0 = ROWS; hv1 = g.gh[0]; hv2 = g.gh[1]; hv3 = g.gh[2]; hv4 = g.gh[3]; hv5 = g.gh[4]; }
            else if (part == 1) { q0 = c[4]; q1 = c[5]; q2 = c[6]; q3 = c[7]; q4 = c[8]; hv0 = g.gh[3]; hv1 = g.gh[4]; hv2 = g.gh[5]; hv3 = g.gh[6]; hv4 = g.gh[7]; hv5 = g.gh[8]; }
            else { q0 = c[7]; q1 = c[8]; q2 = c[9]; q3 = 0u; q4 = 0u; hv0 = g.gh[6]; hv1 = g.gh[7]; hv2 = g.gh[8]; hv3 = g.gh[9]; hv4 = ROWS; hv5 = ROWS; }
        }
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
        auto piece_at = [&](uint32_t tt) -> int {          // base <= tt < base + 64, uniform tt
            const uint32_t idx = tt - base;
            const uint32_t v = __shfl_sync(FULLMASK, win, (int)(idx & 31u));
            return (int)((v >> ((idx >> 5) * 8u)) & 255u);
        };
        const uint32_t lines0 = lines;
        const uint32_t stop = T < prm.max_moves ? T : prm.max_moves;
        int p_cur = piece_at(t), p_next = piece_at(t + 1u);
        const uint4* dq = reinterpret_cast<const uint4*>(&S.qd[p_cur < 7 ? p_cur : 0][myslot][part]);
        uint4 dA = dq[0], dB = dq[1];
        uint32_t par = 0;
        uint32_t fin_seen = 0;
        if (staged) { if (threadIdx.x == 0) S.fin[0] = *reinterpret_cast<volatile const uint32_t*>(finished); __syncthreads(); fin_seen = S.fin[0]; __syncthreads(); }
        bool handed = staged && prm.st_cut != 0u && prm.n_games - fin_seen <= prm.st_cut;
        bool over = false;

        while (!handed && t < stop && p_cur < 7) {
            if (t + 2u - base >= 64u) {              // advance the piece window
                base += 32u;
                win = (win >> 8) | load_piece(base + 32u + (uint32_t)lane) << 8;
            }
            const int p_after = piece_at(t + 2u);
            const uint4* dqn = reinterpret_cast<const uint4*>(&S.qd[p_next < 7 ? p_next : 0][myslot][part]);
            const uint4 dA_next = dqn[0], dB_next = dqn[1];
            const int ns = (int)(dA.w >> 16);
            const int cells = cells0 + 4 * (int)t - 10 * (int)(lines - lines0);      // filled cells on the board before this piece
            uint32_t hi = 0u, lo = 0u, wy = 0u, wfull = 0u;
            if (warp * QGROUPS < ns) {               // this warp's slot range is not empty for this piece (uniform per warp)
                // ---- landing row and legality from the packed heights (every lane) ----
                const uint32_t sel = dA.w & 0xffffu;
                const uint32_t pair = sel >> 8;
                const uint32_t wlo = pair == 0 ? H.H0 : (pair == 1 ? H.H1 : H.H2);
                const uint32_t whi = pair == 0 ? H.H1 : (pair == 1 ? H.H2 : 0u);
                const uint32_t hw = funnel_r(wlo, whi, sel & 31u);
                bool legal = ((hw + dA.z) & 0x80808080u) == 0u && slot < ns && !shadow;
                const uint32_t v = hw + dA.y;
                uint32_t m01 = byte_of(v, 0), b1 = byte_of(v, 1), b2 = byte_of(v, 2), b3 = v >> 24;
                m01 = m01 > b1 ? m01 : b1; b2 = b2 > b3 ? b2 : b3; m01 = m01 > b2 ? m01 : b2;
                const int y = (int)m01 - 32;
                // ---- imprint the own columns (OR == ADD: piece cells never overlap board cells) ----
                const uint32_t one_y = 1u << (y & 31);
                uint32_t n0 = byte_of(dB.x, 0) * one_y + q0, n1 = byte_of(dB.x, 1) * one_y + q1, n2 = byte_of(dB.x, 2) * one_y + q2,
                         n3 = (dB.x >> 24) * one_y + q3, n4 = (dB.y & 255u) * one_y + q4;
                // ---- full rows: AND over my own columns, then over the three lanes of the group ----
                const uint32_t fpart = n0 & n1 & n2 & (n3 | ~own3);
                const uint32_t f1 = __shfl_sync(FULLMASK, fpart, src1), f2 = __shfl_sync(FULLMASK, fpart, src2);
                uint32_t full = fpart & f1 & f2;
                // ---- heights of the view columns when no row is completed ----
                const int yb = y - 32;
                int g0 = (int)byte_of(dB.z, 0) + yb, g1 = (int)byte_of(dB.z, 1) + yb, g2 = (int)byte_of(dB.z, 2) + yb, g3h = (int)(dB.z >> 24) + yb,
                    g4 = (int)byte_of(dB.w, 0) + yb, g5 = (int)byte_of(dB.w, 1) + yb;
                g0 = g0 > hv0 ? g0 : hv0; g1 = g1 > hv1 ? g1 : hv1; g2 = g2 > hv2 ? g2 : hv2; g3h = g3h > hv3 ? g3h : hv3;
                g4 = g4 > hv4 ? g4 : hv4; g5 = g5 > hv5 ? g5 : hv5;
                int k = 0, eroded = 0;
                if (legal && full != 0u) {           // rows completed by this placement: clear my columns, heights by clz
                    k = popc(full);
                    const uint32_t fr = full >> y, rc = dA.x >> 12;
                    eroded = (int)(((fr & 1u) ? (rc & 7u) : 0u) + ((fr & 2u) ? ((rc >> 3) & 7u) : 0u) +
                                   ((fr & 4u) ? ((rc >> 6) & 7u) : 0u) + ((fr & 8u) ? ((rc >> 9) & 7u) : 0u));
                    uint32_t f = full;
                    do {
                        const uint32_t low = (1u << ctz(f)) - 1u;
                        n0 = (n0 & low) | ((n0 >> 1) & ~low); n1 = (n1 & low) | ((n1 >> 1) & ~low); n2 = (n2 & low) | ((n2 >> 1) & ~low);
                        n3 = (n3 & low) | ((n3 >> 1) & ~low); n4 = (n4 & low) | ((n4 >> 1) & ~low);
                        f = (f & low) | ((f >> 1) & ~low);
                    } while (f != 0u);
                    g1 = fls(n0); g2 = fls(n1); g3h = fls(n2); g4 = fls(n3); g5 = fls(n4);
                    g4 = g4 > wall4 ? g4 : wall4;
                }
                // left neighbour's height after a clear comes from the lane that owns that column (one shuffle; the three lanes
                // of a group take the branch above together, legality and the full mask being the same in all of them)
                {
                    const int send = part == 0 ? g4 : g3h;             // height of my last own column
                    const int recv = __shfl_sync(FULLMASK, send, lane > 0 ? lane - 1 : 0);
                    if (k != 0) g0 = part == 0 ? ROWS : recv;
                }
                // ---- wells (features.py:106-125), heights sum, transitions of my range ----
                int wsum, w2sum, hsum;
                {
                    int a = (g0 < g2 ? g0 : g2) - g1; a = a > 0 ? a : 0;
                    int b = (g1 < g3h ? g1 : g3h) - g2; b = b > 0 ? b : 0;
                    int cc = (g2 < g4 ? g2 : g4) - g3h; cc = cc > 0 ? cc : 0;
                    int dd = (g3h < g5 ? g3h : g5) - g4; dd = dd > 0 ? dd : 0;
                    dd = own3 ? dd : 0;
                    wsum = a + b + cc + dd; w2sum = a * a + b * b + cc * cc + dd * dd;
                    hsum = g1 + g2 + g3h + (own3 ? g4 : 0);
                }
                const int rt = popc(n0 ^ n1) + popc(n1 ^ n2) + popc((n2 ^ n3) & pm2) + popc((n3 ^ n4) & own3);        // features.py:89-91
                const int ct = popc((n0 ^ (n0 * 2u)) & 0xFFFFEu) + popc((n1 ^ (n1 * 2u)) & 0xFFFFEu) + popc((n2 ^ (n2 * 2u)) & 0xFFFFEu) +
                               popc((n3 ^ (n3 * 2u)) & 0xFFFFEu & own3);                                            // features.py:83-85
                // ---- the three lanes meet: packed partial sums, one shuffle round ----
                const uint32_t pa = (uint32_t)rt | (uint32_t)ct << 8 | (uint32_t)hsum << 16, pb = (uint32_t)(wsum + w2sum);
                const uint32_t sa = pa + __shfl_sync(FULLMASK, pa, src1) + __shfl_sync(FULLMASK, pa, src2);
                const uint32_t sb = pb + __shfl_sync(FULLMASK, pb, src1) + __shfl_sync(FULLMASK, pb, src2);
                const int RT = (int)(sa & 255u), CT = (int)((sa >> 8) & 255u), SH = (int)(sa >> 16);
                const int pits = SH - (cells + 4 - 10 * k);                                                         // features.py:146-160
                const int cw = (int)(sb >> 1);
                double acc = 0.0;                                                                                    // canonical order
                acc = dadd(acc, dmul(u2d(y), w[0]));
                acc = dadd(acc, dmul(u2d(eroded), w[1]));
                acc = dadd(acc, dmul(u2d(RT), w[2]));
                acc = dadd(acc, dmul(u2d(CT), w[3]));
                acc = dadd(acc, dmul(u2d(pits), w[4]));
                acc = dadd(acc, dmul(u2d(cw), w[5]));
                const uint64_t key = legal ? score_key(acc) : 0ull;
                cnt += (uint32_t)(legal && part == 0);
                // ---- warp-local argmax over the part-0 lanes ----
                const uint32_t khi = part == 0 ? (uint32_t)(key >> 32) : 0u;
                const uint32_t mhi = __reduce_max_sync(FULLMASK, khi);
                const uint32_t tie = __ballot_sync(FULLMASK, khi == mhi);
                int src = __ffs((int)tie) - 1;
                if (mhi != 0u && (tie & (tie - 1u))) {
                    const uint32_t klo = (uint32_t)key;
                    const uint32_t mlo = __reduce_max_sync(FULLMASK, khi == mhi ? klo : 0u);
                    const uint32_t mtag = __reduce_min_sync(FULLMASK, (khi == mhi && klo == mlo) ? (uint32_t)lane : 0xffffffffu);
                    src = (int)mtag;
                }
                if (lane == src) { hi = mhi; lo = (uint32_t)key; wy = (uint32_t)slot | (uint32_t)y << 8; wfull = full; }
                if (lane == src) S.rec[par][warp] = make_uint4(hi, lo, wy, wfull);
            } else if (lane == 0) {
                S.rec[par][warp] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (staged && threadIdx.x == 0 && (t & 3u) == 0u) S.fin[par] = *reinterpret_cast<volatile const uint32_t*>(finished);
            __syncthreads();
            // ---- global winner of the four warps: higher key, then lower slot (= lower warp) ----
            uint4 best = S.rec[par][0];
            {
                const uint4 r1 = S.rec[par][1], r2 = S.rec[par][2], r3 = S.rec[par][3];
                uint64_t kb = ((uint64_t)best.x << 32) | best.y;
                const uint64_t k1 = ((uint64_t)r1.x << 32) | r1.y, k2 = ((uint64_t)r2.x << 32) | r2.y, k3 = ((uint64_t)r3.x << 32) | r3.y;
                if (k1 > kb) { kb = k1; best = r1; }
                if (k2 > kb) { kb = k2; best = r2; }
                if (k3 > kb) { kb = k3; best = r3; }
            }
            if (staged && (t & 3u) == 0u) fin_seen = S.fin[par];
            par ^= 1u;
            if (best.x == 0u) { over = true; break; }        // no legal placement in any warp: game over (simulator.py:187-189)
            // ---- commit: every lane re-applies the winning placement to its own columns ----
            const int wslot = (int)(best.z & 255u), y = (int)(best.z >> 8);
            const uint32_t full = best.w;
            const uint4 dW = reinterpret_cast<const uint4*>(&S.qd[p_cur][wslot][part])[1];
            const uint32_t misc = S.qd[p_cur][wslot][0].e[0];
            {
                const uint32_t one_y = 1u << (y & 31);
                q0 += byte_of(dW.x, 0) * one_y; q1 += byte_of(dW.x, 1) * one_y; q2 += byte_of(dW.x, 2) * one_y;
                q3 += (dW.x >> 24) * one_y; q4 += (dW.y & 255u) * one_y;
                const int yb = y - 32;
                int g0 = (int)byte_of(dW.z, 0) + yb, g1 = (int)byte_of(dW.z, 1) + yb, g2 = (int)byte_of(dW.z, 2) + yb, g3h = (int)(dW.z >> 24) + yb,
                    g4 = (int)byte_of(dW.w, 0) + yb, g5 = (int)byte_of(dW.w, 1) + yb;
                hv0 = g0 > hv0 ? g0 : hv0; hv1 = g1 > hv1 ? g1 : hv1; hv2 = g2 > hv2 ? g2 : hv2; hv3 = g3h > hv3 ? g3h : hv3;
                hv4 = g4 > hv4 ? g4 : hv4; hv5 = g5 > hv5 ? g5 : hv5;
            }
            uint32_t wk = 0u;
            if (full != 0u) {                                  // block-uniform
                wk = (uint32_t)popc(full);
                uint32_t f = full;
                do {
                    const uint32_t low = (1u << ctz(f)) - 1u;
                    q0 = (q0 & low) | ((q0 >> 1) & ~low); q1 = (q1 & low) | ((q1 >> 1) & ~low); q2 = (q2 & low) | ((q2 >> 1) & ~low);
                    q3 = (q3 & low) | ((q3 >> 1) & ~low); q4 = (q4 & low) | ((q4 >> 1) & ~low);
                    f = (f & low) | ((f >> 1) & ~low);
                } while (f != 0u);
                hv1 = fls(q0); hv2 = fls(q1); hv3 = fls(q2); hv4 = fls(q3); hv5 = fls(q4);
                hv4 = hv4 > wall4 ? hv4 : wall4;
                if (part == 2) hv5 = ROWS;
                const int send = part == 0 ? hv4 : hv3;
                const int recv = __shfl_sync(FULLMASK, send, lane > 0 ? lane - 1 : 0);
                hv0 = part == 0 ? ROWS : recv;
                lines += wk;
                h1 += wk == 1; h2 += wk == 2; h3 += wk == 3; h4 += wk == 4;
                const uint32_t pts = prm.pts[wk];
                gscore += (uint64_t)pts * (uint64_t)(lines / 10u + 1u);
            } else if (prm.pts[0] != 0u) gscore += (uint64_t)prm.pts[0] * (uint64_t)(lines / 10u + 1u);      // State.score(points) with points[0] != 0
            // packed heights of the whole board, rebuilt from the three parts of my group
            {
                const uint32_t hp = (uint32_t)hv1 | (uint32_t)hv2 << 8 | (uint32_t)hv3 << 16 | (uint32_t)hv4 << 24;
                H.H0 = __shfl_sync(FULLMASK, hp, g3);
                H.H1 = __shfl_sync(FULLMASK, hp, g3 + 1);
                H.H2 = (__shfl_sync(FULLMASK, hp, g3 + 2) >> 8) & 0xffffu;
            }
            if ((prm.trace != nullptr || prm.score_trace != nullptr) && threadIdx.x == 0) {
                if (prm.trace) {
                    const uint32_t wph = (misc >> 9) & 7u;
                    const uint32_t row = (uint32_t)ROWS - (uint32_t)y - wph;
                    reinterpret_cast<uint32_t*>(prm.trace)[(size_t)game * T + t] = ((misc >> 4) & 3u) | row << 8 | (misc & 15u) << 16 | wk << 24;
                }
                if (prm.score_trace) {
                    const uint64_t kk = ((uint64_t)best.x << 32) | best.y;
                    const uint64_t b = (kk & 0x8000000000000000ull) ? (kk & 0x7FFFFFFFFFFFFFFFull) : ~kk;
                    prm.score_trace[(size_t)game * T + t] = __longlong_as_double((long long)b);
                }
            }
            ++t;
            p_cur = p_next; p_next = p_after;
            dA = dA_next; dB = dB_next;
            if (staged && prm.st_cut != 0u) handed = prm.n_games - fin_seen <= prm.st_cut;
        }
        handed = handed && !over && t < stop && p_cur < 7;       // a game that ended anyway is finished, not handed over
        cnt = __reduce_add_sync(FULLMASK, cnt);
        // the whole board lives in lanes 0, 1, 2 of every warp
        if (warp == 0 && lane < 3) {
            uint32_t* dst = nullptr;
            if (handed) dst = prm.st_save[game].c;
            if (dst) {
                if (part == 0) { dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3; }
                else if (part == 1) { dst[4] = q0; dst[5] = q1; dst[6] = q2; }
                else { dst[7] = q0; dst[8] = q1; dst[9] = q2; }
            }
            if (!handed && prm.final_rows) {
                // final board: gather the columns in lane 0 through shared memory is not needed — each lane ORs its own cells into the rows
                for (int r2 = 0; r2 < ROWS; ++r2) {
                    const int i = ROWS - 1 - r2;
                    uint32_t v = 0;
                    if (part == 0) v = ((q0 >> i) & 1u) | ((q1 >> i) & 1u) << 1 | ((q2 >> i) & 1u) << 2 | ((q3 >> i) & 1u) << 3;
                    else if (part == 1) v = ((q0 >> i) & 1u) << 4 | ((q1 >> i) & 1u) << 5 | ((q2 >> i) & 1u) << 6;
                    else v = ((q0 >> i) & 1u) << 7 | ((q1 >> i) & 1u) << 8 | ((q2 >> i) & 1u) << 9;
                    v |= __shfl_sync(7u, v, 1) | __shfl_sync(7u, v, 2);          // lane 0 ends up with all ten bits
                    if (lane == 0) prm.final_rows[(size_t)game * ROWS + r2] = (uint16_t)v;
                }
            }
        }
        if (lane == 0 && cnt) {
            if (handed) atomicAdd(&prm.st_save[game].cnt, cnt);
            else atomicAdd(reinterpret_cast<unsigned long long*>(&prm.results[game].placements), (unsigned long long)cnt);
        }
        if (threadIdx.x == 0) {
            if (handed) {
                GameSave* sv = prm.st_save + game;
                sv->t = t; sv->lines = lines; sv->h1 = h1; sv->h2 = h2; sv->h3 = h3; sv->h4 = h4; sv->gscore = gscore;
                prm.st_list_out[atomicAdd(prm.st_ctl + 9 + prm.st_stage, 1u)] = game;
            } else {
                if (staged) atomicAdd(prm.st_ctl + 4, 1u);
                tb200_game_result* r = prm.results + game;
                r->pieces = t; r->lines = lines; r->hist[0] = h1; r->hist[1] = h2; r->hist[2] = h3; r->hist[3] = h4;
                r->score = gscore;
            }
        }
        __syncthreads();
    }
}

}  // namespace tb
