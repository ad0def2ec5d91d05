// play_common.cuh — launch parameters, hand-over state and the argmax / table helpers shared by the play kernels (part of libtetris_b200.so; included by tetris_b200.cu only).
#pragma once

#include "tetris_core.cuh"
#include "../../include/tetris_b200.h"

namespace tb {

#define FULLMASK 0xffffffffu
constexpr int BLOCK = 128;

__constant__ Tables c_tables;
// The same tables in global memory: every block copies them into shared memory at kernel start, and a copy out of constant
// memory with per-thread addresses is serialised by the constant cache (it showed up as 2.5 % of the samples of a 0.4 ms
// kernel); from global memory it is a few coalesced 128-bit loads per thread.
__device__ Tables g_tables;
// board-independent window constants of every slot (tetris_core.cuh: win_desc), for the window-local lockstep kernel
__constant__ WinDesc c_win[7][TB_MAX_SLOTS];
static_assert(sizeof(Tables) % 16 == 0, "Tables must be a whole number of 128-bit words");
template <int NT>
__device__ __forceinline__ void load_tables(Tables& s_tab)
{
    const uint4* src = reinterpret_cast<const uint4*>(&g_tables);
    uint4* dst = reinterpret_cast<uint4*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 16); i += NT) dst[i] = src[i];
}
#ifdef TB_PHASES
__device__ unsigned long long g_phase[8];
#define PH(i) { const long long now_ = clock64(); ph[i] += now_ - last_; last_ = now_; }
#else
#define PH(i)
#endif

struct PlayParams {
    const double*  weights;
    const uint8_t* pieces;
    const int32_t* seq_of_game;
    const uint16_t* init_rows;
    const uint32_t* init_lines;      // lines already cleared before the first move (level for score / pressure), or null
    tb200_game_result* results;
    uint8_t*  trace;
    double*   score_trace;
    uint16_t* final_rows;
    uint32_t* ticket;
    uint64_t  seed;
    uint64_t  rmask;
    uint32_t  n_games;
    uint32_t  T;
    uint32_t  max_moves;
    int32_t   G;
    int32_t   F;
    int32_t   per_game_weights;      // K1 (one game per board, sequence id = game): weights row = game (1) or row 0 (-1); K2: 0
    uint32_t  seq_major_C;           // kernels with several games per warp: != 0 => tickets run sequence-major (ticket = g * C + candidate), so the games that share a
                                     // warp start on the same piece sequence and stay in step while they live (fewer idle lanes)
    int32_t   move_set;              // bit 0: overhang slides, bit 1: pressure filter (extended generic kernels only)
    Pressure  press;                 // move_with_pressure parameters (press.on mirrors bit 1)
    uint8_t   ids[TB200_MAX_WEIGHTED];
    uint32_t  pts[5];                // State.score(points) table (game.py:159), default (0, 40, 100, 300, 1200)
    // re-spreading stages of the latency regime (K2f / K2p only; st_ctl == null: off) — see launch_play
    uint32_t* st_ctl;                // [0..3] ticket of stage s, [4] games finished so far, [8 + s] games handed to stage s
    const uint32_t* st_list_in;      // game ids handed to this stage, or null: all games, fresh
    uint32_t* st_list_out;           // survivors of this stage (null in the last stage)
    struct GameSave* st_save;        // [n_games] state of the handed-over games
    int32_t   st_stage;
    uint32_t  st_cut;                // leave the stage once (games not yet finished) <= st_cut; 0: run to the end
    // lockstep kernel K2s (see play_games_w6lock_kernel)
    uint32_t* lk_ctl;                // [S][G][4] per (stage, sequence) bucket: {task ticket, tasks done, games entering the stage, -}
    uint32_t* lk_list;               // [2][G][C] game ids entering a stage, by sequence; double-buffered by stage parity
    uint32_t  lk_C, lk_K, lk_S;      // candidates, pieces per stage, stages
    uint32_t  lk_thr[5];             // K2sA: games alive per resident warp from which a bucket plays with 1 / 4 / 8 lanes per game (below: a warp per game);
                                     // [3]: from which it plays with one lane per game in each of two warps (K2sP; lk_thr[1] <= lk_thr[3] <= lk_thr[0]);
                                     // [4]: the one-lane threshold of a STEADY bucket (>= 15/16 of the previous stage's games still alive)
    // piece-binned kernel K2b (see play_games_w6bin_kernel)
    unsigned long long* pb_ctl;      // per shard 16 words: [0..6] queue tails (reserved), [8..14] queue heads; after the shards: games finished, error flag
    uint32_t* pb_ring;               // [shards][7][pb_mask + 1] game ids waiting for a piece of type p; 0xffffffff = empty slot
    uint32_t  pb_mask;               // ring capacity - 1 (a power of two >= games per shard)
};

// state of a game handed from one stage to the next (96 bytes)
struct GameSave {
    uint32_t c[COLS];
    uint32_t t, lines, h1, h2, h3, h4, cnt, pad;
    uint64_t gscore;
    uint32_t H0, H1, H2, cells;      // K2b only: packed column heights and filled cells of the board (saves recomputing them every piece)
};

// ---- argmax over a LANES-wide group: maximise key, ties -> smallest tag (first in enumeration order)
template <int LANES>
TB_D void group_argmax(uint64_t& key, uint32_t& tag)
{
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
        const uint32_t ohi = __shfl_xor_sync(FULLMASK, (uint32_t)(key >> 32), off);
        const uint32_t olo = __shfl_xor_sync(FULLMASK, (uint32_t)key, off);
        const uint32_t otag = __shfl_xor_sync(FULLMASK, tag, off);
        const uint64_t okey = ((uint64_t)ohi << 32) | olo;
        if (okey > key || (okey == key && otag < tag)) { key = okey; tag = otag; }
    }
}

// same, carrying the placement to commit (mv = slot | height << 8) along with the winner
template <int LANES>
TB_D void group_argmax_mv(uint64_t& key, uint32_t& tag, uint32_t& mv)
{
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
        const uint32_t ohi = __shfl_xor_sync(FULLMASK, (uint32_t)(key >> 32), off);
        const uint32_t olo = __shfl_xor_sync(FULLMASK, (uint32_t)key, off);
        const uint32_t otag = __shfl_xor_sync(FULLMASK, tag, off);
        const uint32_t omv = __shfl_xor_sync(FULLMASK, mv, off);
        const uint64_t okey = ((uint64_t)ohi << 32) | olo;
        if (okey > key || (okey == key && otag < tag)) { key = okey; tag = otag; mv = omv; }
    }
}

TB_D SlotDesc load_slot12(const Tables& tab, int piece, int s)
{
    const uint4* p = reinterpret_cast<const uint4*>(&tab.slot[piece][s]);
    const uint4 a = p[0], b = p[1], e = p[2];
    SlotDesc d;
    d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w; d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
    d.w[8] = e.x; d.w[9] = e.y; d.w[10] = e.z; d.w[11] = e.w;
    return d;
}

// Moves derived from one legal hard-drop slot: the drop itself (index 0) and, with SLIDE, the overhang slides
// (simulator.py:15-47) in the reference's order.  list[i] = slot | height << 8 of slide i + 1.
// ---------------------------------------------------------------------------------------------
// K2w: the latency-optimised variant for lookahead 1 with one full warp per game.  Used when the
// population fits in one resident wave (BASELINE config 2: 1000 games), where the run time is the
// serial chain of the longest game, so what matters is the per-piece latency of a single warp:
//   * slots s and s+32 of a 34-slot piece (T, L, J) are evaluated as two interleaved straight-line
//     instruction streams in ONE pass instead of two dependent passes,
//   * argmax by three REDUX (hi word, lo word, tag) instead of a 5-step shuffle butterfly,
//   * commit by broadcasting the winning lane's candidate board (SHFL) instead of recomputing it,
//   * the ticket fetch is outside the per-piece loop.
// ---------------------------------------------------------------------------------------------
TB_D void warp_argmax32(uint64_t& key, uint32_t& tag)
{
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mhi = __reduce_max_sync(FULLMASK, hi);
    const uint32_t mlo = __reduce_max_sync(FULLMASK, hi == mhi ? lo : 0u);
    const bool win = hi == mhi && lo == mlo;
    tag = __reduce_min_sync(FULLMASK, win ? tag : 0xffffffffu);
    key = ((uint64_t)mhi << 32) | mlo;
}

template <int MODE>
TB_D uint64_t eval_key(const Placement& P, const SlotDesc& d, bool legal, const PrevStat& pv, const uint8_t* ids, const double* w, int F, uint64_t rmask)
{
    Feat f;
    eval_features<ModeTraits<MODE>::CMASK>(P, d, pv, rmask, f);
    const uint64_t k = score_key(score_features<MODE>(f, ids, w, F));
    return legal ? k : 0ull;
}

}  // namespace tb
