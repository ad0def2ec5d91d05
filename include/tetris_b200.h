/* tetris_b200.h — C ABI of the B200-native batched Tetris placement engine (libtetris_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of CogWorks/Tetris-AI: for every live game of a
 * cross-entropy population, enumerate each rotation x column placement of the current piece
 * (optionally also of the next piece), drop it, clear lines, compute the board features, score
 * them against the candidate's weight vector in fp64 and commit the argmax move.
 *
 * The reference has no FFI of its own — its extension points are Python callables — so each entry
 * point below names the reference callable(s) whose work it takes over (paths are under
 * /root/reference/cogworks/):
 *
 *   tb200_play_games     simulate()            tetris/simulator.py:152-189  (whole games)
 *                        move_drop()           tetris/simulator.py:3-13
 *                        policy_best()         tetris/simulator.py:123-150  (first-of-max tie rule)
 *                        State.future()        tetris/game.py:166-180  (+ Board.imprint/full/clear :62-87)
 *                        State.lines_cleared/level/score   tetris/game.py:150-164
 *                        the test_f / map_f call of cross_entropy()   learning.py:38
 *   tb200_best_move(_ex) one simulate() iteration: pool_stack + move_policy   simulator.py:156-180 (any move_gen of simulator.py:3-120)
 *   tb200_eval_features  feature.evaluate(state, features)   feature.py:50-55 over tetris/features.py:18-347
 *   tb200_hashed_pieces  the caller-supplied zoid_gen (simulator.py:152) as a counter-based stream
 *
 * Conventions
 *   - every function returns 0 on success, a negative TB200_E_* code otherwise; nothing throws
 *     across this boundary; tb200_last_error(ctx) returns a message for the last failure;
 *   - the caller owns every buffer; the library never frees caller memory and keeps no caller
 *     pointer after a call returns;
 *   - `mem` says where ALL buffers of a call live: TB200_MEM_HOST (the library stages them through
 *     its own device buffers and the call is synchronous) or TB200_MEM_DEVICE (pointers are device
 *     pointers; work is enqueued on `stream`, the call returns without synchronising);
 *   - a ctx is bound to one device, is not thread-safe, and different ctxs are independent;
 *   - ONE CALL IN FLIGHT PER CTX: every call on a ctx shares one device workspace (game tickets, hand-over lists, queues, the
 *     staging arena).  TB200_MEM_HOST calls return synchronised.  TB200_MEM_DEVICE calls only enqueue; the library orders
 *     them itself — a later call on the same ctx, on any stream or in host mode, first waits (on the device, by event) for
 *     the previous call's last kernel — so two calls never share the workspace, but they also never overlap.  Use one ctx
 *     per stream for concurrency.  tb200_check_status(ctx) reports a bounded wait that expired inside an enqueued call;
 *   - there is NO CPU fallback: if no CUDA device is usable tb200_create fails with TB200_E_NO_DEVICE.
 *
 * Board exchange format (same as the reference's numpy grid, packed): uint16_t rows[20], rows[0]
 * is the TOP row, bit c of a row is column c.  Piece ids: I,T,L,J,O,Z,S = 0..6 (tetris/game.py:24-34).
 * Feature ids: 0..44 in tetris/features.py source order (TB200_F_* and TB200_FEATURE_NAMES in tetris_b200_features.h).
 *
 * Limits (refused with TB200_E_BAD_ARG, never approximated): lookahead 1..TB200_MAX_LOOKAHEAD (simulate() accepts any
 * lookahead > 0, simulator.py:153; the configs use 1 and 2); boards are 20 x 10 (game.py:38 takes any size; every config
 * uses 20 x 10); weights must be finite with |w| <= 1e300 so that no score is NaN or infinite (all_max's NaN ordering,
 * simulator.py:130-134, is therefore never exercised; checked for TB200_MEM_HOST, required for TB200_MEM_DEVICE).
 */
#ifndef TETRIS_B200_H
#define TETRIS_B200_H

#if !defined(__CUDACC_RTC__)      /* run-time compilation of the kernels has no host headers; tetris_core.cuh supplies the types */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_ROWS 20
#define TB200_COLS 10
#define TB200_NUM_FEATURES 45
#define TB200_MAX_WEIGHTED 48
#define TB200_MAX_LOOKAHEAD 2

#define TB200_MOVES_DROP            0
#define TB200_MOVES_OVERHANG_SLIDE  1
#define TB200_MOVES_PRESSURE        2
#define TB200_MOVES_WIGGLE          4      /* move_wiggle: flood fill from the top of every column (simulator.py:81-85) */
#define TB200_MOVES_WIGGLE_DROP     8      /* move_with_wiggle(move_drop) (simulator.py:49-77) */

#define TB200_MEM_HOST   0
#define TB200_MEM_DEVICE 1

#define TB200_OK             0
#define TB200_E_NO_DEVICE   -1
#define TB200_E_BAD_ARG     -2
#define TB200_E_CUDA        -3
#define TB200_E_NO_FEATURES -4
#define TB200_E_ALLOC       -5

typedef struct tb200_ctx tb200_ctx;

/* per-game result record, written once per game (40 bytes) */
typedef struct {
    uint32_t pieces;        /* pieces placed (game length)                         simulator.py:180 */
    uint32_t lines;         /* State.lines_cleared()                               game.py:150-154  */
    uint32_t hist[4];       /* State.cleared[k], k = 1..4                          game.py:176      */
    uint64_t score;         /* State.score() with the default points table          game.py:159-164  */
    uint64_t placements;    /* scored leaves = scorer calls                        simulator.py:129 */
} tb200_game_result;

typedef struct {
    uint32_t struct_size;            /* = sizeof(tb200_play_args) */
    int32_t  mem;                    /* TB200_MEM_HOST / TB200_MEM_DEVICE for every pointer below */
    void*    stream;                 /* cudaStream_t for TB200_MEM_DEVICE (NULL = legacy default stream) */
    int32_t  n_candidates;           /* C */
    int32_t  games_per_candidate;    /* G; game g belongs to candidate g / G */
    const double*  weights;          /* f64[C][F], F and the order from tb200_set_features (learning.py:32-35) */
    const uint8_t* pieces;           /* u8[S][T] piece ids (an id > 6 ends the game there), or NULL => counter-based stream (seed, stream id, t) */
    int32_t  n_sequences;            /* S */
    uint32_t T;                      /* pieces per sequence = cap on the game length */
    uint32_t max_moves;              /* stop (not game over) after this many placements; 0 => T */
    const int32_t* seq_of_game;      /* i32[C*G] sequence / stream id of each game, or NULL => g % G */
    uint64_t seed;                   /* seed of the counter-based stream */
    int32_t  lookahead;              /* 1 or 2 (simulate(..., lookahead)) */
    int32_t  lanes;                  /* lanes per game: 0 = auto, or 8 / 16 / 32; for the w6 set with lookahead 1 also 64 (two warps per game), 4, 2, 1.
                                        An explicit value runs ONE persistent kernel of that width.  0 lets the library schedule the play: for the w6
                                        set with lookahead 1 that can be several stream-ordered launches (2 - 16 games per SM), the lockstep kernel
                                        (shared sequences, > 80 games per SM) or the piece-binned kernel (>= 32 768 games); results never depend on it.
                                        Those two kernels bound every wait: a TB200_MEM_HOST call returns TB200_E_CUDA instead of hanging if one expires. */
    const uint16_t* init_rows;       /* u16[C*G][20] initial boards, or NULL => empty boards */
    tb200_game_result* results;      /* [C*G], required */
    uint8_t*  trace;                 /* u8[C*G][T][4] = {rot,row,col,k} per move, or NULL */
    double*   score_trace;           /* f64[C*G][T] winning score per move, or NULL */
    uint16_t* final_rows;            /* u16[C*G][20] final boards, or NULL */
    double*   fitness;               /* f64[C] = total lines of the candidate's games / G, or NULL */
    int32_t   move_set;              /* TB200_MOVES_DROP = move_drop (simulator.py:3-13), optionally OR-ed with
                                        TB200_MOVES_OVERHANG_SLIDE: move_with_overhang_slide(...) (simulator.py:15-47), or with
                                        TB200_MOVES_WIGGLE / TB200_MOVES_WIGGLE_DROP (simulator.py:49-85), and / or
                                        TB200_MOVES_PRESSURE: move_with_pressure(...) as the OUTERMOST wrapper (simulator.py:88-120) */
    int32_t   pressure_two_but_rot;  /* move_with_pressure(two_but_rot=...) */
    double    pressure_avg_lat;      /* move_with_pressure(avg_lat=250.9, resp_time=79.8, move_eff=1.23) — no defaults here */
    double    pressure_resp_time;
    double    pressure_move_eff;
    const uint32_t* init_lines;      /* u32[C*G] lines already cleared before the first move (continuing games: the level used
                                        by score and by the pressure filter; results.lines then reports the running total), or NULL */
    uint32_t  points[5];             /* State.score(points) (game.py:159-164): results.score = sum over moves of points[rows cleared by
                                        the move] * (level after the move + 1).  All five zero => the reference's default table
                                        (0, 40, 100, 300, 1200).  points[0] may be non-zero (every non-clearing move then scores). */
    uint32_t  reserved0;
} tb200_play_args;

/* timing of the last tb200_play_games call on this ctx (CUDA events on the launching stream) */
typedef struct {
    float kernel_ms;                 /* play kernel only (valid after the stream has been synchronised; 0 for a CUDA-graph replay) */
    float total_ms;                  /* host mode: H2D + kernels + D2H */
    int32_t lanes;                   /* lanes per game actually used */
    int32_t grid, block;             /* launch configuration */
    int32_t kernels_launched;        /* number of kernels of this library launched by the call */
    uint64_t h2d_bytes, d2h_bytes;   /* host mode: bytes staged each way */
    int32_t specialised;             /* 1: the play kernel was one compiled at run time for the ctx's feature list (tb200_jit_build) */
    int32_t reserved;
} tb200_timing;

/* page-locked host memory for staging buffers (optional).  A TB200_MEM_HOST play_games call whose buffers are
 * all pinned and that repeats the previous call's arguments (one CE generation after another) is replayed as a
 * CUDA graph: one launch instead of a dozen runtime calls.  Set TB200_NO_GRAPHS=1 to disable. */
int  tb200_alloc_pinned(size_t bytes, void** out);
void tb200_free_pinned(void* p);

int  tb200_create(int device, tb200_ctx** out);
void tb200_destroy(tb200_ctx* ctx);
const char* tb200_last_error(const tb200_ctx* ctx);
const char* tb200_version(void);

/* number of SMs / SM clock (kHz) / device name of the ctx's device */
int tb200_device_info(tb200_ctx* ctx, int32_t* sms, int32_t* clock_khz, char* name, int32_t name_len);

/* weighted features of the linear scorer, in summation order (the key order of the weights
 * mapping handed to test_f, learning.py:38).  1 <= F <= TB200_MAX_WEIGHTED, ids in 0..44. */
int tb200_set_features(tb200_ctx* ctx, const int32_t* feature_ids, int32_t F);

/* Feature lists other than the two compiled into the library (the w6 set, all 45 in id order) are played by kernels that
 * tb200_play_games specialises for the list at run time — the library's own CUDA sources (csrc/, next to the .so), compiled
 * once more with NVRTC for sm_100a with the list known to the compiler, cached on disk ($TB200_JIT_CACHE, default
 * /tmp/tetris_b200_jit-<uid>) — about 4x faster than the run-time-list code, identical results.  It happens on the first play
 * of >= 2^22 game-moves with lookahead 1 and hard drops (~8 s once; TB200_JIT=1 always, =0 never); if NVRTC or the sources are
 * missing the precompiled run-time-list kernels run instead (tb200_last_error says why).
 * tb200_jit_build compiles (or finds in the cache) the kernels for a list without needing a device: returns the cubin size
 * in KiB (> 0) or a negative TB200_E_* code; log (may be NULL) receives "compiled" / "cached" or the reason of the failure. */
int tb200_jit_build(const int32_t* feature_ids, int32_t F, char* log, int32_t log_len);

/* whole games for a population: C candidates x G games (K2) */
int tb200_play_games(tb200_ctx* ctx, const tb200_play_args* args);
int tb200_last_timing(tb200_ctx* ctx, tb200_timing* out);

/* Status of the work the last TB200_MEM_DEVICE call enqueued: waits for it (event of its last kernel), then returns TB200_OK, or
 * TB200_E_CUDA if a CUDA error surfaced or one of the bounded device-side waits of the lockstep / piece-binned schedulers
 * expired (results incomplete).  TB200_MEM_HOST calls check this themselves before they return. */
int tb200_check_status(tb200_ctx* ctx);

/* one placement decision per board (K1).  boards u16[N][20]; piece/next u8[N] (next[i] = 255 or
 * next == NULL => no visible second piece); weights f64[N][F] if per_board_weights else f64[1][F].
 * out_move u8[N][4] = {rot,row,col,k} (row = 255 => no legal move), out_score f64[N] (may be NULL),
 * out_count u32[N] scored leaves (may be NULL), out_rows u16[N][20] boards after the move (may be NULL). */
int tb200_best_move(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                    const uint16_t* boards, const uint8_t* piece, const uint8_t* next,
                    const double* weights, int32_t per_board_weights, int32_t lookahead,
                    uint8_t* out_move, double* out_score, uint32_t* out_count, uint16_t* out_rows);

/* the same decision under the other move generators of the reference (simulator.py:15-120): opts->move_set and the
 * pressure_* fields mean what they mean in tb200_play_args; opts->lines u32[N] = lines cleared so far on each board (the
 * level the pressure filter reads, game.py:156-157), or NULL = 0.  opts == NULL is tb200_best_move. */
typedef struct {
    uint32_t struct_size;            /* = sizeof(tb200_move_opts) */
    int32_t  move_set;
    int32_t  pressure_two_but_rot;
    int32_t  reserved;
    double   pressure_avg_lat;
    double   pressure_resp_time;
    double   pressure_move_eff;
    const uint32_t* lines;
} tb200_move_opts;
int tb200_best_move_ex(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                       const uint16_t* boards, const uint8_t* piece, const uint8_t* next,
                       const double* weights, int32_t per_board_weights, int32_t lookahead, const tb200_move_opts* opts,
                       uint8_t* out_move, double* out_score, uint32_t* out_count, uint16_t* out_rows);

/* all 45 features of prev.future(piece, rot, <drop row>, col) for N (board, placement) pairs (K3).
 * out_features f64[N][45] (ints are exact), out_info i32[N][3] = {legal, row, k}, out_rows u16[N][20]
 * post-clear boards (may be NULL). */
int tb200_eval_features(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                        const uint16_t* prev_boards, const uint8_t* piece, const uint8_t* rot, const uint8_t* col,
                        double* out_features, int32_t* out_info, uint16_t* out_rows);

/* same, for placements at an explicit row (row u8[N] = row of the piece's top bbox row, 255 => hard drop): the piece
 * only has to fit there — State.future(zoid, rot, row, col) (game.py:166) does not check support either — which
 * covers the states produced by the overhang-slide move set.  row may be NULL (= tb200_eval_features). */
int tb200_eval_features_at(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                           const uint16_t* prev_boards, const uint8_t* piece, const uint8_t* rot, const uint8_t* col, const uint8_t* row,
                           double* out_features, int32_t* out_info, uint16_t* out_rows);

/* materialise the counter-based piece stream (K4): out u8[n_streams][T], stream ids first_stream.. */
int tb200_hashed_pieces(tb200_ctx* ctx, int32_t mem, void* stream, uint64_t seed, uint64_t first_stream,
                        int32_t n_streams, uint32_t T, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* TETRIS_B200_H */
