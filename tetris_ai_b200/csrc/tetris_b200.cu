// tetris_b200.cu — the single translation unit of libtetris_b200.so (sm_100a only): small kernels (K3, K4, fitness, K1 packers),
// kernel selection, launch scheduling and the C ABI.  The play kernels live in the headers included below:
//   play_common.cuh      PlayParams, GameSave, argmax / table helpers
//   play_generic.cuh     K2 generic, K2w            play_latency.cuh     K2f, K2p
//   play_population.cuh  K2g, K2s, K2b              play_lookahead2.cuh  K2L          play_wiggle.cuh  K2x
//   tetris_core.cuh      the host/device core (placement, clear, 45 features, scorer) shared with the host emulation in tests/emu
//
//   K2  play_games_kernel<LANES,MODE,LOOKAHEAD,EXT>  persistent, one LANES-wide lane group per game (8 / 16 / 32), lookahead
//        1 or 2, any feature list; placements fan out across the lanes, features by popc/clz bit formulas
//        (tetris_core.cuh), fp64 dot product without FMA, shuffle argmax with the reference's first-of-max tie rule,
//        dynamic game tickets (atomic counter) for the heavy-tailed game lengths.  EXT = 1 adds the overhang-slide
//        move set and the pressure filter.
//   K2f play_games_w6fast_kernel        one warp per game, lookahead 1, the w6 feature set: the latency path (config 2).
//   K2p play_games_w6pair_kernel        two warps per game (tiny populations).
//   K2g play_games_w6grp_kernel<LANES>  16 / 8 / 4 / 2 / 1 lanes per game: the throughput path (configs 3, 5).
//   K2s play_games_w6lock_kernel<LANES> the lockstep population kernel (config 3): every warp holds games of one piece sequence at one
//        move index; stages of 64 pieces, survivors re-packed per sequence, one persistent launch with per-bucket hand-over.
//   K2b play_games_w6bin_kernel<LANES>  the piece-binned kernel (config 5, independent streams): games travel through device-wide
//        per-piece-type queues; a warp pops 32 / LANES games on the same piece, plays one piece each, pushes the survivors.
//   K2q play_games_w6quad_kernel        four warps per game, three lanes per placement (each a column range): the sparse stages of the latency regime.
//   K2f<staged> + K2q as re-spreading stages for the latency regime (config 2): see launch_play.
//   K2w play_games_w32_kernel<MODE>     one warp per game, lookahead 1, all-45 / generic feature lists.
//   K2L play_games_l2flat_kernel<MODE>  one warp per game, lookahead 2: layer-1 boards in shared memory, leaves flattened over lanes.
//   K2x play_games_wiggle_kernel<MODE>  the wiggle move sets (8f rank 4): warp-parallel set flood fill + scoring; the reference-order search only breaks ties.
//   K1 = K2* with per-game initial boards, two-piece sequences and max_moves = 1 (tb200_best_move).
//   K3 eval_features_kernel   one thread per (board, placement): all 45 features.
//   K4 hashed_pieces_kernel   counter-based piece stream.
//   fitness_kernel            per-candidate  sum(lines) / G.
//
// Reference lines replaced: see include/tetris_b200.h.  There is no CPU path in this file.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <new>

#include "play_common.cuh"
#include "play_generic.cuh"
#include "play_latency.cuh"
#include "play_quad.cuh"
#include "play_population.cuh"
#include "play_lookahead2.cuh"
#include "play_wiggle.cuh"
#include "jit.inc"

namespace tb {

// K3: all 45 features of one placement per thread
// at_row (optional): row of the piece's top bbox row; 255 / null => hard drop.  With an explicit row the placement only
// has to fit (State.future does not check support either), which covers the overhang-slide placements.
__global__ void eval_features_kernel(int N, const uint16_t* prev_boards, const uint8_t* piece, const uint8_t* rot,
                                     const uint8_t* col, const uint8_t* at_row, double* out_features, int32_t* out_info, uint16_t* out_rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint16_t rows[ROWS];
    for (int r = 0; r < ROWS; ++r) rows[r] = prev_boards[(size_t)i * ROWS + r];
    uint32_t c[COLS];
    rows_to_cols(rows, c);
    const PrevStat pv = prev_stat(c);
    const Heights H = pack_gh(pv.gh);
    const int p = piece[i];
    int legal = 0, row = -1, k = 0;
    Feat f;
    Placement P;
    SlotDesc dsel;
    bool found = false;
    if (p < 7) {
        const int ns = c_tables.nslots[p];
        for (int s = 0; s < ns; ++s) {
            const SlotDesc d = c_tables.slot[p][s];
            if ((int)((d.w[0] >> 4) & 3u) == (int)rot[i] && (int)(d.w[0] & 15u) == (int)col[i]) {
                found = true;
                dsel = d;
                if (at_row && at_row[i] != 255) {
                    const int ph = (int)((d.w[0] >> 9) & 7u), y = ROWS - (int)at_row[i] - ph;
                    if (fits_at(c, d, y)) {
                        uint32_t full;
                        place_at(c, d, y, P, &full);
                        if (full != 0u) place_clear(d, P, full);
                        legal = 1; row = (int)at_row[i]; k = P.k;
                    }
                } else if (place_slot(c, H, d, P)) { legal = 1; row = ROWS - P.y - P.ph; k = P.k; }
                break;
            }
        }
    }
    if (legal) {
        eval_features<0>(P, dsel, pv, ALL45_MASK, f);
        for (int j = 0; j < NFEAT; ++j) out_features[(size_t)i * NFEAT + j] = feat_value(f, j);
        if (out_rows) { cols_to_rows(P.n, rows); for (int r = 0; r < ROWS; ++r) out_rows[(size_t)i * ROWS + r] = rows[r]; }
    } else {
        for (int j = 0; j < NFEAT; ++j) out_features[(size_t)i * NFEAT + j] = 0.0;
        if (out_rows) for (int r = 0; r < ROWS; ++r) out_rows[(size_t)i * ROWS + r] = rows[r];
    }
    out_info[(size_t)i * 3 + 0] = found ? legal : -1;
    out_info[(size_t)i * 3 + 1] = row;
    out_info[(size_t)i * 3 + 2] = k;
}

// K4
__global__ void hashed_pieces_kernel(uint64_t seed, uint64_t first_stream, int n_streams, uint32_t T, uint8_t* out)
{
    const size_t total = (size_t)n_streams * T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / T, t = i % T;
        out[i] = (uint8_t)hashed_piece(seed, first_stream + s, t);
    }
}

// fitness[c] = sum(lines of its G games) / G   (one exact int sum, one fp64 division)
__global__ void fitness_kernel(int C, int G, const tb200_game_result* res, double* fitness)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long s = 0;
    for (int g = 0; g < G; ++g) s += res[(size_t)c * G + g].lines;
    fitness[c] = __ddiv_rn((double)s, (double)G);
}

// K1 prologue: sequence of board i = {piece[i], next[i] or 255}
__global__ void pair_pieces_kernel(int N, const uint8_t* piece, const uint8_t* next, uint8_t* seq)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    seq[2 * (size_t)i] = piece[i];
    seq[2 * (size_t)i + 1] = next ? next[i] : (uint8_t)255;
}

// K1 epilogue: repack per-game outputs of a max_moves = 1 play into the best_move layout
__global__ void best_move_pack_kernel(int N, const tb200_game_result* res, const uint32_t* trace, const double* strace,
                                      uint8_t* out_move, double* out_score, uint32_t* out_count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const bool moved = res[i].pieces > 0;
    const uint32_t rec = moved ? trace[(size_t)i * 2] : 0x0000ff00u;
    reinterpret_cast<uint32_t*>(out_move)[i] = rec;
    if (out_score) out_score[i] = moved ? strace[(size_t)i * 2] : 0.0;
    if (out_count) out_count[i] = (uint32_t)res[i].placements;
}

typedef void (*play_fn)(const PlayParams);
template <int MODE, int LOOKAHEAD, int SLIDE>
static play_fn pick_lanes(int lanes)
{
    switch (lanes) {
    case 8:  return play_games_kernel<8, MODE, LOOKAHEAD, SLIDE>;
    case 16: return play_games_kernel<16, MODE, LOOKAHEAD, SLIDE>;
    default: return play_games_kernel<32, MODE, LOOKAHEAD, SLIDE>;
    }
}
template <int SLIDE>
static play_fn pick_generic(int mode, int lookahead, int lanes)
{
    if (lookahead == 2) {
        if (mode == MODE_W6) return pick_lanes<MODE_W6, 2, SLIDE>(lanes);
        if (mode == MODE_ALL45) return pick_lanes<MODE_ALL45, 2, SLIDE>(lanes);
        return pick_lanes<MODE_GENERIC, 2, SLIDE>(lanes);
    }
    if (mode == MODE_W6) return pick_lanes<MODE_W6, 1, SLIDE>(lanes);
    if (mode == MODE_ALL45) return pick_lanes<MODE_ALL45, 1, SLIDE>(lanes);
    return pick_lanes<MODE_GENERIC, 1, SLIDE>(lanes);
}
static play_fn pick_kernel(int mode, int lookahead, int lanes, int move_set)
{
    if (move_set & (TB200_MOVES_WIGGLE | TB200_MOVES_WIGGLE_DROP)) {
        if (lookahead == 2) {
            if (mode == MODE_W6) return play_games_wiggle_kernel<MODE_W6, 2>;
            if (mode == MODE_ALL45) return play_games_wiggle_kernel<MODE_ALL45, 2>;
            return play_games_wiggle_kernel<MODE_GENERIC, 2>;
        }
        if (mode == MODE_W6) return play_games_wiggle_kernel<MODE_W6, 1>;
        if (mode == MODE_ALL45) return play_games_wiggle_kernel<MODE_ALL45, 1>;
        return play_games_wiggle_kernel<MODE_GENERIC, 1>;
    }
    if (move_set != 0) return pick_generic<1>(mode, lookahead, lanes);     // overhang slides / pressure: extended generic kernels
    if (lookahead == 2 && lanes == 32) {                   // leaves flattened over the lanes
        if (mode == MODE_W6) return play_games_l2flat_kernel<MODE_W6>;
        if (mode == MODE_ALL45) return play_games_l2flat_kernel<MODE_ALL45>;
        return play_games_l2flat_kernel<MODE_GENERIC>;
    }
    if (lanes == 64) return play_games_w6pair_kernel;      // caller guarantees lookahead 1 and the w6 set
    if (lookahead == 1 && mode == MODE_W6 && lanes == 16) return play_games_w6grp_kernel<16>;
    if (lookahead == 1 && mode == MODE_W6 && lanes == 8) return play_games_w6grp_kernel<8>;
    if (lookahead == 1 && mode == MODE_W6 && lanes == 4) return play_games_w6grp_kernel<4>;
    if (lookahead == 1 && mode == MODE_W6 && lanes == 2) return play_games_w6grp_kernel<2>;
    if (lookahead == 1 && mode == MODE_W6 && lanes == 1) return play_games_w6grp_kernel<1>;
    if (lookahead == 1 && lanes == 32) {
        if (mode == MODE_W6) return play_games_w6fast_kernel<false>;
        if (mode == MODE_ALL45) return play_games_w32_kernel<MODE_ALL45>;
        return play_games_w32_kernel<MODE_GENERIC>;
    }
    return pick_generic<0>(mode, lookahead, lanes);
}

}  // namespace tb

// =================================================================================================
// C ABI
// =================================================================================================
struct tb200_ctx {
    int device = 0;
    int sms = 0;
    int clock_khz = 0;
    char name[128] = {0};
    std::string err;
    int F = 0;
    int mode = tb::MODE_GENERIC;
    uint64_t rmask = 0;
    uint8_t ids[TB200_MAX_WEIGHTED] = {0};
    uint32_t* d_ticket = nullptr;
    // workspace of the re-spreading stages (launch_play): control words, two hand-over lists, saved game states
    char* d_stage = nullptr;
    uint32_t stage_cap = 0;              // games the workspace holds
    bool staging_enabled = true;
    // workspace of the lockstep kernel K2s (grown on demand): saved games, two list buffers, bucket control words
    char* d_lock = nullptr;
    size_t lock_cap = 0;
    bool lockstep_enabled = true;
    // run-time specialised kernels for the current (generic) feature list: see jit.inc
    tbjit::Module* jit = nullptr;
    bool jit_failed = false;             // compilation / loading failed for the current list: do not retry on every call
    int jit_policy = -1;                 // TB200_JIT: 0 never, 1 always, unset: plays of >= 2^22 game-moves
    bool win_enabled = true;             // K2sW: window-local evaluation in the one-lane-per-game lockstep kernel (TB200_NO_WIN: the all-columns K2s)
    int lock_lanes = 0, lock_K = 64;     // tuning aids: TB200_LOCK_LANES, TB200_LOCK_K
    int lock_thr[3] = {24, 8, 2};        // K2sA width thresholds (games alive per resident warp): TB200_LOCK_THR=a,b,c
    int lock_steady = 10;                // K2sA: one-lane threshold of a steady bucket (games alive per resident warp; 0: same as lock_thr[0]): TB200_LOCK_STEADY
    int lock_pair = 0;                   // K2sP (opt-in, measured and not adopted — DESIGN.md §4): games alive per resident warp from which a bucket is played by
                                         // pairs of warps (0: never): TB200_LOCK_PAIR
    int lock_minb = 4;                   // K2sA resident blocks per SM (4: 128 registers — the 4- and 8-lane tasks lose 20 % at 96; 5: 96 registers, 25 % more
                                         // resident warps): TB200_LOCK_MINB
    bool bin_enabled = true;             // K2b; TB200_NO_BINS, TB200_BIN_LANES, TB200_BIN_MIN_GAMES
    int bin_lanes = 0; long long bin_min_games = 32768;
    const uint32_t* lock_flag = nullptr; // device word set by K2s if a bucket wait ever timed out (checked after host-mode calls)
    int stage_cut1 = 0, stage_cut2 = 0, stage_cut3 = 0;  // hand-over thresholds (games still unfinished): 4, 2 and 1 per SM unless TB200_STAGE_CUTS=a,b[,c]
    bool quad_enabled = false;           // K2q as the sparse stages: measured SLOWER than K2p (1450 vs 1170 cycles per piece), so off unless TB200_QUAD=1
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // one call in flight per ctx: `busy` is recorded after the last kernel of every call; the next call makes its stream wait for it
    cudaEvent_t busy = nullptr;
    bool busy_recorded = false;
    bool seq_major = true;               // TB200_NO_SEQ_MAJOR clears it (read once at tb200_create)
    bool timing_valid = false;
    bool kernel_events_valid = true;     // false after a graph replay: ev[0] / ev[1] were not recorded by it
    tb200_timing timing{};
    // staging arena for TB200_MEM_HOST calls (grown on demand, reused across calls)
    char* d_arena = nullptr;
    size_t arena_cap = 0;
    // CUDA-graph cache of the last host-mode play_games call: when the same call (same pinned host buffers,
    // sizes and options) is repeated — one CE generation after another — the whole H2D / memset / kernels /
    // D2H sequence is replayed with one cudaGraphLaunch instead of ~12 runtime calls.
    bool graphs_enabled = true;
    bool graph_valid = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    tb200_play_args graph_args{};
    int graph_F = 0, graph_mode = 0;
    uint8_t graph_ids[TB200_MAX_WEIGHTED] = {0};
    char* graph_arena = nullptr;
    tb200_timing graph_timing{};
    const uint32_t* graph_lock_flag = nullptr;
};

static int fail(tb200_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    return code;
}
#define TB_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(ctx, TB200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static void drop_graph(tb200_ctx* ctx)
{
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->graph) cudaGraphDestroy(ctx->graph);
    ctx->graph_exec = nullptr; ctx->graph = nullptr; ctx->graph_valid = false;
}

static bool is_pinned(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// One call in flight per ctx (include/tetris_b200.h): the stream of a new call first waits for the last kernel of the
// previous call on this ctx, whatever stream that ran on, because all calls share the ctx workspace.
static int enter_call(tb200_ctx* ctx, cudaStream_t st)
{
    if (ctx->busy_recorded) TB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->busy, 0));
    return TB200_OK;
}
static int leave_call(tb200_ctx* ctx, cudaStream_t st)
{
    TB_CUDA(ctx, cudaEventRecord(ctx->busy, st));
    ctx->busy_recorded = true;
    return TB200_OK;
}
// before any workspace buffer is freed or reallocated: nothing enqueued on this ctx may still use it
static void quiesce(tb200_ctx* ctx)
{
    if (ctx->busy_recorded && cudaEventSynchronize(ctx->busy) != cudaSuccess) cudaGetLastError();
}

static int ensure_arena(tb200_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->arena_cap) return TB200_OK;
    quiesce(ctx);
    drop_graph(ctx);
    if (ctx->d_arena) { TB_CUDA(ctx, cudaFree(ctx->d_arena)); ctx->d_arena = nullptr; ctx->arena_cap = 0; }
    const size_t cap = bytes + bytes / 4 + (1u << 20);
    TB_CUDA(ctx, cudaMalloc(&ctx->d_arena, cap));
    ctx->arena_cap = cap;
    return TB200_OK;
}

extern "C" {

const char* tb200_version(void) { return "tetris_b200 0.1 (sm_100a)"; }
#ifdef TB_PHASES
int tb200_debug_phases(unsigned long long* out8, int reset)
{
    unsigned long long z[8] = {0,0,0,0,0,0,0,0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, tb::g_phase, sizeof z);
    if (reset) cudaMemcpyToSymbol(tb::g_phase, z, sizeof z);
    return 0;
}
#endif

int tb200_create(int device, tb200_ctx** out)
{
    if (!out) return TB200_E_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return TB200_E_NO_DEVICE;
    tb200_ctx* ctx = new (std::nothrow) tb200_ctx();
    if (!ctx) return TB200_E_ALLOC;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return TB200_E_NO_DEVICE; }
    ctx->sms = prop.multiProcessorCount;
    cudaDeviceGetAttribute(&ctx->clock_khz, cudaDevAttrClockRate, device);
    strncpy(ctx->name, prop.name, sizeof(ctx->name) - 1);
    static const tb::Tables host_tab = { TB_SLOT_TABLE_INIT, TB_NSLOTS_INIT };
    cudaError_t e = cudaMemcpyToSymbol(tb::c_tables, &host_tab, sizeof(tb::Tables));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(tb::g_tables, &host_tab, sizeof(tb::Tables));
    if (e == cudaSuccess) {
        static tb::WinDesc host_win[7][TB_MAX_SLOTS];
        for (int p = 0; p < 7; ++p) for (int s2 = 0; s2 < TB_MAX_SLOTS; ++s2) host_win[p][s2] = tb::win_desc(host_tab.slot[p][s2]);
        e = cudaMemcpyToSymbol(tb::c_win, host_win, sizeof host_win);
    }
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_ticket, sizeof(uint32_t));
    ctx->stage_cap = (uint32_t)(16 * ctx->sms);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_stage, 256 + (size_t)ctx->stage_cap * (3 * sizeof(uint32_t) + sizeof(tb::GameSave)));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->busy, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        fprintf(stderr, "tb200_create: %s\n", cudaGetErrorString(e));
        tb200_destroy(ctx);
        return TB200_E_CUDA;
    }
    if (getenv("TB200_NO_GRAPHS")) ctx->graphs_enabled = false;
    if (getenv("TB200_NO_STAGES")) ctx->staging_enabled = false;
    if (getenv("TB200_NO_LOCKSTEP")) ctx->lockstep_enabled = false;
    if (getenv("TB200_NO_WIN")) ctx->win_enabled = false;
    if (const char* ej = getenv("TB200_JIT")) ctx->jit_policy = atoi(ej) != 0 ? 1 : 0;
    if (getenv("TB200_NO_BINS")) ctx->bin_enabled = false;
    if (getenv("TB200_NO_SEQ_MAJOR")) ctx->seq_major = false;
    if (const char* e5 = getenv("TB200_BIN_LANES")) ctx->bin_lanes = atoi(e5);
    if (const char* e6 = getenv("TB200_BIN_MIN_GAMES")) { const long long v = atoll(e6); if (v >= 1024) ctx->bin_min_games = v; }
    if (const char* e3 = getenv("TB200_LOCK_LANES")) ctx->lock_lanes = atoi(e3);
    if (const char* e7 = getenv("TB200_LOCK_THR")) { int a = 0, b = 0, c2 = 0; if (sscanf(e7, "%d,%d,%d", &a, &b, &c2) == 3 && a >= b && b >= c2 && c2 >= 0) { ctx->lock_thr[0] = a; ctx->lock_thr[1] = b; ctx->lock_thr[2] = c2; } }
    if (const char* e10 = getenv("TB200_LOCK_STEADY")) { const int v = atoi(e10); if (v >= 0) ctx->lock_steady = v; }
    if (const char* e9 = getenv("TB200_LOCK_PAIR")) { const int v = atoi(e9); if (v >= 0) ctx->lock_pair = v; }
    if (const char* e8 = getenv("TB200_LOCK_MINB")) { const int v = atoi(e8); if (v == 4 || v == 5) ctx->lock_minb = v; }
    if (const char* e4 = getenv("TB200_LOCK_K")) { const int k = atoi(e4); if (k >= 8) ctx->lock_K = k; }
    ctx->stage_cut1 = 4 * ctx->sms; ctx->stage_cut2 = 2 * ctx->sms; ctx->stage_cut3 = ctx->sms;
    if (getenv("TB200_QUAD")) ctx->quad_enabled = true;
    if (const char* e2 = getenv("TB200_STAGE_CUTS")) {          // tuning aid (tools/stage_probe.py)
        int c1 = 0, c2 = 0, c3 = -1;
        const int nr = sscanf(e2, "%d,%d,%d", &c1, &c2, &c3);
        if (nr >= 2 && c1 >= c2 && c2 >= 1 && c1 <= 4 * ctx->sms) {
            ctx->stage_cut1 = c1; ctx->stage_cut2 = c2;
            if (nr == 3 && c3 >= 0 && c3 <= c2) ctx->stage_cut3 = c3; else if (ctx->stage_cut3 > c2) ctx->stage_cut3 = c2;
        }
    }
    *out = ctx;
    return TB200_OK;
}

int tb200_alloc_pinned(size_t bytes, void** out)
{
    if (!out || bytes == 0) return TB200_E_BAD_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? TB200_OK : TB200_E_ALLOC;
}

void tb200_free_pinned(void* p) { if (p) cudaFreeHost(p); }

void tb200_destroy(tb200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    drop_graph(ctx);
    tbjit::unload(ctx->jit); ctx->jit = nullptr;
    if (ctx->d_ticket) cudaFree(ctx->d_ticket);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->d_lock) cudaFree(ctx->d_lock);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->busy) cudaEventDestroy(ctx->busy);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* tb200_last_error(const tb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int tb200_device_info(tb200_ctx* ctx, int32_t* sms, int32_t* clock_khz, char* name, int32_t name_len)
{
    if (!ctx) return TB200_E_BAD_ARG;
    if (sms) *sms = ctx->sms;
    if (clock_khz) *clock_khz = ctx->clock_khz;
    if (name && name_len > 0) { strncpy(name, ctx->name, (size_t)name_len - 1); name[name_len - 1] = 0; }
    return TB200_OK;
}

int tb200_set_features(tb200_ctx* ctx, const int32_t* feature_ids, int32_t F)
{
    if (!ctx) return TB200_E_BAD_ARG;
    if (!feature_ids || F < 1 || F > TB200_MAX_WEIGHTED) return fail(ctx, TB200_E_BAD_ARG, "set_features: F must be 1..48");
    uint64_t mask = 0;
    for (int i = 0; i < F; ++i) {
        if (feature_ids[i] < 0 || feature_ids[i] >= TB200_NUM_FEATURES) return fail(ctx, TB200_E_BAD_ARG, "set_features: feature id out of range");
        ctx->ids[i] = (uint8_t)feature_ids[i];
        mask |= 1ull << feature_ids[i];
    }
    if (ctx->jit && (ctx->jit->F != F || memcmp(ctx->jit->ids, ctx->ids, (size_t)F) != 0)) {      // another list: the specialised kernels go
        quiesce(ctx); drop_graph(ctx);
        tbjit::unload(ctx->jit); ctx->jit = nullptr;
    }
    if (ctx->F != F || ctx->rmask != mask) ctx->jit_failed = false;
    ctx->F = F; ctx->rmask = mask;
    static const int w6[6] = {tb::F_LANDING_HEIGHT, tb::F_ERODED_CELLS, tb::F_ROW_TRANS, tb::F_COL_TRANS, tb::F_PITS, tb::F_CUML_WELLS};
    bool is_w6 = F == 6, is_all = F == TB200_NUM_FEATURES;
    for (int i = 0; i < F && is_w6; ++i) is_w6 = feature_ids[i] == w6[i];
    for (int i = 0; i < F && is_all; ++i) is_all = feature_ids[i] == i;
    ctx->mode = is_w6 ? tb::MODE_W6 : (is_all ? tb::MODE_ALL45 : tb::MODE_GENERIC);
    return TB200_OK;
}

int tb200_last_timing(tb200_ctx* ctx, tb200_timing* out)
{
    if (!ctx || !out) return TB200_E_BAD_ARG;
    if (!ctx->timing_valid) return fail(ctx, TB200_E_BAD_ARG, "last_timing: no play_games call yet");
    ctx->timing.kernel_ms = 0.f;
    if (ctx->kernel_events_valid) {
        float ms = 0.f;
        if (cudaEventSynchronize(ctx->ev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->timing.kernel_ms = ms;
        else cudaGetLastError();         // never leave a stale error behind for the next launch check
    }
    *out = ctx->timing;
    return TB200_OK;
}

// ---- lockstep kernel K2s: when it applies, and its workspace (allocated outside any stream capture) ----
struct LockPlan { bool on; int kind; uint32_t K, S; size_t b_save, b_list, b_ctl; uint32_t ring_cap; };   // kind 1: K2s, 2: K2b
static LockPlan lock_plan(const tb200_ctx* ctx, const tb200_play_args& a, int per_game_weights)
{
    LockPlan lp{};
    const long long n_games = (long long)a.n_candidates * a.games_per_candidate;
    const int l = a.lanes;
    const bool lanes_auto = !(l == 1 || l == 2 || l == 4 || l == 8 || l == 16 || l == 32 || l == 64 || l == 128);
    const uint32_t stop_moves = a.max_moves && a.max_moves < a.T ? a.max_moves : a.T;
    // hard drops; inside move_with_pressure too, for the w6 set with the per-bucket-width kernel (the only one instantiated with the filter)
    const bool press_ok = a.move_set == TB200_MOVES_PRESSURE && ctx->mode == tb::MODE_W6 && ctx->win_enabled &&
                          !(ctx->lock_lanes == 1 || ctx->lock_lanes == 2 || ctx->lock_lanes == 4 || ctx->lock_lanes == 8);
    lp.on = ctx->lockstep_enabled && lanes_auto && a.lookahead == 1 && (a.move_set == 0 || press_ok) && per_game_weights == 0 &&
           
            !a.seq_of_game && a.games_per_candidate >= 4 && a.games_per_candidate <= 4096 &&
            (n_games + ctx->sms - 1) / ctx->sms > 80 && stop_moves >= (uint32_t)(2 * ctx->lock_K);
    if (!lp.on) {
        // independent piece streams (or too few games per sequence to stay in step): the piece-binned kernel K2b, once the
        // population is several times what the machine holds in flight
        const bool pb = ctx->bin_enabled && lanes_auto && ctx->mode == tb::MODE_W6 && a.lookahead == 1 && a.move_set == 0 && per_game_weights == 0 &&
                        n_games >= (long long)ctx->bin_min_games && n_games <= 0x40000000LL;
        if (!pb) return lp;
        lp.on = true; lp.kind = 2;
        uint32_t cap = 256;
        while ((long long)cap * tb::PB_SHARDS < n_games + tb::PB_SHARDS) cap <<= 1;           // >= games per shard
        lp.ring_cap = cap;
        lp.b_save = align_up(sizeof(tb::GameSave) * (size_t)n_games);
        lp.b_list = align_up(7 * sizeof(uint32_t) * (size_t)cap * tb::PB_SHARDS);
        lp.b_ctl = align_up(8 * (16 * tb::PB_SHARDS + 16));
        return lp;
    }
    lp.kind = 1;
    lp.K = (uint32_t)ctx->lock_K; lp.S = (stop_moves + lp.K - 1) / lp.K + 2;      // see lock_stage_begin
    lp.b_save = align_up(sizeof(tb::GameSave) * (size_t)n_games);
    lp.b_list = align_up(2 * sizeof(uint32_t) * (size_t)n_games);
    lp.b_ctl = align_up(16 * (size_t)(lp.S + 1) * (size_t)a.games_per_candidate + 256 * 4 * sizeof(uint32_t));      // + K2sA: warps seen per (SM, scheduler)
    return lp;
}
static int ensure_lock_ws(tb200_ctx* ctx, const tb200_play_args& a)
{
    const LockPlan lp = lock_plan(ctx, a, 0);
    if (!lp.on || lp.b_save + lp.b_list + lp.b_ctl <= ctx->lock_cap) return TB200_OK;
    quiesce(ctx);
    drop_graph(ctx);
    if (ctx->d_lock) { TB_CUDA(ctx, cudaFree(ctx->d_lock)); ctx->d_lock = nullptr; ctx->lock_cap = 0; }
    const size_t cap = (lp.b_save + lp.b_list + lp.b_ctl) * 5 / 4;
    TB_CUDA(ctx, cudaMalloc(&ctx->d_lock, cap));
    ctx->lock_cap = cap;
    return TB200_OK;
}

// ---- run-time specialised kernels (jit.inc): which plays use them, and their one-off compilation (outside any stream capture) ----
static bool jit_applies(const tb200_ctx* ctx, const tb200_play_args& a)
{
    if (ctx->mode != tb::MODE_GENERIC || a.lookahead != 1 || a.move_set != 0 || ctx->jit_policy == 0) return false;
    const uint32_t moves = a.max_moves && a.max_moves < a.T ? a.max_moves : a.T;
    const unsigned long long work = (unsigned long long)a.n_candidates * a.games_per_candidate * moves;
    return ctx->jit_policy == 1 || work >= (1ull << 22);
}
static void ensure_jit(tb200_ctx* ctx, const tb200_play_args& a)
{
    if (!jit_applies(ctx, a) || ctx->jit || ctx->jit_failed) return;
    static const tb::Tables host_tab = { TB_SLOT_TABLE_INIT, TB_NSLOTS_INIT };
    static tb::WinDesc host_win[7][TB_MAX_SLOTS];
    for (int p = 0; p < 7; ++p) for (int s2 = 0; s2 < TB_MAX_SLOTS; ++s2) host_win[p][s2] = tb::win_desc(host_tab.slot[p][s2]);
    std::string why;
    quiesce(ctx);
    drop_graph(ctx);
    ctx->jit = tbjit::load_module(ctx->ids, ctx->F, host_tab, &host_win[0][0], sizeof host_win, why);
    if (!ctx->jit) {
        ctx->jit_failed = true;
        ctx->err = "run-time specialisation unavailable (" + why + "): the precompiled generic kernels run instead";
        fprintf(stderr, "tetris_b200: %s\n", ctx->err.c_str());
    }
}
// launch a kernel of the library (fn) or its run-time specialised twin (cf != null)
static int launch_any(tb200_ctx* ctx, tb::play_fn fn, CUfunction cf, unsigned grid, unsigned block, size_t smem, cudaStream_t st, tb::PlayParams& p)
{
    if (cf) {
        void* params[] = { &p };
        if (tbjit::g_api.LaunchKernel(cf, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)st, params, nullptr) != CUDA_SUCCESS)
            return fail(ctx, TB200_E_CUDA, "cuLaunchKernel of a specialised kernel failed");
        return TB200_OK;
    }
    fn<<<grid, block, smem, st>>>(p);
    TB_CUDA(ctx, cudaGetLastError());
    return TB200_OK;
}

// launch K2 (+ optional fitness) on device buffers
static int launch_play(tb200_ctx* ctx, const tb200_play_args& a, cudaStream_t st, int per_game_weights, int* launches, bool record_events = true)
{
    const long long n_games_ll = (long long)a.n_candidates * a.games_per_candidate;
    const uint32_t n_games = (uint32_t)n_games_ll;
    int lanes = a.lanes;
    const int move_set = a.move_set;
    const bool pair_ok = ctx->mode == tb::MODE_W6 && a.lookahead == 1 && move_set == 0;
    const bool lanes_auto = !(lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32 || lanes == 64 || lanes == 128);
    if (lanes == 128 && !pair_ok) lanes = 32;
    if (move_set != 0 && (lanes == 64 || lanes == 4 || lanes == 2 || lanes == 1)) lanes = lanes == 64 ? 32 : 8;
    if (move_set & (TB200_MOVES_WIGGLE | TB200_MOVES_WIGGLE_DROP)) lanes = 32;
    if (lanes == 64 && !pair_ok) lanes = 32;
    const bool narrow_ok = pair_ok && (lanes == 4 || lanes == 2 || lanes == 1);
    if (!narrow_ok && lanes != 8 && lanes != 16 && lanes != 32 && lanes != 64 && lanes != 128) {
        // auto (measured on B200, tools/lanes_sweep.py): two warps per game only while that still leaves warp
        // schedulers idle; one warp per game while the population is a few resident waves (lowest latency per
        // piece); narrower groups — more games per warp, better lane utilisation — as the population grows.
        const long long per_sm = (n_games_ll + ctx->sms - 1) / ctx->sms;
        if (a.lookahead == 2 && move_set == 0) lanes = 32;        // K2L: leaves flattened over a full warp
        else if (pair_ok && ctx->quad_enabled && per_sm <= 2) lanes = 128;      // K2q: four warps per game while every warp can still have a scheduler (almost) to itself
        else if (pair_ok && per_sm <= 4) lanes = 64;
        else if (per_sm <= 80) lanes = 32;
        else if (pair_ok) lanes = per_sm <= 320 ? 8 : (per_sm <= 700 ? 4 : 2);
        else lanes = per_sm <= 160 ? 16 : 8;
    }
    tb::play_fn fn = lanes == 128 ? tb::play_games_w6quad_kernel : tb::pick_kernel(ctx->mode, a.lookahead, lanes, move_set);
    const int block = lanes == 64 ? 64 : tb::BLOCK;
    const size_t dyn_smem = lanes == 128 ? sizeof(tb::QShared) : 0;
    int occ = 0;
    TB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, block, dyn_smem));
    if (occ < 1) occ = 1;
    const int groups_per_block = lanes >= 64 ? 1 : tb::BLOCK / lanes;
    long long blocks = (n_games_ll + groups_per_block - 1) / groups_per_block;
    const long long max_blocks = (long long)ctx->sms * occ;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;

    tb::PlayParams p;
    memset(&p, 0, sizeof p);
    p.weights = a.weights; p.pieces = a.pieces; p.seq_of_game = a.seq_of_game; p.init_rows = a.init_rows; p.init_lines = a.init_lines;
    p.results = a.results; p.trace = a.trace; p.score_trace = a.score_trace; p.final_rows = a.final_rows;
    p.ticket = ctx->d_ticket; p.seed = a.seed; p.rmask = ctx->rmask; p.n_games = n_games; p.T = a.T;
    p.max_moves = a.max_moves ? a.max_moves : a.T; p.G = a.games_per_candidate; p.F = ctx->F;
    p.per_game_weights = per_game_weights;
    if (per_game_weights == 0 && !a.seq_of_game && a.games_per_candidate > 1 && ctx->seq_major) p.seq_major_C = (uint32_t)a.n_candidates;
    p.move_set = move_set;
    p.press.on = (move_set & TB200_MOVES_PRESSURE) ? 1 : 0;
    p.press.two_but_rot = a.pressure_two_but_rot ? 1 : 0;
    p.press.avg_lat = a.pressure_avg_lat; p.press.resp_time = a.pressure_resp_time; p.press.move_eff = a.pressure_move_eff;
    memcpy(p.ids, ctx->ids, sizeof p.ids);
    {
        static const uint32_t default_pts[5] = {0u, 40u, 100u, 300u, 1200u};          // game.py:159
        const bool custom = (a.points[0] | a.points[1] | a.points[2] | a.points[3] | a.points[4]) != 0u;
        memcpy(p.pts, custom ? a.points : default_pts, sizeof p.pts);
    }

    // Latency regime (between 2 and 16 games per SM, w6, lookahead 1, automatic lanes): the kernel time is the serial chain
    // of the longest games, and a game's per-piece latency depends on how many other live games share its warp scheduler
    // (measured, tools/latency_probe.py: 1880 cycles per piece with two K2f warps per scheduler, 1495 alone, 1225 as a K2p
    // pair with a scheduler per warp).  Games die at very different ages, so the survivors end up unevenly spread.  The
    // play is therefore split into up to three stream-ordered stages: K2f on everything until at most 4 games per SM are
    // unfinished, K2f again on the survivors — one block per SM, one warp per scheduler — until at most 2 per SM are
    // left, then K2p with one scheduler per warp.  A stage ends by itself: every warp polls the finished-games counter
    // and hands its game (board + counters, 96 bytes) to the next stage's list.  Results do not depend on the staging.
    const uint32_t stop_moves = p.max_moves < p.T ? p.max_moves : p.T;
    const bool staged = ctx->staging_enabled && lanes_auto && pair_ok && n_games > (uint32_t)(2 * ctx->sms) && n_games <= ctx->stage_cap && stop_moves >= 128u;
    // Population regime (more than 80 games per SM, w6, lookahead 1, G shared sequences, automatic lanes): the lockstep
    // kernel K2s keeps every warp on one sequence at one move index.
    const long long per_sm_games = (n_games_ll + ctx->sms - 1) / ctx->sms;
    const LockPlan lp = lock_plan(ctx, a, per_game_weights);
    const bool lockstep = lp.on && lp.b_save + lp.b_list + lp.b_ctl <= ctx->lock_cap;      // workspace comes from ensure_lock_ws()
    bool used_jit = false;
    TB_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(uint32_t), st));
    TB_CUDA(ctx, cudaMemsetAsync(a.results, 0, sizeof(tb200_game_result) * (size_t)n_games, st));
    if (lockstep && lp.kind == 2) {
        p.st_save = reinterpret_cast<tb::GameSave*>(ctx->d_lock);
        p.pb_ring = reinterpret_cast<uint32_t*>(ctx->d_lock + lp.b_save);
        p.pb_ctl = reinterpret_cast<unsigned long long*>(ctx->d_lock + lp.b_save + lp.b_list);
        p.pb_mask = lp.ring_cap - 1u;
        ctx->lock_flag = reinterpret_cast<const uint32_t*>(p.pb_ctl + 16 * tb::PB_SHARDS + 1);
        int bl = ctx->bin_lanes;
        if (bl != 1 && bl != 2 && bl != 4 && bl != 8) bl = per_sm_games <= 330 ? 4 : (per_sm_games <= 1300 ? 2 : 1);   // tools/bin_probe.py
        tb::play_fn bf = bl == 8 ? tb::play_games_w6bin_kernel<8> : (bl == 4 ? tb::play_games_w6bin_kernel<4> : (bl == 2 ? tb::play_games_w6bin_kernel<2> : tb::play_games_w6bin_kernel<1>));
        int bocc = 0;
        TB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bocc, bf, tb::BLOCK, 0));
        if (bocc < 1) bocc = 1;
        const long long bblocks = (long long)ctx->sms * bocc;
        TB_CUDA(ctx, cudaMemsetAsync(p.pb_ring, 0xff, lp.b_list, st));
        TB_CUDA(ctx, cudaMemsetAsync(p.pb_ctl, 0, lp.b_ctl, st));
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
        tb::pb_init_kernel<<<(unsigned)((n_games + 255u) / 256u), 256, 0, st>>>(p);
        TB_CUDA(ctx, cudaGetLastError());
        bf<<<(unsigned)bblocks, tb::BLOCK, 0, st>>>(p);
        TB_CUDA(ctx, cudaGetLastError());
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
        *launches += 2;
        lanes = bl; blocks = bblocks;
    } else if (lockstep) {
        const uint32_t K = lp.K, S = lp.S;
        const size_t b_save = lp.b_save, b_list = lp.b_list, b_ctl = lp.b_ctl;
        p.st_save = reinterpret_cast<tb::GameSave*>(ctx->d_lock);
        p.lk_list = reinterpret_cast<uint32_t*>(ctx->d_lock + b_save);
        p.lk_ctl = reinterpret_cast<uint32_t*>(ctx->d_lock + b_save + b_list);
        p.lk_C = (uint32_t)a.n_candidates; p.lk_K = K; p.lk_S = S;
        p.lk_thr[0] = (uint32_t)ctx->lock_thr[0]; p.lk_thr[1] = (uint32_t)ctx->lock_thr[1]; p.lk_thr[2] = (uint32_t)ctx->lock_thr[2];
        {   // pairs of warps take the range [lock_pair, lock_thr[0]) of games alive per resident warp, when that range is not empty
            int pr = ctx->lock_pair > 0 ? ctx->lock_pair : ctx->lock_thr[0];
            if (pr > ctx->lock_thr[0]) pr = ctx->lock_thr[0];
            if (pr < ctx->lock_thr[1]) pr = ctx->lock_thr[1];
            p.lk_thr[3] = (uint32_t)pr;
        }
        // The pair task can only matter for plays that do not fill the machine with one warp per 32 games from the start (a rank's slice of a
        // sharded population); only they get the kernel variant that has it (it costs the plain one-lane path 4 %).
        const unsigned long long warps4 = (unsigned long long)ctx->sms * 4ull * (tb::BLOCK / 32);
        p.lk_thr[4] = ctx->lock_steady > 0 && ctx->lock_steady < ctx->lock_thr[0] ? (uint32_t)ctx->lock_steady : p.lk_thr[0];
        const bool pair_on = p.lk_thr[3] < p.lk_thr[0] && (unsigned long long)n_games_ll < (unsigned long long)p.lk_thr[0] * warps4;
        if (!pair_on) p.lk_thr[3] = p.lk_thr[0];
        const int minb = ctx->lock_minb;
        ctx->lock_flag = p.lk_ctl + (size_t)S * (size_t)a.games_per_candidate * 4 + 3;
        int ll = ctx->lock_lanes;
        if (ll != 1 && ll != 2 && ll != 4 && ll != 8) {
            ll = per_sm_games <= 320 ? 8 : (per_sm_games <= 760 ? 4 : (per_sm_games <= 1500 ? 2 : 1));   // tools/lock_probe.py
            if (ctx->mode != tb::MODE_W6 && ll < 2) ll = 2;     // the general evaluation is register-heavy: two lanes per game (tools/generic_pop_probe.py)
        }
        tb::play_fn lf;
        const bool lock_auto = ctx->mode == tb::MODE_W6 && ctx->win_enabled && !(ctx->lock_lanes == 1 || ctx->lock_lanes == 2 || ctx->lock_lanes == 4 || ctx->lock_lanes == 8);
        if (lock_auto)
            lf = move_set == TB200_MOVES_PRESSURE ? (pair_on ? tb::play_games_w6lockauto_kernel<4, true, true> : tb::play_games_w6lockauto_kernel<4, true>) :
                 (pair_on ? (minb == 4 ? tb::play_games_w6lockauto_kernel<4, false, true> : tb::play_games_w6lockauto_kernel<5, false, true>) :
                            (minb == 4 ? tb::play_games_w6lockauto_kernel<4> : tb::play_games_w6lockauto_kernel<5>));      // lanes per game chosen per (stage, sequence) bucket from the live count
        else if (ctx->mode == tb::MODE_W6 && ll == 1 && ctx->win_enabled)
            lf = tb::play_games_w6lockwin_kernel;
        else if (ctx->mode == tb::MODE_W6)
            lf = ll == 8 ? tb::play_games_w6lock_kernel<8> : (ll == 4 ? tb::play_games_w6lock_kernel<4> : (ll == 2 ? tb::play_games_w6lock_kernel<2> : tb::play_games_w6lock_kernel<1>));
        else if (ctx->mode == tb::MODE_ALL45)
            lf = ll == 8 ? tb::play_games_w6lock_kernel<8, tb::MODE_ALL45> : (ll == 4 ? tb::play_games_w6lock_kernel<4, tb::MODE_ALL45> :
                 (ll == 2 ? tb::play_games_w6lock_kernel<2, tb::MODE_ALL45> : tb::play_games_w6lock_kernel<1, tb::MODE_ALL45>));
        else
            lf = ll == 8 ? tb::play_games_w6lock_kernel<8, tb::MODE_GENERIC> : (ll == 4 ? tb::play_games_w6lock_kernel<4, tb::MODE_GENERIC> :
                 (ll == 2 ? tb::play_games_w6lock_kernel<2, tb::MODE_GENERIC> : tb::play_games_w6lock_kernel<1, tb::MODE_GENERIC>));
        CUfunction lcf = nullptr;
        if (ctx->jit && jit_applies(ctx, a) && ctx->mode == tb::MODE_GENERIC) {
            if (ll == 1) ll = 2;
            lcf = ctx->jit->fn[ll == 8 ? tbjit::K_LOCK8 : (ll == 4 ? tbjit::K_LOCK4 : tbjit::K_LOCK2)];
            used_jit = true;
        }
        int locc = 0;
        if (lcf) { if (tbjit::g_api.Occupancy(&locc, lcf, tb::BLOCK, 0) != CUDA_SUCCESS) locc = 1; }
        else TB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&locc, lf, tb::BLOCK, 0));
        if (locc < 1) locc = 1;
        long long lblocks = (n_games_ll * ll + tb::BLOCK - 1) / tb::BLOCK;           // the grid must be fully resident: tasks wait on tasks
        if (lblocks > (long long)ctx->sms * locc || lock_auto) lblocks = (long long)ctx->sms * locc;
        if (lock_auto) {                                        // reported width = the one the first stage plays with (tb::lock_auto_lanes)
            const unsigned long long alive = (unsigned long long)n_games_ll, warps = (unsigned long long)lblocks * (tb::BLOCK / 32);
            ll = alive >= (unsigned long long)p.lk_thr[0] * warps ? 1 : (alive >= (unsigned long long)p.lk_thr[3] * warps ? 2 : (alive >= (unsigned long long)p.lk_thr[1] * warps ? 4 : (alive >= (unsigned long long)p.lk_thr[2] * warps ? 8 : 32)));
        }
        TB_CUDA(ctx, cudaMemsetAsync(p.lk_ctl, 0, b_ctl, st));
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
        { const int rcl = launch_any(ctx, lf, lcf, (unsigned)lblocks, tb::BLOCK, 0, st, p); if (rcl) return rcl; }
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
        *launches += 1;
        lanes = ll; blocks = lblocks;
    } else if (staged) {
        uint32_t* ctl = reinterpret_cast<uint32_t*>(ctx->d_stage);
        uint32_t* list1 = reinterpret_cast<uint32_t*>(ctx->d_stage + 256);
        uint32_t* list2 = list1 + ctx->stage_cap;
        uint32_t* list3 = list2 + ctx->stage_cap;
        p.st_ctl = ctl;
        p.st_save = reinterpret_cast<tb::GameSave*>(list3 + ctx->stage_cap);
        TB_CUDA(ctx, cudaMemsetAsync(ctl, 0, 64, st));
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
        const bool dense = n_games > (uint32_t)ctx->stage_cut1;
        if (dense) {
            p.st_stage = 0; p.st_list_in = nullptr; p.st_list_out = list1; p.st_cut = (uint32_t)ctx->stage_cut1;
            tb::play_games_w6fast_kernel<true><<<(unsigned)blocks, tb::BLOCK, 0, st>>>(p);
            TB_CUDA(ctx, cudaGetLastError());
            *launches += 1;
        }
        p.st_stage = 1; p.st_list_in = dense ? list1 : nullptr; p.st_list_out = list2; p.st_cut = (uint32_t)ctx->stage_cut2;
        tb::play_games_w6fast_kernel<true><<<(unsigned)ctx->sms, tb::BLOCK, 0, st>>>(p);
        TB_CUDA(ctx, cudaGetLastError());
        if (ctx->quad_enabled) {
            // K2q, two games per SM (two warps per scheduler, each with a third of the per-piece instructions), then one game per SM
            const bool two = ctx->stage_cut3 > 0 && ctx->stage_cut3 < ctx->stage_cut2;
            if (two) {
                p.st_stage = 2; p.st_list_in = list2; p.st_list_out = list3; p.st_cut = (uint32_t)ctx->stage_cut3;
                tb::play_games_w6quad_kernel<<<(unsigned)ctx->stage_cut2, tb::QBLOCK, sizeof(tb::QShared), st>>>(p);      // one block per game still alive
                TB_CUDA(ctx, cudaGetLastError());
                *launches += 1;
            }
            p.st_stage = two ? 3 : 2; p.st_list_in = two ? list3 : list2; p.st_list_out = nullptr; p.st_cut = 0;
            tb::play_games_w6quad_kernel<<<(unsigned)(two ? ctx->stage_cut3 : ctx->stage_cut2), tb::QBLOCK, sizeof(tb::QShared), st>>>(p);
            TB_CUDA(ctx, cudaGetLastError());
        } else {
            p.st_stage = 2; p.st_list_in = list2; p.st_list_out = nullptr; p.st_cut = 0;
            tb::play_games_w6pair_kernel<<<(unsigned)(2 * ctx->sms), 64, 0, st>>>(p);
            TB_CUDA(ctx, cudaGetLastError());
        }
        *launches += 2;
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
        if (!dense) blocks = ctx->sms;
    } else {
        CUfunction cf = nullptr;
        if (ctx->jit && jit_applies(ctx, a) && (lanes == 8 || lanes == 16 || lanes == 32)) {
            cf = ctx->jit->fn[lanes == 8 ? tbjit::K_GEN8 : (lanes == 16 ? tbjit::K_GEN16 : tbjit::K_W32)];
            used_jit = true;
            int jocc = 0;
            if (tbjit::g_api.Occupancy(&jocc, cf, tb::BLOCK, 0) == CUDA_SUCCESS && jocc >= 1) {
                const long long per_block = tb::BLOCK / lanes;
                blocks = (n_games_ll + per_block - 1) / per_block;
                if (blocks > (long long)ctx->sms * jocc) blocks = (long long)ctx->sms * jocc;
            }
        }
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
        { const int rcl = launch_any(ctx, fn, cf, (unsigned)blocks, (unsigned)block, dyn_smem, st, p); if (rcl) return rcl; }
        if (record_events) TB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
        *launches += 1;
    }
    if (a.fitness) {
        tb::fitness_kernel<<<(a.n_candidates + 127) / 128, 128, 0, st>>>(a.n_candidates, a.games_per_candidate, a.results, a.fitness);
        TB_CUDA(ctx, cudaGetLastError());
        *launches += 1;
    }
    ctx->timing.lanes = staged ? 32 : lanes; ctx->timing.grid = (int)blocks; ctx->timing.block = staged ? tb::BLOCK : block;
    ctx->timing.specialised = used_jit ? 1 : 0;
    return TB200_OK;
}

// Limits (include/tetris_b200.h): no NaN / infinite score can arise from weights that pass this
static bool weights_ok(const double* w, size_t n)
{
    for (size_t i = 0; i < n; ++i) if (!(w[i] >= -1e300 && w[i] <= 1e300)) return false;
    return true;
}

static int check_play_args(tb200_ctx* ctx, const tb200_play_args* a)
{
    if (!a || a->struct_size != sizeof(tb200_play_args)) return fail(ctx, TB200_E_BAD_ARG, "play_games: bad struct_size");
    if (ctx->F <= 0) return fail(ctx, TB200_E_NO_FEATURES, "play_games: call tb200_set_features first");
    if (a->n_candidates < 1 || a->games_per_candidate < 1) return fail(ctx, TB200_E_BAD_ARG, "play_games: C and G must be >= 1");
    if ((long long)a->n_candidates * a->games_per_candidate > 0x7fffffffLL) return fail(ctx, TB200_E_BAD_ARG, "play_games: too many games");
    if (!a->weights || !a->results) return fail(ctx, TB200_E_BAD_ARG, "play_games: weights and results are required");
    if (a->T < 1) return fail(ctx, TB200_E_BAD_ARG, "play_games: T must be >= 1");
    if (a->lookahead < 1 || a->lookahead > TB200_MAX_LOOKAHEAD) return fail(ctx, TB200_E_BAD_ARG, "play_games: lookahead must be 1 or 2 (TB200_MAX_LOOKAHEAD)");
    if (a->pieces && a->n_sequences < 1) return fail(ctx, TB200_E_BAD_ARG, "play_games: n_sequences must be >= 1 with explicit pieces");
    if (a->pieces && !a->seq_of_game && a->n_sequences < a->games_per_candidate) return fail(ctx, TB200_E_BAD_ARG, "play_games: need S >= G without seq_of_game");
    if (a->mem != TB200_MEM_HOST && a->mem != TB200_MEM_DEVICE) return fail(ctx, TB200_E_BAD_ARG, "play_games: bad mem");
    {
        const int ms = a->move_set, inner = ms & ~TB200_MOVES_PRESSURE;
        if (ms < 0 || (inner != 0 && inner != TB200_MOVES_OVERHANG_SLIDE && inner != TB200_MOVES_WIGGLE && inner != TB200_MOVES_WIGGLE_DROP))
            return fail(ctx, TB200_E_BAD_ARG, "play_games: bad move_set");
    }
    return TB200_OK;
}

int tb200_play_games(tb200_ctx* ctx, const tb200_play_args* args)
{
    if (!ctx) return TB200_E_BAD_ARG;
    int rc = check_play_args(ctx, args);
    if (rc) return rc;
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    const tb200_play_args& a = *args;
    const size_t n_games = (size_t)a.n_candidates * a.games_per_candidate;
    rc = ensure_lock_ws(ctx, a);
    if (rc) return rc;
    ensure_jit(ctx, a);
    ctx->lock_flag = nullptr;
    ctx->timing = tb200_timing{};
    ctx->timing_valid = false;
    int launches = 0;
    if (a.mem == TB200_MEM_DEVICE) {
        rc = enter_call(ctx, (cudaStream_t)a.stream);
        if (rc) return rc;
        rc = launch_play(ctx, a, (cudaStream_t)a.stream, 0, &launches);
        if (rc) return rc;
        rc = leave_call(ctx, (cudaStream_t)a.stream);
        if (rc) return rc;
        ctx->timing.kernels_launched = launches;
        ctx->timing_valid = true;
        ctx->kernel_events_valid = true;
        return TB200_OK;
    }
    // ---- host buffers: stage through the ctx arena on the ctx's own stream; synchronous ----------
    cudaStream_t st = ctx->own_stream;
    if (!weights_ok(a.weights, (size_t)a.n_candidates * ctx->F)) return fail(ctx, TB200_E_BAD_ARG, "play_games: weights must be finite with |w| <= 1e300");
    const size_t b_w = sizeof(double) * (size_t)a.n_candidates * ctx->F;
    const size_t b_p = a.pieces ? (size_t)a.n_sequences * a.T : 0;
    const size_t b_s = a.seq_of_game ? sizeof(int32_t) * n_games : 0;
    const size_t b_i = a.init_rows ? sizeof(uint16_t) * TB200_ROWS * n_games : 0;
    const size_t b_il = a.init_lines ? sizeof(uint32_t) * n_games : 0;
    const size_t b_r = sizeof(tb200_game_result) * n_games;
    const size_t b_t = a.trace ? 4 * (size_t)a.T * n_games : 0;
    const size_t b_st = a.score_trace ? sizeof(double) * (size_t)a.T * n_games : 0;
    const size_t b_f = a.final_rows ? sizeof(uint16_t) * TB200_ROWS * n_games : 0;
    const size_t b_fit = a.fitness ? sizeof(double) * (size_t)a.n_candidates : 0;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += align_up(b); return o; };
    const size_t o_w = take(b_w), o_p = take(b_p), o_s = take(b_s), o_i = take(b_i), o_r = take(b_r), o_t = take(b_t),
                 o_st = take(b_st), o_f = take(b_f), o_fit = take(b_fit), o_il = take(b_il);
    rc = ensure_arena(ctx, off);
    if (rc) return rc;
    char* base = ctx->d_arena;
    tb200_play_args d = a;
    d.mem = TB200_MEM_DEVICE;
    d.weights = (const double*)(base + o_w);
    d.pieces = a.pieces ? (const uint8_t*)(base + o_p) : nullptr;
    d.seq_of_game = a.seq_of_game ? (const int32_t*)(base + o_s) : nullptr;
    d.init_rows = a.init_rows ? (const uint16_t*)(base + o_i) : nullptr;
    d.init_lines = a.init_lines ? (const uint32_t*)(base + o_il) : nullptr;
    d.results = (tb200_game_result*)(base + o_r);
    d.trace = a.trace ? (uint8_t*)(base + o_t) : nullptr;
    d.score_trace = a.score_trace ? (double*)(base + o_st) : nullptr;
    d.final_rows = a.final_rows ? (uint16_t*)(base + o_f) : nullptr;
    d.fitness = a.fitness ? (double*)(base + o_fit) : nullptr;

    // timing events are recorded outside a captured graph (event nodes of a graph cannot be used for elapsed time),
    // so a graph replay reports total_ms only
    auto enqueue = [&](bool events) -> int {
        TB_CUDA(ctx, cudaMemcpyAsync((void*)d.weights, a.weights, b_w, cudaMemcpyHostToDevice, st));
        if (b_p) TB_CUDA(ctx, cudaMemcpyAsync((void*)d.pieces, a.pieces, b_p, cudaMemcpyHostToDevice, st));
        if (b_s) TB_CUDA(ctx, cudaMemcpyAsync((void*)d.seq_of_game, a.seq_of_game, b_s, cudaMemcpyHostToDevice, st));
        if (b_i) TB_CUDA(ctx, cudaMemcpyAsync((void*)d.init_rows, a.init_rows, b_i, cudaMemcpyHostToDevice, st));
        if (b_il) TB_CUDA(ctx, cudaMemcpyAsync((void*)d.init_lines, a.init_lines, b_il, cudaMemcpyHostToDevice, st));
        if (b_t) TB_CUDA(ctx, cudaMemsetAsync(d.trace, 0, b_t, st));
        if (b_st) TB_CUDA(ctx, cudaMemsetAsync(d.score_trace, 0, b_st, st));
        int rc2 = launch_play(ctx, d, st, 0, &launches, events);
        if (rc2) return rc2;
        TB_CUDA(ctx, cudaMemcpyAsync(a.results, d.results, b_r, cudaMemcpyDeviceToHost, st));
        if (b_t) TB_CUDA(ctx, cudaMemcpyAsync(a.trace, d.trace, b_t, cudaMemcpyDeviceToHost, st));
        if (b_st) TB_CUDA(ctx, cudaMemcpyAsync(a.score_trace, d.score_trace, b_st, cudaMemcpyDeviceToHost, st));
        if (b_f) TB_CUDA(ctx, cudaMemcpyAsync(a.final_rows, d.final_rows, b_f, cudaMemcpyDeviceToHost, st));
        if (b_fit) TB_CUDA(ctx, cudaMemcpyAsync(a.fitness, d.fitness, b_fit, cudaMemcpyDeviceToHost, st));
        return TB200_OK;
    };
    bool kernel_events = true;
    rc = enter_call(ctx, st);                  // outside any capture: a device-mode call may still be running on another stream
    if (rc) return rc;
    ctx->busy_recorded = false;                // this call ends synchronised
    TB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    // ---- graph replay: identical call (same pinned buffers, sizes, options, feature list) as last time ----
    const bool same = ctx->graph_valid && memcmp(&ctx->graph_args, &a, sizeof a) == 0 && ctx->graph_F == ctx->F &&
                      ctx->graph_mode == ctx->mode && memcmp(ctx->graph_ids, ctx->ids, sizeof ctx->ids) == 0 && ctx->graph_arena == base;
    bool done = false;
    if (same) {
        if (cudaGraphLaunch(ctx->graph_exec, st) == cudaSuccess) { ctx->timing = ctx->graph_timing; launches = ctx->graph_timing.kernels_launched; ctx->lock_flag = ctx->graph_lock_flag; done = true; kernel_events = false; }
        else { cudaGetLastError(); drop_graph(ctx); }
    } else if (ctx->graphs_enabled && is_pinned(a.weights) && is_pinned(a.pieces) && is_pinned(a.seq_of_game) && is_pinned(a.init_rows) && is_pinned(a.init_lines) &&
               is_pinned(a.results) && is_pinned(a.trace) && is_pinned(a.score_trace) && is_pinned(a.final_rows) && is_pinned(a.fitness)) {
        drop_graph(ctx);
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc2 = enqueue(false);
            cudaGraph_t gr = nullptr;
            const cudaError_t e1 = cudaStreamEndCapture(st, &gr);
            if (rc2 == TB200_OK && e1 == cudaSuccess && gr && cudaGraphInstantiate(&ctx->graph_exec, gr, 0) == cudaSuccess) {
                ctx->graph = gr; ctx->graph_args = a; ctx->graph_F = ctx->F; ctx->graph_mode = ctx->mode;
                memcpy(ctx->graph_ids, ctx->ids, sizeof ctx->ids); ctx->graph_arena = base;
                ctx->timing.kernels_launched = launches;
                ctx->graph_timing = ctx->timing; ctx->graph_lock_flag = ctx->lock_flag; ctx->graph_valid = true;
                if (cudaGraphLaunch(ctx->graph_exec, st) == cudaSuccess) { done = true; kernel_events = false; }
                else { cudaGetLastError(); drop_graph(ctx); ctx->graphs_enabled = false; }
            } else {
                cudaGetLastError();
                if (gr) cudaGraphDestroy(gr);
                ctx->graph_exec = nullptr; ctx->graphs_enabled = false;      // capture not possible here: use the direct path from now on
                launches = 0;
            }
        } else {
            cudaGetLastError(); ctx->graphs_enabled = false;
        }
    }
    if (!done) {
        launches = 0;
        rc = enqueue(true);
        if (rc) return rc;
    }
    TB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    if (ctx->lock_flag) {                   // K2s ran: its bucket waits are bounded; a timeout means incomplete results
        uint32_t flag = 0;
        TB_CUDA(ctx, cudaMemcpy(&flag, ctx->lock_flag, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag) return fail(ctx, TB200_E_CUDA, "play_games: lockstep kernel timed out waiting for a stage (results incomplete)");
    }
    ctx->kernel_events_valid = kernel_events;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    ctx->timing.total_ms = ms;
    ctx->timing.h2d_bytes = b_w + b_p + b_s + b_i + b_il;
    ctx->timing.d2h_bytes = b_r + b_t + b_st + b_f + b_fit;
    ctx->timing.kernels_launched = launches;
    ctx->timing_valid = true;
    return TB200_OK;
}

int tb200_jit_build(const int32_t* feature_ids, int32_t F, char* log, int32_t log_len)
{
    if (!feature_ids || F < 1 || F > TB200_MAX_WEIGHTED) return TB200_E_BAD_ARG;
    uint8_t ids[TB200_MAX_WEIGHTED];
    for (int i = 0; i < F; ++i) { if (feature_ids[i] < 0 || feature_ids[i] >= TB200_NUM_FEATURES) return TB200_E_BAD_ARG; ids[i] = (uint8_t)feature_ids[i]; }
    std::vector<char> cubin;
    std::string why;
    bool cached = false;
    const bool ok = tbjit::build_cubin(ids, F, cubin, why, &cached);
    if (log && log_len > 0) snprintf(log, (size_t)log_len, "%s", ok ? (cached ? "cached" : "compiled") : why.c_str());
    return ok ? (int)(cubin.size() >> 10) : TB200_E_CUDA;
}

int tb200_check_status(tb200_ctx* ctx)
{
    if (!ctx) return TB200_E_BAD_ARG;
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->busy_recorded) TB_CUDA(ctx, cudaEventSynchronize(ctx->busy));
    if (ctx->lock_flag) {
        uint32_t flag = 0;
        TB_CUDA(ctx, cudaMemcpy(&flag, ctx->lock_flag, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag) return fail(ctx, TB200_E_CUDA, "check_status: a bounded device-side wait of the lockstep / piece-binned scheduler expired (results incomplete)");
    }
    return TB200_OK;
}

int tb200_best_move(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                    const uint16_t* boards, const uint8_t* piece, const uint8_t* next,
                    const double* weights, int32_t per_board_weights, int32_t lookahead,
                    uint8_t* out_move, double* out_score, uint32_t* out_count, uint16_t* out_rows)
{
    return tb200_best_move_ex(ctx, mem, stream, N, boards, piece, next, weights, per_board_weights, lookahead, nullptr,
                              out_move, out_score, out_count, out_rows);
}

int tb200_best_move_ex(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                       const uint16_t* boards, const uint8_t* piece, const uint8_t* next,
                       const double* weights, int32_t per_board_weights, int32_t lookahead, const tb200_move_opts* opts,
                       uint8_t* out_move, double* out_score, uint32_t* out_count, uint16_t* out_rows)
{
    if (!ctx) return TB200_E_BAD_ARG;
    if (opts) {
        if (opts->struct_size != sizeof(tb200_move_opts)) return fail(ctx, TB200_E_BAD_ARG, "best_move: bad opts struct_size");
        const int ms = opts->move_set, inner = ms & ~TB200_MOVES_PRESSURE;
        if (ms < 0 || (inner != 0 && inner != TB200_MOVES_OVERHANG_SLIDE && inner != TB200_MOVES_WIGGLE && inner != TB200_MOVES_WIGGLE_DROP))
            return fail(ctx, TB200_E_BAD_ARG, "best_move: bad move_set");
    }
    const uint32_t* lines = opts ? opts->lines : nullptr;
    if (ctx->F <= 0) return fail(ctx, TB200_E_NO_FEATURES, "best_move: call tb200_set_features first");
    if (N < 1 || !boards || !piece || !weights || !out_move) return fail(ctx, TB200_E_BAD_ARG, "best_move: missing argument");
    if (lookahead < 1 || lookahead > TB200_MAX_LOOKAHEAD) return fail(ctx, TB200_E_BAD_ARG, "best_move: lookahead must be 1 or 2 (TB200_MAX_LOOKAHEAD)");
    if (mem != TB200_MEM_HOST && mem != TB200_MEM_DEVICE) return fail(ctx, TB200_E_BAD_ARG, "best_move: bad mem");
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = mem == TB200_MEM_HOST;
    cudaStream_t st = host ? ctx->own_stream : (cudaStream_t)stream;
    const size_t n = (size_t)N;
    const size_t nw = per_board_weights ? n : 1;
    // arena: [staged inputs/outputs when host] + internal (pieces[N][2], results, trace[N][2], strace[N][2])
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += align_up(b); return o; };
    const size_t o_seq = take(2 * n), o_res = take(sizeof(tb200_game_result) * n), o_tr = take(8 * n), o_str = take(16 * n);
    const size_t o_b = take(host ? 40 * n : 0), o_p = take(host ? n : 0), o_n = take(host && next ? n : 0),
                 o_w = take(host ? sizeof(double) * nw * ctx->F : 0), o_mv = take(host ? 4 * n : 0),
                 o_sc = take(host && out_score ? 8 * n : 0), o_ct = take(host && out_count ? 4 * n : 0),
                 o_rows = take(host && out_rows ? 40 * n : 0), o_ln = take(host && lines ? 4 * n : 0);
    if (host && !weights_ok(weights, nw * (size_t)ctx->F)) return fail(ctx, TB200_E_BAD_ARG, "best_move: weights must be finite with |w| <= 1e300");
    int rc = ensure_arena(ctx, off);
    if (rc) return rc;
    rc = enter_call(ctx, st);
    if (rc) return rc;
    ctx->lock_flag = nullptr;
    char* base = ctx->d_arena;
    const uint16_t* d_boards = boards; const uint8_t* d_piece = piece; const uint8_t* d_next = next; const double* d_w = weights;
    uint8_t* d_mv = out_move; double* d_sc = out_score; uint32_t* d_ct = out_count; uint16_t* d_rows = out_rows;
    if (host) {
        d_boards = (const uint16_t*)(base + o_b); d_piece = (const uint8_t*)(base + o_p); d_next = next ? (const uint8_t*)(base + o_n) : nullptr;
        d_w = (const double*)(base + o_w); d_mv = (uint8_t*)(base + o_mv); d_sc = out_score ? (double*)(base + o_sc) : nullptr;
        d_ct = out_count ? (uint32_t*)(base + o_ct) : nullptr; d_rows = out_rows ? (uint16_t*)(base + o_rows) : nullptr;
        TB_CUDA(ctx, cudaMemcpyAsync((void*)d_boards, boards, 40 * n, cudaMemcpyHostToDevice, st));
        TB_CUDA(ctx, cudaMemcpyAsync((void*)d_piece, piece, n, cudaMemcpyHostToDevice, st));
        if (next) TB_CUDA(ctx, cudaMemcpyAsync((void*)d_next, next, n, cudaMemcpyHostToDevice, st));
        TB_CUDA(ctx, cudaMemcpyAsync((void*)d_w, weights, sizeof(double) * nw * ctx->F, cudaMemcpyHostToDevice, st));
        if (lines) {
            TB_CUDA(ctx, cudaMemcpyAsync(base + o_ln, lines, 4 * n, cudaMemcpyHostToDevice, st));
            lines = (const uint32_t*)(base + o_ln);
        }
    }
    // two-piece sequence per board: {piece, next}; T = 2 when a next piece is visible for lookahead 2, else 1.
    uint8_t* d_seq = (uint8_t*)(base + o_seq);
    tb::pair_pieces_kernel<<<(N + 255) / 256, 256, 0, st>>>(N, d_piece, d_next, d_seq);
    TB_CUDA(ctx, cudaGetLastError());
    tb200_play_args a;
    memset(&a, 0, sizeof a);
    a.struct_size = sizeof a; a.mem = TB200_MEM_DEVICE; a.stream = st;
    a.n_candidates = N; a.games_per_candidate = 1; a.weights = d_w; a.pieces = d_seq; a.n_sequences = N;
    a.T = 2; a.max_moves = 1; a.lookahead = lookahead; a.lanes = 0; a.init_rows = d_boards;
    a.results = (tb200_game_result*)(base + o_res); a.trace = (uint8_t*)(base + o_tr); a.score_trace = (double*)(base + o_str);
    a.final_rows = d_rows;
    if (opts) {
        a.move_set = opts->move_set; a.pressure_two_but_rot = opts->pressure_two_but_rot;
        a.pressure_avg_lat = opts->pressure_avg_lat; a.pressure_resp_time = opts->pressure_resp_time; a.pressure_move_eff = opts->pressure_move_eff;
        a.init_lines = lines;
    }
    int launches = 1;
    ctx->timing = tb200_timing{};
    rc = launch_play(ctx, a, st, per_board_weights ? 1 : -1, &launches);
    if (rc) return rc;
    tb::best_move_pack_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, a.results, (const uint32_t*)a.trace, a.score_trace, d_mv, d_sc, d_ct);
    TB_CUDA(ctx, cudaGetLastError());
    launches += 1;
    ctx->timing.kernels_launched = launches;
    ctx->timing_valid = true;
    if (host) {
        TB_CUDA(ctx, cudaMemcpyAsync(out_move, d_mv, 4 * n, cudaMemcpyDeviceToHost, st));
        if (out_score) TB_CUDA(ctx, cudaMemcpyAsync(out_score, d_sc, 8 * n, cudaMemcpyDeviceToHost, st));
        if (out_count) TB_CUDA(ctx, cudaMemcpyAsync(out_count, d_ct, 4 * n, cudaMemcpyDeviceToHost, st));
        if (out_rows) TB_CUDA(ctx, cudaMemcpyAsync(out_rows, d_rows, 40 * n, cudaMemcpyDeviceToHost, st));
        TB_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->busy_recorded = false;
        return TB200_OK;
    }
    return leave_call(ctx, st);
}

int tb200_eval_features_at(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                           const uint16_t* prev_boards, const uint8_t* piece, const uint8_t* rot, const uint8_t* col, const uint8_t* row,
                           double* out_features, int32_t* out_info, uint16_t* out_rows)
{
    if (!ctx) return TB200_E_BAD_ARG;
    if (N < 1 || !prev_boards || !piece || !rot || !col || !out_features || !out_info) return fail(ctx, TB200_E_BAD_ARG, "eval_features: missing argument");
    if (mem != TB200_MEM_HOST && mem != TB200_MEM_DEVICE) return fail(ctx, TB200_E_BAD_ARG, "eval_features: bad mem");
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)N;
    if (mem == TB200_MEM_DEVICE) {
        tb::eval_features_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, prev_boards, piece, rot, col, row, out_features, out_info, out_rows);
        TB_CUDA(ctx, cudaGetLastError());
        return TB200_OK;
    }
    cudaStream_t st = ctx->own_stream;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += align_up(b); return o; };
    const size_t o_b = take(40 * n), o_p = take(n), o_r = take(n), o_c = take(n), o_f = take(8 * 45 * n), o_i = take(12 * n), o_rows = take(out_rows ? 40 * n : 0),
                 o_ar = take(row ? n : 0);
    int rc = ensure_arena(ctx, off);
    if (rc) return rc;
    rc = enter_call(ctx, st);
    if (rc) return rc;
    char* base = ctx->d_arena;
    TB_CUDA(ctx, cudaMemcpyAsync(base + o_b, prev_boards, 40 * n, cudaMemcpyHostToDevice, st));
    TB_CUDA(ctx, cudaMemcpyAsync(base + o_p, piece, n, cudaMemcpyHostToDevice, st));
    TB_CUDA(ctx, cudaMemcpyAsync(base + o_r, rot, n, cudaMemcpyHostToDevice, st));
    TB_CUDA(ctx, cudaMemcpyAsync(base + o_c, col, n, cudaMemcpyHostToDevice, st));
    if (row) TB_CUDA(ctx, cudaMemcpyAsync(base + o_ar, row, n, cudaMemcpyHostToDevice, st));
    tb::eval_features_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, (const uint16_t*)(base + o_b), (const uint8_t*)(base + o_p),
        (const uint8_t*)(base + o_r), (const uint8_t*)(base + o_c), row ? (const uint8_t*)(base + o_ar) : nullptr, (double*)(base + o_f), (int32_t*)(base + o_i),
        out_rows ? (uint16_t*)(base + o_rows) : nullptr);
    TB_CUDA(ctx, cudaGetLastError());
    TB_CUDA(ctx, cudaMemcpyAsync(out_features, base + o_f, 8 * 45 * n, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaMemcpyAsync(out_info, base + o_i, 12 * n, cudaMemcpyDeviceToHost, st));
    if (out_rows) TB_CUDA(ctx, cudaMemcpyAsync(out_rows, base + o_rows, 40 * n, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->busy_recorded = false;
    return TB200_OK;
}

int tb200_eval_features(tb200_ctx* ctx, int32_t mem, void* stream, int32_t N,
                        const uint16_t* prev_boards, const uint8_t* piece, const uint8_t* rot, const uint8_t* col,
                        double* out_features, int32_t* out_info, uint16_t* out_rows)
{
    return tb200_eval_features_at(ctx, mem, stream, N, prev_boards, piece, rot, col, nullptr, out_features, out_info, out_rows);
}

int tb200_hashed_pieces(tb200_ctx* ctx, int32_t mem, void* stream, uint64_t seed, uint64_t first_stream,
                        int32_t n_streams, uint32_t T, uint8_t* out)
{
    if (!ctx) return TB200_E_BAD_ARG;
    if (n_streams < 1 || T < 1 || !out) return fail(ctx, TB200_E_BAD_ARG, "hashed_pieces: missing argument");
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t total = (size_t)n_streams * T;
    const int grid = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    if (mem == TB200_MEM_DEVICE) {
        tb::hashed_pieces_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, first_stream, n_streams, T, out);
        TB_CUDA(ctx, cudaGetLastError());
        return TB200_OK;
    }
    int rc = ensure_arena(ctx, total);
    if (rc) return rc;
    cudaStream_t st = ctx->own_stream;
    rc = enter_call(ctx, st);
    if (rc) return rc;
    tb::hashed_pieces_kernel<<<grid, 256, 0, st>>>(seed, first_stream, n_streams, T, (uint8_t*)ctx->d_arena);
    TB_CUDA(ctx, cudaGetLastError());
    TB_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_arena, total, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->busy_recorded = false;
    return TB200_OK;
}

}  // extern "C"
