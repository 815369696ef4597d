// bk_env_kernels.cuh — bodies of the game-batch kernels, written as warp-collective device
// functions of (game index g, lane) so that the thin __global__ wrappers in bk_env.cu and the
// CPU warp emulator used by the tests (tests/warp_emu/) execute the very same source.
#pragma once
#include "bk_game.cuh"
#include "bk_rng.cuh"
#include "bk_playout.cuh"

#ifndef BK_ERR_ILLEGAL_MOVE_CODE
#define BK_ERR_ILLEGAL_MOVE_CODE (-3)
#endif
#define BK_PLAYOUT_HASH_FLAG 1u
#define BK_PLAYOUT_MIN_TILE_FLAG 2u
#define BK_PLAYOUT_MAX_TILE_FLAG 4u
#define BK_PLAYOUT_NEW_GAME_FLAG 8u

// per-game scalars gathered by one kernel for the cheap accessors
struct BkSummary {
    int32_t cur;
    int32_t terminal;
    int32_t active[4];
    int32_t scores[4];
    float payoff[4];
    uint32_t pieces[4];
    int32_t lastlens[4];
    int32_t ply;
    int32_t pad;
    uint64_t digest;
};

__device__ __forceinline__ void bk_flush_counters(const BkCounters& ctr, uint32_t steps, int lane,
                                                  unsigned long long* counters) {
    const unsigned crem = __reduce_add_sync(BK_FULL, ctr.crem);
    if (lane == 0 && counters) {
        atomicAdd(&counters[0], (unsigned long long)steps);
        atomicAdd(&counters[1], (unsigned long long)ctr.movegens);
        atomicAdd(&counters[2], 120ull * (unsigned long long)crem);
    }
}

__device__ __forceinline__ void kb_reset(BkState* __restrict__ states, int g, int lane) {
    BkRegs G;
    BkCounters ctr = {0u, 0u};
    bk_reset(G, lane, ctr);
    bk_store(&states[g], lane, G);
}

__device__ __forceinline__ void kb_apply(BkState* __restrict__ states, uint16_t* __restrict__ hist,
                                         const int32_t* __restrict__ tiles, const int32_t* __restrict__ finish,
                                         int32_t* __restrict__ status, unsigned long long* counters, int g, int lane,
                                         const BkTabs& tabs) {
    const int tile = tiles[g];
    if (tile < 0) { if (lane == 0) status[g] = 1; return; }
    BkRegs G;
    bk_load(&states[g], lane, G);
    BkCounters ctr = {0u, 0u};
    const int p = bk_cur(G);
    const uint32_t ply = G.ply;
    const bool ok = bk_apply(G, tile, finish ? finish[g] : -1, lane, tabs, ctr);
    if (ok) {
        bk_store(&states[g], lane, G);
        if (lane == 0 && ply < BK_HIST_CAP) hist[size_t(g) * BK_HIST_CAP + ply] = uint16_t(tile | (p << 9));
    }
    if (lane == 0) status[g] = ok ? 0 : BK_ERR_ILLEGAL_MOVE_CODE;
    bk_flush_counters(ctr, ok ? 1u : 0u, lane, counters);
}

// Game::place_piece (game.rs:116-144): validity of (piece index in remaining list, variant, offset)
// against the current board (Board::is_valid_move, board.rs:62-92), then the whole piece at once —
// the end state equals the reference's tile-by-tile apply with Some(p) on the last tile.
__device__ __forceinline__ void kb_place_piece(BkState* __restrict__ states, uint16_t* __restrict__ hist,
                                               const int32_t* __restrict__ pp, const int32_t* __restrict__ vv,
                                               const int32_t* __restrict__ oo, int32_t* __restrict__ status, int g,
                                               int lane) {
    const int pi = pp[g];
    if (pi < 0) { if (lane == 0) status[g] = 1; return; }
    BkRegs G;
    bk_load(&states[g], lane, G);
    BkCounters ctr = {0u, 0u};
    const int p = bk_cur(G);
    const uint32_t pieces = bk_sel4(p, G.pc0, G.pc1, G.pc2, G.pc3);
    const int pid = bk_nth_set_bit(pieces, pi);
    const int vi = vv[g], off = oo[g];
    bool ok = !bk_terminal(G) && ((G.meta >> 6) & 7u) == 0u && pid >= 0 && vi >= 0 && off >= 0;
    int gv = 0;
    if (ok) {
        gv = int(c_piece_first_variant[pid]) + vi;
        ok = gv < int(c_piece_first_variant[pid + 1]);
    }
    int ncells = 0;
    uint32_t mask = 0u;
    if (ok) {
        ncells = int(c_variant_ncells[gv]);
        const int len = (int(c_variant_height[gv]) - 1) * 20 + int(c_variant_width[gv]);
        ok = (off + len <= 400) && (off % 20 + int(c_variant_width[gv]) <= 20);  // board.rs:71-75
    }
    if (ok) {  // warp-uniform
        for (int j = 0; j < ncells; ++j) {
            const int t = off + int(c_variant_offsets[gv][j]);
            if (t / 20 == lane) mask |= 1u << (t % 20);
        }
        uint32_t free_, anch;
        bk_free_anchor(bk_sel4(p, G.o0, G.o1, G.o2, G.o3), G.o0 | G.o1 | G.o2 | G.o3, p, lane, free_, anch);
        const bool blocked = __any_sync(BK_FULL, (mask & ~free_) != 0u);
        const bool anchored = __any_sync(BK_FULL, (mask & anch) != 0u);
        ok = !blocked && anchored;
    }
    if (ok) {
        if (p == 0) G.o0 |= mask; else if (p == 1) G.o1 |= mask; else if (p == 2) G.o2 |= mask; else G.o3 |= mask;
        if (lane == 0)
            for (int j = 0; j < ncells; ++j)
                if (G.ply + j < BK_HIST_CAP)
                    hist[size_t(g) * BK_HIST_CAP + G.ply + j] = uint16_t((off + int(c_variant_offsets[gv][j])) | (p << 9));
        G.ply += uint32_t(ncells);
        const uint32_t clr = ~(1u << pid);
        if (p == 0) G.pc0 &= clr; else if (p == 1) G.pc1 &= clr; else if (p == 2) G.pc2 &= clr; else G.pc3 &= clr;
        G.lastlens = (G.lastlens & ~(0xFFu << (8 * p))) | (uint32_t(c_piece_points[pid]) << (8 * p));
        bk_advance(G, lane, ctr);
        bk_store(&states[g], lane, G);
    }
    if (lane == 0) status[g] = ok ? 0 : BK_ERR_ILLEGAL_MOVE_CODE;
}

// Persistent lockstep playout: one warp plays its game to the end without leaving the SM.  bk_playout.cuh keeps the
// turn in progress in registers and touches the bitboards once per piece.
struct BkPlayoutCtx {
    uint64_t seed;
    uint32_t game_id, flags, ply_end;
    uint16_t* __restrict__ h16;
    uint64_t h;
};

template <bool HASH>
__device__ __forceinline__ void bk_playout_digest(BkPlayoutCtx& C, const BkRegs& G, const BkTurn& T, int p, int tile, int lane) {
    if (!HASH || !(C.flags & BK_PLAYOUT_HASH_FLAG)) return;
    BkRegs H = G;
    if (T.nT) bk_turn_materialise(H, T, lane);
    C.h = bk_splitmix64(C.h ^ bk_digest(H, lane));
    C.h = bk_splitmix64(C.h ^ (uint64_t(p) | (uint64_t(tile) << 8)));
}

// CHECKED = the instantiation that also serves the tests' seed-free policies (lowest / highest tile) and the per-ply digest
template <bool CHECKED>
__device__ __forceinline__ int bk_playout_draw(const BkPlayoutCtx& C, uint32_t ply, int cnt, BkPlayoutRng& rng) {
    if (CHECKED && (C.flags & BK_PLAYOUT_MIN_TILE_FLAG)) return 0;
    if (CHECKED && (C.flags & BK_PLAYOUT_MAX_TILE_FLAG)) return cnt - 1;
    return int(bk_playout_index(C.seed, C.game_id, ply, uint32_t(cnt), rng));
}

// the moves of a turn after its first tile, until the piece is complete (legal set empty) or the ply budget ends
template <bool HASH>
__device__ __forceinline__ void bk_playout_turn_moves(BkPlayoutCtx& C, BkRegs& G, BkTurn& T, BkPlayoutRng& rng, int p,
                                                      int lane, const BkTabs& tabs) {
    while ((T.w0 | T.w1 | T.w2) != 0u && G.ply < C.ply_end) {
        const int c0 = __popc(T.w0), c1 = __popc(T.w1), cnt = c0 + c1 + __popc(T.w2);
        const int idx = bk_playout_draw<HASH>(C, G.ply, cnt, rng);
        int k, wbit;
        uint32_t b;
        bk_turn_pick(T, idx, c0, c1, k, b, wbit);
        const int tile = bk_window_tile(wbit, T.tr, T.tc);
        T.tq = (T.tq << 7) | uint32_t(wbit);
        T.nT += 1;
        bk_turn_next(G, T, k, b, lane, tabs);
        G.ply += 1u;
        if (HASH && (T.w0 | T.w1 | T.w2) != 0u) bk_playout_digest<HASH>(C, G, T, p, tile, lane);
        if (HASH) T.last = tile;
    }
}

// game.rs:176-191: the piece is complete — its squares join the mover's board, the piece leaves the hand, its size
// is remembered, the turn passes
template <bool HASH>
__device__ __forceinline__ void bk_playout_commit(BkPlayoutCtx& C, BkRegs& G, BkTurn& T, int p, int lane,
                                                  const BkTabs& tabs, BkCounters& ctr, uint32_t& free_, uint32_t& anch) {
    const int pid = bk_turn_piece(G, T, lane, tabs);
    bk_turn_record(T, G.ply, p, lane, C.h16);
    const uint32_t rows = bk_window_to_row(G.tw0, G.tw1, G.tw2, T.tr, T.tc, lane);
    const uint32_t clr = ~(1u << pid);
    if (p == 0) { G.o0 |= rows; G.pc0 &= clr; } else if (p == 1) { G.o1 |= rows; G.pc1 &= clr; }
    else if (p == 2) { G.o2 |= rows; G.pc2 &= clr; } else { G.o3 |= rows; G.pc3 &= clr; }
    G.lastlens = (G.lastlens & ~(0xFFu << (8 * p))) | (uint32_t(T.nT) << (8 * p));
    T.nT = 0;
    bk_advance(G, lane, ctr, free_, anch);            // leaves the next mover's turn-start rows in free_/anch
    if (HASH) bk_playout_digest<HASH>(C, G, T, p, T.last, lane);
}

template <bool HASH>
__device__ __forceinline__ void kb_playout(BkState* __restrict__ states, uint16_t* __restrict__ hist, uint64_t seed,
                                           uint32_t game_id, int max_plies, uint32_t flags,
                                           int32_t* __restrict__ steps_out, uint64_t* __restrict__ hash_out,
                                           unsigned long long* counters, int g, int lane, const BkTabs& tabs) {
    BkRegs G;
    BkCounters ctr = {0u, 0u};
    if (flags & BK_PLAYOUT_NEW_GAME_FLAG) bk_reset(G, lane, ctr);     // Game::reset inside the launch (its move generation counts as the launch's work)
    else bk_load(&states[g], lane, G);
    BkPlayoutCtx C;
    C.seed = seed; C.game_id = game_id; C.flags = flags; C.h = 0ull;
    C.h16 = hist + size_t(g) * BK_HIST_CAP;
    C.ply_end = max_plies < 0 ? 0xffffffffu : G.ply + uint32_t(max_plies);
    BkPlayoutRng rng;
    rng.w0 = rng.w1 = rng.w2 = rng.w3 = 0u;
    rng.block = 0xffffffffu;
    const uint32_t ply0 = G.ply;
    BkTurn T;
    bk_turn_resume(G, T, lane, tabs);                    // the stored state may be in the middle of a turn
    G.meta &= ~(7u << 6);                                // |T| lives in T.nT while the kernel runs
    G.t01 = 0u; G.t23 = 0u;
    uint32_t free_ = 0u, anch = 0u;                      // turn-start rows of the seat to move
    if (!bk_terminal(G) && T.nT == 0) {
        const int p = bk_cur(G);
        bk_free_anchor(bk_sel4(p, G.o0, G.o1, G.o2, G.o3), G.o0 | G.o1 | G.o2 | G.o3, p, lane, free_, anch);
    }
    if (!bk_terminal(G) && T.nT) {
        const int p = bk_cur(G);
        bk_playout_turn_moves<HASH>(C, G, T, rng, p, lane, tabs);
        if ((T.w0 | T.w1 | T.w2) == 0u) bk_playout_commit<HASH>(C, G, T, p, lane, tabs, ctr, free_, anch);
    }
    while (T.nT == 0 && !bk_terminal(G) && G.ply < C.ply_end) {
        // turn start: the legal set is the board rows left by the move generator
        const int p = bk_cur(G);
        const int cnt = bk_legal_count(G.legal);
        const int idx = bk_playout_draw<HASH>(C, G.ply, cnt, rng);
        int tr, tc;
        bk_legal_select_rc(G.legal, idx, lane, tr, tc);
        const int tile = tr * 20 + tc;
        bk_turn_first(G, T, free_, anch, bk_sel4(p, G.pc0, G.pc1, G.pc2, G.pc3), tr, tc, lane, tabs);
        G.ply += 1u;
        if (HASH) { T.last = tile; if ((T.w0 | T.w1 | T.w2) != 0u) bk_playout_digest<HASH>(C, G, T, p, tile, lane); }
        bk_playout_turn_moves<HASH>(C, G, T, rng, p, lane, tabs);
        if ((T.w0 | T.w1 | T.w2) == 0u) bk_playout_commit<HASH>(C, G, T, p, lane, tabs, ctr, free_, anch);
    }
    if (T.nT) {
        bk_turn_record(T, G.ply, bk_cur(G), lane, C.h16);
        bk_turn_materialise(G, T, lane);
    }
    bk_store(&states[g], lane, G);
    if (flags & BK_PLAYOUT_NEW_GAME_FLAG)                    // bk_env_reset clears the history: the rest of the row
        for (uint32_t i = G.ply + uint32_t(lane); i < BK_HIST_CAP; i += 32u) C.h16[i] = 0;
    const int steps = int(G.ply - ply0);
    if (lane == 0) { steps_out[g] = steps; hash_out[g] = C.h; }
    bk_flush_counters(ctr, uint32_t(steps), lane, counters);
}

__device__ __forceinline__ void kb_scores(const BkState* __restrict__ states, int32_t* __restrict__ plies,
                                          int32_t* __restrict__ scores, int g, int lane) {
    BkRegs G;
    bk_load(&states[g], lane, G);
    int sc[4];
    bk_scores(G, sc);
    if (lane == 0) {
        plies[g] = int(G.ply);
        for (int p = 0; p < 4; ++p) scores[g * 4 + p] = sc[p];
    }
}

__device__ __forceinline__ void kb_summary(const BkState* __restrict__ states, BkSummary* __restrict__ out, int g,
                                           int lane) {
    BkRegs G;
    bk_load(&states[g], lane, G);
    int sc[4];
    float pay[4];
    bk_scores(G, sc);
    bk_payoff(G, pay);
    const uint64_t dg = bk_digest(G, lane);
    if (lane == 0) {
        BkSummary s;
        s.cur = bk_cur(G);
        s.terminal = bk_terminal(G) ? 1 : 0;
        const uint32_t pcs[4] = {G.pc0, G.pc1, G.pc2, G.pc3};
        for (int p = 0; p < 4; ++p) {
            s.active[p] = ((bk_elim(G) >> p) & 1u) ? 0 : 1;
            s.scores[p] = sc[p];
            s.payoff[p] = pay[p];
            s.pieces[p] = pcs[p];
            s.lastlens[p] = int((G.lastlens >> (8 * p)) & 0xFFu);
        }
        s.ply = int(G.ply);
        s.pad = 0;
        s.digest = dg;
        out[g] = s;
    }
}

// what: 0 legal mask, 1 reference board bytes (board.rs:95-119), 2 anchors of `player` (<0: current).
// Plain per-cell code (no collectives): thread tid of nthreads strides over the 400 cells of game g.
__device__ __forceinline__ void kb_cells(const BkState* __restrict__ states, uint8_t* __restrict__ out, int what,
                                         int player, int g, int tid, int nthreads) {
    const BkState* s = &states[g];
    for (int t = tid; t < 400; t += nthreads) {
        const int r = t / 20, c = t % 20;
        uint8_t v = 0;
        if (what == 0) {
            v = (s->legal[r] >> c) & 1u;
        } else {
            int owner = 0;
            uint8_t adjbits = 0;
            for (int p = 0; p < 4; ++p) {
                if ((s->own[r][p] >> c) & 1u) owner = p + 1;
                bool adj = false;
                if (c > 0) adj |= (s->own[r][p] >> (c - 1)) & 1u;
                if (c < 19) adj |= (s->own[r][p] >> (c + 1)) & 1u;
                if (r > 0) adj |= (s->own[r - 1][p] >> c) & 1u;
                if (r < 19) adj |= (s->own[r + 1][p] >> c) & 1u;
                if (adj) adjbits |= uint8_t(1u << (4 + p));
            }
            if (what == 1) {
                v = owner ? uint8_t(0xF0u | owner) : adjbits;
            } else {
                const int p = player < 0 ? int(s->meta & 3u) : player;
                bool diag = false;
                if (r > 0 && c > 0) diag |= (s->own[r - 1][p] >> (c - 1)) & 1u;
                if (r > 0 && c < 19) diag |= (s->own[r - 1][p] >> (c + 1)) & 1u;
                if (r < 19 && c > 0) diag |= (s->own[r + 1][p] >> (c - 1)) & 1u;
                if (r < 19 && c < 19) diag |= (s->own[r + 1][p] >> (c + 1)) & 1u;
                const int start = p == 0 ? 0 : (p == 1 ? 19 : (p == 2 ? 399 : 380));
                const bool restricted = owner != 0 || (adjbits & (1u << (4 + p)));
                v = ((diag || t == start) && !restricted) ? 1 : 0;
            }
        }
        out[size_t(g) * 400 + t] = v;
    }
}

// Game::get_board_state (game.rs:283-311).  Plane k < 4 holds the squares of seat (cur + k) % 4,
// plane 4 the legal tiles; the result is turned `cur` quarter turns: new[j][k] = old[k][19-j].
template <typename T>
__device__ __forceinline__ void kb_planes(const BkState* __restrict__ s, T* __restrict__ out, int tid, int nthreads) {
    const int cur = int(s->meta & 3u);
    for (int e = tid; e < 2000; e += nthreads) {
        const int plane = e / 400, j = (e % 400) / 20, k = e % 20;
        int r, c;
        if (cur == 0) { r = j; c = k; }
        else if (cur == 1) { r = k; c = 19 - j; }
        else if (cur == 2) { r = 19 - j; c = 19 - k; }
        else { r = 19 - k; c = j; }
        const uint32_t row = plane < 4 ? s->own[r][(plane + cur) & 3] : s->legal[r];
        out[e] = T((row >> c) & 1u);
    }
}
