// bk_playout.cuh — the persistent random playout, organised by TURN (config 2's dominant kernel).
//
// Same rules as bk_game.cuh's Game::apply (game.rs:150-194), same random policy, same traces — only the
// bookkeeping is arranged for a warp that plays a whole game without leaving its registers:
//   * while a turn is in progress the legal set, the tiles laid (T) and — once at most 32 placements survive —
//     the surviving candidates' window masks all live in REGISTERS in the 9x9-window form round T[0]; a mid-turn
//     move is then: count (3 popc), draw, pick a window bit, one register test per lane, three redux.or;
//   * the mover's bitboard receives the whole piece at the COMMIT (the window mask of T scattered back to rows);
//     nothing reads it in between (narrowing works from the turn-start boards, game.rs:165-173 only intersects);
//   * the piece id is looked up once, at the commit, from the unique surviving placement;
//   * BkState's turn fields (|T|, the tiles, the cached legal rows, the narrowing cache) are materialised only when
//     the kernel stops in the middle of a turn (max_plies) or when a digest is asked for after every ply.
#pragma once
#include "bk_game.cuh"
#include "bk_rng.cuh"

struct BkTurn {
    uint32_t w0, w1, w2;   // legal set of the turn in progress, window form (warp-uniform); all zero = piece complete
    uint32_t m0, m1, m2;   // compact form only: window masks of this lane's surviving candidate (m0 carries pid << 27), 0 = none
    uint32_t tq;           // window-bit indices (7 bits each) of the turn's tiles AFTER the first (always the centre), newest in the low bits
    int tr, tc;            // T[0]
    int nT;                // |T|
    int last;              // the last tile laid (only kept when a digest per ply is asked for)
};

// window form of board rows (inverse of bk_window_to_row)
__device__ __forceinline__ void bk_rows_to_window(uint32_t row, int tr, int tc, int lane, uint32_t& W0, uint32_t& W1,
                                                  uint32_t& W2) {
    const uint32_t s = ((row << 4) >> tc) & 0x1FFu;
    const int wr = lane - tr + 4;
    const bool inw = (wr >= 0) && (wr < 9) && (lane < 20);
    const int wk = inw ? wr / 3 : 3;
    const int sh = inw ? 9 * (wr - 3 * wk) : 0;
    W0 = __reduce_or_sync(BK_FULL, wk == 0 ? s << sh : 0u);
    W1 = __reduce_or_sync(BK_FULL, wk == 1 ? s << sh : 0u);
    W2 = __reduce_or_sync(BK_FULL, wk == 2 ? s << sh : 0u);
}

// board tile of window bit wbit (0..80) for the window centred on (tr, tc)
__device__ __forceinline__ int bk_window_tile(int wbit, int tr, int tc) {
    const int wr = (wbit * 57) >> 9;                       // wbit / 9 for wbit < 81
    return (tr - 4) * 20 + (tc - 4) + wbit + 11 * wr;      // (tr-4+wr)*20 + (tc-4+wbit-9*wr)
}
__device__ __forceinline__ int bk_tile_window_bit(int t, int tr, int tc) {
    return (t / 20 - tr + 4) * 9 + (t % 20 - tc + 4);
}

// Chunk form -> compact form (bk_narrow_compact) with a shortcut: when no lane holds more than one survivor — the usual
// case once a tile or two are laid — every lane simply keeps its own, and the scan / shared-memory exchange is skipped.
// (The compact form does not ask for the survivors to sit on the low lanes.)
__device__ __forceinline__ void bk_turn_compact(BkRegs& G, int lane, const BkTabs& tabs) {
    const int cnt = __popc(G.smask);
    if (__any_sync(BK_FULL, cnt > 1)) { bk_narrow_compact(G, lane, tabs); return; }
    G.smask = cnt ? uint32_t((__ffs(G.smask) - 1) * 32 + lane) : BK_CAND_NONE;     // (at most 32 survivors: one per lane)
    G.alive = BK_NARROW_COMPACT;
}

__device__ __forceinline__ void bk_turn_load_masks(const BkRegs& G, BkTurn& T, const BkTabs& tabs) {
    T.m0 = T.m1 = T.m2 = 0u;
    if ((G.alive & BK_NARROW_COMPACT) && G.smask != BK_CAND_NONE) {
        T.m0 = tabs.w0[G.smask]; T.m1 = tabs.w1[G.smask]; T.m2 = tabs.w2[G.smask];
    }
}

// piece id of the unique surviving placement (called once per turn, when the legal set has become empty)
__device__ __forceinline__ int bk_turn_piece(const BkRegs& G, const BkTurn& T, int lane, const BkTabs& tabs) {
    uint32_t f = 0u;
    if (G.alive & BK_NARROW_COMPACT) {
        if (T.m0 | T.m1 | T.m2) f = (T.m0 >> 27) + 1u;
    } else if (G.smask) {
        f = (tabs.w0[(__ffs(G.smask) - 1) * 32 + lane] >> 27) + 1u;
    }
    return int(__reduce_max_sync(BK_FULL, f)) - 1;
}

// First tile (tr, tc) of a turn; free_/anch are the mover's turn-start rows.  Same candidate scan as
// bk_narrow_first, without the per-candidate piece bookkeeping.
__device__ __forceinline__ void bk_turn_first(BkRegs& G, BkTurn& T, uint32_t free_, uint32_t anch, uint32_t pieces,
                                              int tr, int tc, int lane, const BkTabs& tabs) {
    uint32_t FW0, FW1, FW2, AW0, AW1, AW2;
    bk_rows_to_window(free_, tr, tc, lane, FW0, FW1, FW2);
    bk_rows_to_window(anch, tr, tc, lane, AW0, AW1, AW2);
    const uint32_t NF0 = ~FW0 & 0x7FFFFFFu, NF1 = ~FW1, NF2 = ~FW2;
    uint32_t L0 = 0u, L1 = 0u, L2 = 0u, smask = 0u;
#pragma unroll
    for (int ch = 0; ch < BK_NUM_CAND_CHUNKS; ++ch) {
        // warp-uniform skip: no piece of this chunk is held (a branch-free scan of all 13 chunks measured 7 % slower)
        if ((c_cand_chunk_pieces[ch] & pieces) == 0u) continue;
        const int idx = ch * 32 + lane;
        const uint32_t w0 = tabs.w0[idx], m1 = tabs.w1[idx], m2 = tabs.w2[idx];
        const bool fits = ((w0 & NF0) | (m1 & NF1) | (m2 & NF2)) == 0u;
        const bool hits = ((w0 & AW0) | (m1 & AW1) | (m2 & AW2)) != 0u;   // AW0 has no bits above 26
        if (((pieces >> (w0 >> 27)) & 1u) && fits && hits) {
            L0 |= w0; L1 |= m1; L2 |= m2;
            smask |= 1u << ch;
        }
    }
    const uint32_t TW1 = 1u << 13;   // (tr, tc) is the window centre: bit 4*9+4 = 40 = word 1, bit 13
    T.w0 = __reduce_or_sync(BK_FULL, L0) & 0x7FFFFFFu;
    T.w1 = __reduce_or_sync(BK_FULL, L1) & ~TW1;
    T.w2 = __reduce_or_sync(BK_FULL, L2);
    T.tr = tr; T.tc = tc; T.nT = 1; T.tq = 0u;
    T.m0 = T.m1 = T.m2 = 0u;
    G.smask = smask;
    G.tw0 = 0u; G.tw1 = TW1; G.tw2 = 0u;
    if ((T.w0 | T.w1 | T.w2) != 0u) {                    // the turn goes on: later tiles re-test survivors only
        G.alive = __reduce_or_sync(BK_FULL, smask);
        bk_turn_compact(G, lane, tabs);
        bk_turn_load_masks(G, T, tabs);
    } else {
        G.alive = 0u;                                    // chunk form; bk_turn_piece reads smask
    }
}

// Later tile of a turn: window bit b of window word k.
__device__ __forceinline__ void bk_turn_next(BkRegs& G, BkTurn& T, int k, uint32_t b, int lane, const BkTabs& tabs) {
    if (k == 0) G.tw0 |= b; else if (k == 1) G.tw1 |= b; else G.tw2 |= b;
    if (G.alive & BK_NARROW_COMPACT) {
        const uint32_t mk = k == 0 ? T.m0 : (k == 1 ? T.m1 : T.m2);
        if ((mk & b) == 0u) { T.m0 = 0u; T.m1 = 0u; T.m2 = 0u; G.smask = BK_CAND_NONE; }
        T.w0 = __reduce_or_sync(BK_FULL, T.m0) & 0x7FFFFFFu & ~G.tw0;
        T.w1 = __reduce_or_sync(BK_FULL, T.m1) & ~G.tw1;
        T.w2 = __reduce_or_sync(BK_FULL, T.m2) & ~G.tw2;
        return;
    }
    const uint32_t* __restrict__ wk = k == 0 ? tabs.w0 : (k == 1 ? tabs.w1 : tabs.w2);
    uint32_t L0 = 0u, L1 = 0u, L2 = 0u;
    uint32_t smask = G.smask;
    for (uint32_t cm = G.alive; cm; cm &= cm - 1u) {   // warp-uniform
        const int ch = __ffs(cm) - 1;
        const int idx = ch * 32 + lane;
        if (((smask >> ch) & 1u) && (wk[idx] & b)) {
            L0 |= tabs.w0[idx]; L1 |= tabs.w1[idx]; L2 |= tabs.w2[idx];
        } else {
            smask &= ~(1u << ch);
        }
    }
    T.w0 = __reduce_or_sync(BK_FULL, L0) & 0x7FFFFFFu & ~G.tw0;
    T.w1 = __reduce_or_sync(BK_FULL, L1) & ~G.tw1;
    T.w2 = __reduce_or_sync(BK_FULL, L2) & ~G.tw2;
    G.smask = smask;
    if ((T.w0 | T.w1 | T.w2) != 0u) {
        G.alive = __reduce_or_sync(BK_FULL, smask);
        bk_turn_compact(G, lane, tabs);
        bk_turn_load_masks(G, T, tabs);
    }
}

// idx-th (ascending board order = window row-major order) tile of the turn's legal set: window word and bit
__device__ __forceinline__ void bk_turn_pick(const BkTurn& T, int idx, int c0, int c1, int& k, uint32_t& b, int& wbit) {
    uint32_t w;
    if (idx < c0) { w = T.w0; k = 0; }
    else if (idx < c0 + c1) { w = T.w1; k = 1; idx -= c0; }
    else { w = T.w2; k = 2; idx -= c0 + c1; }
#pragma unroll 1                                         // the sets are small (a handful of tiles): no unrolled copies
    for (int i = 0; i < idx; ++i) w &= w - 1u;
    b = w & (0u - w);
    wbit = 27 * k + __ffs(w) - 1;
}

// BkState's view of a turn in progress (bk_apply_t's bookkeeping): the mover's squares include T, the cached
// legal rows are the narrowed set, |T| and the tiles are recorded.
__device__ __forceinline__ void bk_turn_materialise(BkRegs& G, const BkTurn& T, int lane) {
    const int p = bk_cur(G);
    const uint32_t rows = bk_window_to_row(G.tw0, G.tw1, G.tw2, T.tr, T.tc, lane);
    if (p == 0) G.o0 |= rows; else if (p == 1) G.o1 |= rows; else if (p == 2) G.o2 |= rows; else G.o3 |= rows;
    G.legal = bk_window_to_row(T.w0, T.w1, T.w2, T.tr, T.tc, lane);
    G.meta = (G.meta & ~(7u << 6)) | (uint32_t(T.nT) << 6);
    uint32_t t[4] = {uint32_t(T.tr * 20 + T.tc), 0u, 0u, 0u};
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < T.nT) t[i] = uint32_t(bk_window_tile(int((T.tq >> (7 * (T.nT - 1 - i))) & 127u), T.tr, T.tc));
    G.t01 = t[0] | (t[1] << 16);
    G.t23 = t[2] | (t[3] << 16);
}

// The turn's tiles go to the move history together: lane i writes the i-th tile of the turn (one store per piece
// instead of one per tile).  ply = the game's ply count, the turn's tiles included.
__device__ __forceinline__ void bk_turn_record(const BkTurn& T, uint32_t ply, int p, int lane, uint16_t* __restrict__ h16) {
    const int sh = 7 * (T.nT - 1 - lane);
    const int wbit = lane == 0 ? 40 : int((T.tq >> (sh & 31)) & 127u);
    const uint32_t at = ply - uint32_t(T.nT) + uint32_t(lane);
    if (lane < T.nT && at < BK_HIST_CAP) h16[at] = uint16_t(bk_window_tile(wbit, T.tr, T.tc) | (p << 9));
}

// the reverse: pick up a stored state whose turn is in progress
__device__ __forceinline__ void bk_turn_resume(const BkRegs& G, BkTurn& T, int lane, const BkTabs& tabs) {
    T.nT = int((G.meta >> 6) & 7u);
    T.w0 = T.w1 = T.w2 = 0u; T.m0 = T.m1 = T.m2 = 0u; T.tq = 0u; T.tr = T.tc = 0; T.last = 0;
    if (T.nT == 0) return;
    const int t0 = int(G.t01 & 0xFFFFu);
    T.tr = t0 / 20; T.tc = t0 % 20;
    bk_rows_to_window(G.legal, T.tr, T.tc, lane, T.w0, T.w1, T.w2);
    const uint32_t ts[4] = {G.t01 & 0xFFFFu, G.t01 >> 16, G.t23 & 0xFFFFu, G.t23 >> 16};
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < T.nT) T.tq = (T.tq << 7) | uint32_t(bk_tile_window_bit(int(ts[i]), T.tr, T.tc));
    bk_turn_load_masks(G, T, tabs);
}
