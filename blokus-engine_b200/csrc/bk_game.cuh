// bk_game.cuh — warp-per-game Blokus rules engine for sm_100a.
//
// One warp owns one game.  Lane r (0..19) holds row r of every 20x20 bitboard as a uint32 with
// 20 live bits (bit c = column c); lanes 20..31 hold zeros.  Vertical neighbours come from
// __shfl_up/down, horizontal ones from shifts; everything is integer ALU work on registers.
//
// What is restated (behaviour only — the reference is hash maps of placements on a byte board):
//   Board::place_tile / is_valid_move      blokus/src/board.rs:62-141
//   get_piece_moves/get_moves/get_tile_moves  blokus/src/game.rs:12-74
//   Game::apply / advance_player            blokus/src/game.rs:150-223
//   get_scores / get_payoff                 blokus/src/board.rs:155-181, game.rs:252-272
//
// Closed forms used, with own_p = player p's squares:
//   restricted_p = occupied | adj4(own_p)            (byte bit 1<<(4+p) or an occupied cell)
//   anchors_p    = (diag4(own_p) | start_p) & ~restricted_p
// A turn is a sequence of single-tile applies; the set of placements still consistent with the
// tiles T laid this turn is recomputed from the TURN-START boards (own_p minus T) every time:
//   S = { valid placements at turn start that contain T },  legal = union(cells(S)) \ T,
// which is exactly what game.rs:156-173 maintains by set intersection.  The piece is committed
// when legal becomes empty; then cells(S) == T, so the piece length is |T|.
#pragma once
#include <stdint.h>

#include "bk_tables_gen.h"

#define BK_FULL 0xffffffffu
#define BK_ROWMASK 0xFFFFFu
#define BK_HIST_CAP 360

// Per-game record in HBM.  Lane r reads own[r] as one 16-byte vector: a warp's 20 loads are one
// contiguous 320-byte segment.
struct __align__(16) BkState {
    uint32_t own[20][4];  // [row][player]
    uint32_t legal[20];   // current legal-tile set (narrowed mid-piece), cached
    uint32_t pieces[4];   // remaining piece-id masks
    uint32_t meta;        // bits 0-1 current player, 2-5 eliminated mask, 6-8 |T|
    uint32_t lastlens;    // byte p = last_piece_lens[p]            (game.rs:98)
    uint32_t t01, t23;    // tiles laid this turn, 16 bits each (at most 4 are ever stored)
    uint32_t ply;         // history length
    uint32_t pad[3];      // [0],[1]: child block offset / count when the record is an MCTS node
    // Narrowing cache of the turn in progress (derived data; meaningful only while |T| >= 1).  Two forms:
    //   chunk form   — alive = chunks of the candidate table with a surviving placement; smask[l] bit ch set
    //                  iff candidate ch*32+l is still consistent with T;
    //   compact form — (alive bit 31 set) at most 32 survivors, smask[l] = candidate index held by lane l or
    //                  0xFFFF: later tiles of the turn are then one pass, not one pass per alive chunk.
    uint32_t alive;
    uint32_t tw[3];       // 9x9-window mask of T round T[0]
    uint16_t smask[32];
};
#define BK_NARROW_COMPACT 0x80000000u
#define BK_CAND_NONE 0xFFFFu
static_assert(sizeof(BkState) == 528, "BkState layout");

struct BkRegs {
    uint32_t o0, o1, o2, o3, legal;
    uint32_t pc0, pc1, pc2, pc3;
    uint32_t meta, lastlens, t01, t23, ply;
    uint32_t smask, alive, tw0, tw1, tw2;   // narrowing cache (see BkState)
};

struct BkTabs {  // candidate table staged in shared memory
    const uint32_t* w0;
    const uint32_t* w1;
    const uint32_t* w2;
    uint16_t* scratch;  // 32 entries private to this warp (survivor compaction)
};
#ifndef BK_MAX_WARPS_PER_CTA
#define BK_MAX_WARPS_PER_CTA 4
#endif

struct BkCounters {  // lane-local partial sums, reduced once per kernel
    uint32_t movegens;  // counted on lane 0 only
    uint32_t crem;      // lane l < 21: sum over movegens of cells(piece l) if piece l was held
};

static __device__ const uint32_t g_cand_w0[BK_NUM_CANDS_PAD] = BK_CAND_W0_INIT;
static __device__ const uint32_t g_cand_w1[BK_NUM_CANDS_PAD] = BK_CAND_W1_INIT;
static __device__ const uint32_t g_cand_w2[BK_NUM_CANDS_PAD] = BK_CAND_W2_INIT;
static __constant__ uint32_t c_cand_chunk_pieces[BK_NUM_CAND_CHUNKS] = BK_CAND_CHUNK_PIECES_INIT;
static __constant__ uint8_t c_piece_points[BK_NUM_PIECES] = BK_PIECE_POINTS_INIT;
static __constant__ uint8_t c_piece_first_variant[BK_NUM_PIECES + 1] = BK_PIECE_FIRST_VARIANT_INIT;
static __constant__ uint8_t c_variant_width[BK_NUM_VARIANTS] = BK_VARIANT_WIDTH_INIT;
static __constant__ uint8_t c_variant_height[BK_NUM_VARIANTS] = BK_VARIANT_HEIGHT_INIT;
static __constant__ uint8_t c_variant_ncells[BK_NUM_VARIANTS] = BK_VARIANT_NCELLS_INIT;
static __constant__ uint16_t c_variant_offsets[BK_NUM_VARIANTS][5] = BK_VARIANT_OFFSETS_INIT;

#define BK_TABS_SMEM_WORDS (3 * BK_NUM_CANDS_PAD + 16 * BK_MAX_WARPS_PER_CTA)

// All threads of the CTA call this once; the caller must __syncthreads() afterwards.
__device__ __forceinline__ BkTabs bk_stage_tables(uint32_t* smem) {
    for (int i = threadIdx.x; i < BK_NUM_CANDS_PAD; i += blockDim.x) {
        smem[i] = g_cand_w0[i];
        smem[BK_NUM_CANDS_PAD + i] = g_cand_w1[i];
        smem[2 * BK_NUM_CANDS_PAD + i] = g_cand_w2[i];
    }
    BkTabs t;
    t.w0 = smem;
    t.w1 = smem + BK_NUM_CANDS_PAD;
    t.w2 = smem + 2 * BK_NUM_CANDS_PAD;
    t.scratch = reinterpret_cast<uint16_t*>(smem + 3 * BK_NUM_CANDS_PAD) + 32 * (threadIdx.x >> 5);
    return t;
}

// The same tables read in place (global memory through L1): for kernels whose candidate scans are a small share of the
// work, the 5 KB per-CTA copy costs more shared memory (and L1 carve-out) than it saves.  `scratch`: 32 uint16 per warp.
__device__ __forceinline__ BkTabs bk_global_tables(uint16_t* scratch) {
    BkTabs t;
    t.w0 = g_cand_w0; t.w1 = g_cand_w1; t.w2 = g_cand_w2;
    t.scratch = scratch + 32 * (threadIdx.x >> 5);
    return t;
}

__device__ __forceinline__ uint32_t bk_sel4(int p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return p == 0 ? a : (p == 1 ? b : (p == 2 ? c : d));
}
__device__ __forceinline__ uint32_t bk_smear4(uint32_t x) { uint32_t t = x | (x << 1); return t | (t << 2); }
__device__ __forceinline__ uint32_t bk_smear5(uint32_t x) { uint32_t t = x | (x << 1); return t | (t << 2) | (x << 4); }

__device__ __forceinline__ void bk_load(const BkState* __restrict__ s, int lane, BkRegs& G) {
    if (lane < 20) {
        const uint4 v = reinterpret_cast<const uint4*>(s->own)[lane];
        G.o0 = v.x; G.o1 = v.y; G.o2 = v.z; G.o3 = v.w;
        G.legal = s->legal[lane];
    } else {
        G.o0 = G.o1 = G.o2 = G.o3 = G.legal = 0u;
    }
    const uint4 pc = *reinterpret_cast<const uint4*>(s->pieces);
    G.pc0 = pc.x; G.pc1 = pc.y; G.pc2 = pc.z; G.pc3 = pc.w;
    const uint4 m = *reinterpret_cast<const uint4*>(&s->meta);
    G.meta = m.x; G.lastlens = m.y; G.t01 = m.z; G.t23 = m.w;
    G.ply = s->ply;
    const uint4 cw = *reinterpret_cast<const uint4*>(&s->alive);
    G.alive = cw.x; G.tw0 = cw.y; G.tw1 = cw.z; G.tw2 = cw.w;
    G.smask = s->smask[lane];
}

__device__ __forceinline__ void bk_store(BkState* __restrict__ s, int lane, const BkRegs& G) {
    if (lane < 20) {
        reinterpret_cast<uint4*>(s->own)[lane] = make_uint4(G.o0, G.o1, G.o2, G.o3);
        s->legal[lane] = G.legal;
    }
    if (lane == 0) {
        *reinterpret_cast<uint4*>(s->pieces) = make_uint4(G.pc0, G.pc1, G.pc2, G.pc3);
        *reinterpret_cast<uint4*>(&s->meta) = make_uint4(G.meta, G.lastlens, G.t01, G.t23);
        s->ply = G.ply;
        *reinterpret_cast<uint4*>(&s->alive) = make_uint4(G.alive, G.tw0, G.tw1, G.tw2);
    }
    s->smask[lane] = uint16_t(G.smask);
}

__device__ __forceinline__ int bk_warp_sum_i(int v) { return int(__reduce_add_sync(BK_FULL, unsigned(v))); }
__device__ __forceinline__ int bk_cur(const BkRegs& G) { return int(G.meta & 3u); }
__device__ __forceinline__ uint32_t bk_elim(const BkRegs& G) { return (G.meta >> 2) & 0xFu; }
__device__ __forceinline__ bool bk_terminal(const BkRegs& G) { return bk_elim(G) == 0xFu; }

// free_p / anchors_p rows of this lane for player p (see header).
__device__ __forceinline__ void bk_free_anchor(uint32_t mine, uint32_t occ, int p, int lane, uint32_t& free_,
                                               uint32_t& anch) {
    uint32_t up = __shfl_up_sync(BK_FULL, mine, 1);
    uint32_t dn = __shfl_down_sync(BK_FULL, mine, 1);
    if (lane == 0) up = 0u;
    if (lane >= 19) dn = 0u;
    const uint32_t ud = up | dn;
    const uint32_t adj = (mine << 1) | (mine >> 1) | ud;
    const uint32_t diag = (ud << 1) | (ud >> 1);
    // board.rs:44-53: start corners 0, 19, 399, 380 for players 0..3
    uint32_t start = 0u;
    if (lane == ((p >> 1) ? 19 : 0)) start = (p == 1 || p == 2) ? (1u << 19) : 1u;
    const uint32_t rowmask = lane < 20 ? BK_ROWMASK : 0u;
    free_ = ~(occ | adj) & rowmask;
    anch = (diag | start) & free_;
}

// Turn-start legal-tile board of a player: union of the cells of every valid placement of every
// remaining piece (get_tile_moves keys, game.rs:60-74).  Fully unrolled over the 91 variants.
static __device__ __noinline__ uint32_t bk_movegen_rows(uint32_t bk_free, uint32_t bk_anch, uint32_t bk_pieces, int lane) {
#include "bk_movegen_gen.inc"
    return bk_legal & BK_ROWMASK;
}

// free_/anch: the mover's turn-start rows, handed back because the first tile of the turn narrows against them
__device__ __forceinline__ uint32_t bk_movegen_start(const BkRegs& G, int p, int lane, BkCounters& ctr, uint32_t& free_,
                                                     uint32_t& anch) {
    const uint32_t mine = bk_sel4(p, G.o0, G.o1, G.o2, G.o3);
    const uint32_t occ = G.o0 | G.o1 | G.o2 | G.o3;
    const uint32_t pieces = bk_sel4(p, G.pc0, G.pc1, G.pc2, G.pc3);
    bk_free_anchor(mine, occ, p, lane, free_, anch);
    if (lane == 0) ctr.movegens += 1u;
    if (lane < BK_NUM_PIECES && ((pieces >> lane) & 1u))
        ctr.crem += uint32_t(c_piece_points[lane]) * uint32_t(c_piece_first_variant[lane + 1] - c_piece_first_variant[lane]);
    if (pieces == 0u || !__any_sync(BK_FULL, anch != 0u)) return 0u;
    return bk_movegen_rows(free_, anch, pieces, lane);
}
__device__ __forceinline__ uint32_t bk_movegen_start(const BkRegs& G, int p, int lane, BkCounters& ctr) {
    uint32_t free_, anch;
    return bk_movegen_start(G, p, lane, ctr, free_, anch);
}

struct BkNarrow {
    uint32_t w0, w1, w2;  // the narrowed legal set inside the 9x9 window centred on T[0] = (tr, tc): warp-uniform
    int tr, tc;
    int pid;              // piece id of a surviving placement (the committed piece when the set is empty)
    bool any_valid;       // some turn-start placement contains T (always true after a legal tile)
};

// window word index / bit of board tile t relative to the window centred on (tr, tc)
__device__ __forceinline__ void bk_window_bit(int t, int tr, int tc, int& k, uint32_t& b) {
    const int bit = (t / 20 - tr + 4) * 9 + (t % 20 - tc + 4);
    k = bit / 27;
    b = 1u << (bit - 27 * k);
}

// scatter the 9x9 window legal mask back to this lane's board row
__device__ __forceinline__ uint32_t bk_window_to_row(uint32_t L0, uint32_t L1, uint32_t L2, int tr, int tc, int lane) {
    const int wr = lane - tr + 4;
    if (wr < 0 || wr >= 9 || lane >= 20) return 0u;
    const int wk = wr / 3;
    const uint32_t Lw = wk == 0 ? L0 : (wk == 1 ? L1 : L2);
    const uint32_t slice = (Lw >> (9 * (wr - 3 * wk))) & 0x1FFu;
    return ((slice << tc) >> 4) & BK_ROWMASK;
}

// Chunk form -> compact form when at most 32 placements survive (deterministic: ascending (lane, chunk) order).
__device__ __forceinline__ void bk_narrow_compact(BkRegs& G, int lane, const BkTabs& tabs) {
    const int cnt = __popc(G.smask);
    const int total = bk_warp_sum_i(cnt);
    if (total > 32 || total == 0) return;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(BK_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    int pos = incl - cnt;
    for (uint32_t m = G.smask; m; m &= m - 1u) tabs.scratch[pos++] = uint16_t((__ffs(m) - 1) * 32 + lane);
    __syncwarp();
    G.smask = lane < total ? uint32_t(tabs.scratch[lane]) : BK_CAND_NONE;
    __syncwarp();
    G.alive = BK_NARROW_COMPACT;
}

// First tile t of a turn: S = { turn-start-valid placements containing t }, evaluated inside the 9x9
// window centred on t (every (variant, cell) candidate contains t by construction).  free_/anch are the
// TURN-START rows of this lane.  Leaves the survivors in G's narrowing cache.  The 13 chunks are
// independent, so the loop is fully unrolled: a lone warp (MCTS runs ~2 warps per scheduler) needs the ILP.
__device__ __forceinline__ BkNarrow bk_narrow_first(BkRegs& G, uint32_t free_, uint32_t anch, uint32_t pieces, int t,
                                                    int lane, const BkTabs& tabs) {
    const int tr = t / 20, tc = t % 20;
    const uint32_t fs = ((free_ << 4) >> tc) & 0x1FFu;
    const uint32_t as = ((anch << 4) >> tc) & 0x1FFu;
    const int wr = lane - tr + 4;
    const bool inw = (wr >= 0) && (wr < 9) && (lane < 20);
    const int wk = inw ? wr / 3 : 3;
    const int sh = inw ? 9 * (wr - 3 * wk) : 0;
    const uint32_t FW0 = __reduce_or_sync(BK_FULL, wk == 0 ? fs << sh : 0u);
    const uint32_t FW1 = __reduce_or_sync(BK_FULL, wk == 1 ? fs << sh : 0u);
    const uint32_t FW2 = __reduce_or_sync(BK_FULL, wk == 2 ? fs << sh : 0u);
    const uint32_t AW0 = __reduce_or_sync(BK_FULL, wk == 0 ? as << sh : 0u);
    const uint32_t AW1 = __reduce_or_sync(BK_FULL, wk == 1 ? as << sh : 0u);
    const uint32_t AW2 = __reduce_or_sync(BK_FULL, wk == 2 ? as << sh : 0u);
    uint32_t L0 = 0u, L1 = 0u, L2 = 0u, smask = 0u;
    int found = 0;
#pragma unroll      // (a rolled loop measured 7 % slower in k_playout as well: 0.689 vs 0.647 ms)
    for (int ch = 0; ch < BK_NUM_CAND_CHUNKS; ++ch) {
        if ((c_cand_chunk_pieces[ch] & pieces) == 0u) continue;  // warp-uniform: no piece of this chunk is held
        const int idx = ch * 32 + lane;
        const uint32_t w0 = tabs.w0[idx];
        const uint32_t m0 = w0 & 0x7FFFFFFu, m1 = tabs.w1[idx], m2 = tabs.w2[idx];
        const uint32_t pid = w0 >> 27;
        const bool fits = ((m0 & ~FW0) | (m1 & ~FW1) | (m2 & ~FW2)) == 0u;
        const bool hits = ((m0 & AW0) | (m1 & AW1) | (m2 & AW2)) != 0u;
        if (((pieces >> pid) & 1u) && fits && hits) {
            L0 |= m0; L1 |= m1; L2 |= m2;
            found = int(pid) + 1;
            smask |= 1u << ch;
        }
    }
    const uint32_t TW1 = 1u << 13;   // t is the window centre: bit 4*9+4 = 40 = word 1, bit 13
    L0 = __reduce_or_sync(BK_FULL, L0);
    L1 = __reduce_or_sync(BK_FULL, L1) & ~TW1;
    L2 = __reduce_or_sync(BK_FULL, L2);
    found = int(__reduce_max_sync(BK_FULL, unsigned(found)));
    G.smask = smask;
    G.alive = __reduce_or_sync(BK_FULL, smask);
    G.tw0 = 0u; G.tw1 = TW1; G.tw2 = 0u;
    BkNarrow out;
    out.pid = found - 1;
    out.any_valid = found > 0;
    out.w0 = L0; out.w1 = L1; out.w2 = L2; out.tr = tr; out.tc = tc;
    if ((L0 | L1 | L2) != 0u) bk_narrow_compact(G, lane, tabs);   // the turn goes on: later tiles re-test survivors
    return out;
}

// Later tiles of a turn: S only shrinks (game.rs:165-173 intersects with the placements through the new
// tile), so only the cached survivors are re-tested, and only against the new tile t.  t0 = T[0].
__device__ __forceinline__ BkNarrow bk_narrow_next(BkRegs& G, int t0, int t, int lane, const BkTabs& tabs) {
    const int tr = t0 / 20, tc = t0 % 20;
    int k; uint32_t b;
    bk_window_bit(t, tr, tc, k, b);
    if (k == 0) G.tw0 |= b; else if (k == 1) G.tw1 |= b; else G.tw2 |= b;
    uint32_t L0 = 0u, L1 = 0u, L2 = 0u;
    int found = 0;
    if (G.alive & BK_NARROW_COMPACT) {                  // one survivor per lane
        const uint32_t idx = G.smask;
        if (idx != BK_CAND_NONE) {
            const uint32_t w0 = tabs.w0[idx], w1 = tabs.w1[idx], w2 = tabs.w2[idx];
            if ((k == 0 ? w0 : (k == 1 ? w1 : w2)) & b) {
                L0 = w0 & 0x7FFFFFFu; L1 = w1; L2 = w2;
                found = int(w0 >> 27) + 1;
            } else {
                G.smask = BK_CAND_NONE;
            }
        }
        L0 = __reduce_or_sync(BK_FULL, L0) & ~G.tw0;
        L1 = __reduce_or_sync(BK_FULL, L1) & ~G.tw1;
        L2 = __reduce_or_sync(BK_FULL, L2) & ~G.tw2;
        found = int(__reduce_max_sync(BK_FULL, unsigned(found)));
        BkNarrow out;
        out.pid = found - 1;
        out.any_valid = found > 0;
        out.w0 = L0; out.w1 = L1; out.w2 = L2; out.tr = tr; out.tc = tc;
        return out;
    }
    const uint32_t* __restrict__ wk = k == 0 ? tabs.w0 : (k == 1 ? tabs.w1 : tabs.w2);
    uint32_t smask = G.smask;
    for (uint32_t cm = G.alive; cm; cm &= cm - 1u) {   // warp-uniform
        const int ch = __ffs(cm) - 1;
        const int idx = ch * 32 + lane;
        if (((smask >> ch) & 1u) && (wk[idx] & b)) {
            const uint32_t w0 = tabs.w0[idx];
            L0 |= w0 & 0x7FFFFFFu; L1 |= tabs.w1[idx]; L2 |= tabs.w2[idx];
            found = int(w0 >> 27) + 1;
        } else {
            smask &= ~(1u << ch);
        }
    }
    L0 = __reduce_or_sync(BK_FULL, L0) & ~G.tw0;
    L1 = __reduce_or_sync(BK_FULL, L1) & ~G.tw1;
    L2 = __reduce_or_sync(BK_FULL, L2) & ~G.tw2;
    found = int(__reduce_max_sync(BK_FULL, unsigned(found)));
    G.smask = smask;
    G.alive = __reduce_or_sync(BK_FULL, smask);
    BkNarrow out;
    out.pid = found - 1;
    out.any_valid = found > 0;
    out.w0 = L0; out.w1 = L1; out.w2 = L2; out.tr = tr; out.tc = tc;
    if ((L0 | L1 | L2) != 0u) bk_narrow_compact(G, lane, tabs);
    return out;
}

// position of the k-th (0-based) set bit of a word whose live bits are below bit 32; branch-free
// binary search on popcounts (the __fns intrinsic is a long software loop)
__device__ __forceinline__ int bk_kth_set_bit(uint32_t m, int k) {
    int pos = 0, c;
    c = __popc(m & 0xFFFFu); if (k >= c) { k -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (k >= c) { k -= c; pos += 8;  m >>= 8; }
    c = __popc(m & 0xFu);    if (k >= c) { k -= c; pos += 4;  m >>= 4; }
    c = __popc(m & 0x3u);    if (k >= c) { k -= c; pos += 2;  m >>= 2; }
    c = int(m & 1u);         if (k >= c) { pos += 1; }
    return pos;
}

__device__ __forceinline__ int bk_nth_set_bit(uint32_t mask, int n) {  // n-th (0-based) set bit, -1 if none
    if (n < 0 || n >= __popc(mask)) return -1;
    return bk_kth_set_bit(mask, n);
}

// Game::advance_player (game.rs:203-223) as a loop: next seat that is not eliminated and has a
// legal tile; seats found blocked are eliminated for good.  Move generation is skipped for seats
// already eliminated — their set is provably empty (boards only fill up), so the result is the same.
__device__ __forceinline__ void bk_advance(BkRegs& G, int lane, BkCounters& ctr, uint32_t& free_, uint32_t& anch) {
    uint32_t elim = bk_elim(G);
    int cur = bk_cur(G);
    uint32_t legal = 0u;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        if (elim == 0xFu) break;
        cur = (cur + 1) & 3;
        if ((elim >> cur) & 1u) continue;
        const uint32_t lg = bk_movegen_start(G, cur, lane, ctr, free_, anch);
        if (!__any_sync(BK_FULL, lg != 0u)) { elim |= 1u << cur; continue; }
        legal = lg;
        break;
    }
    G.legal = legal;
    G.meta = (G.meta & ~0x3Fu) | uint32_t(cur) | (elim << 2);
}
__device__ __forceinline__ void bk_advance(BkRegs& G, int lane, BkCounters& ctr) {
    uint32_t free_, anch;
    bk_advance(G, lane, ctr, free_, anch);
}

// Game::apply(tile, piece_to_finish) (game.rs:150-194).  finish < 0 is None.  Returns false (and
// leaves the game untouched) when the tile is not legal or finish is out of range.
// this lane's row of a narrowed legal set
__device__ __forceinline__ uint32_t bk_narrow_row(const BkNarrow& nw, int lane) {
    return bk_window_to_row(nw.w0, nw.w1, nw.w2, nw.tr, nw.tc, lane);
}

// TRUSTED_LAZY (the persistent playout only): the tile is known to come from the legal set, so the membership vote
// is skipped, and while a turn is in progress G.legal is NOT refreshed — the caller works on the window form
// returned through nw_out and materialises the rows (bk_narrow_row) when something needs them.
template <bool TRUSTED_LAZY>
__device__ __forceinline__ bool bk_apply_t(BkRegs& G, int tile, int finish, int lane, const BkTabs& tabs,
                                           BkCounters& ctr, BkNarrow* nw_out) {
    if (tile < 0 || tile >= 400 || bk_terminal(G)) return false;
    const int p = bk_cur(G);
    const int tr = tile / 20, tc = tile % 20;
    const uint32_t bit = (lane == tr) ? (1u << tc) : 0u;
    if (!TRUSTED_LAZY && !__any_sync(BK_FULL, (G.legal & bit) != 0u)) return false;
    const uint32_t pieces = bk_sel4(p, G.pc0, G.pc1, G.pc2, G.pc3);
    int fin_pid = -1;
    if (finish >= 0) {
        fin_pid = bk_nth_set_bit(pieces, finish);
        if (fin_pid < 0) return false;
    }
    const int nT = int((G.meta >> 6) & 7u) + 1;   // tiles laid this turn, this one included
    const int t0 = nT > 1 ? int(G.t01 & 0xFFFFu) : tile;
    // Board::place_tile (board.rs:95-141): the square joins own_p; restricted/anchor sets are
    // derived from the bitboards on demand.
    if (p == 0) G.o0 |= bit; else if (p == 1) G.o1 |= bit; else if (p == 2) G.o2 |= bit; else G.o3 |= bit;
    G.ply += 1u;
    BkNarrow nw;
    if (nT > 1) {
        nw = bk_narrow_next(G, t0, tile, lane, tabs);
    } else {
        // turn-start boards = current boards minus the tile just laid
        const uint32_t mine0 = bk_sel4(p, G.o0, G.o1, G.o2, G.o3) & ~bit;
        const uint32_t occ0 = (G.o0 | G.o1 | G.o2 | G.o3) & ~bit;
        uint32_t free_, anch;
        bk_free_anchor(mine0, occ0, p, lane, free_, anch);
        nw = bk_narrow_first(G, free_, anch, pieces, tile, lane, tabs);
    }
    const bool done = (nw.w0 | nw.w1 | nw.w2) == 0u;      // warp-uniform: no vote needed
    if (done || fin_pid >= 0) {
        // game.rs:176-191: commit the piece, remember its size, pass the turn
        const int pid = fin_pid >= 0 ? fin_pid : nw.pid;
        const uint32_t len = fin_pid >= 0 ? uint32_t(c_piece_points[pid]) : uint32_t(nT);
        const uint32_t clr = ~(1u << pid);
        if (p == 0) G.pc0 &= clr; else if (p == 1) G.pc1 &= clr; else if (p == 2) G.pc2 &= clr; else G.pc3 &= clr;
        G.lastlens = (G.lastlens & ~(0xFFu << (8 * p))) | (len << (8 * p));
        G.meta &= ~(7u << 6);
        G.t01 = 0u; G.t23 = 0u;
        bk_advance(G, lane, ctr);
    } else {
        if (!TRUSTED_LAZY) G.legal = bk_narrow_row(nw, lane);
        G.meta = (G.meta & ~(7u << 6)) | (uint32_t(nT) << 6);
        if (nT == 1) G.t01 = uint32_t(tile);
        else if (nT == 2) G.t01 |= uint32_t(tile) << 16;
        else if (nT == 3) G.t23 = uint32_t(tile);
        else G.t23 |= uint32_t(tile) << 16;
    }
    if (nw_out) *nw_out = nw;
    return true;
}

__device__ __forceinline__ bool bk_apply(BkRegs& G, int tile, int finish, int lane, const BkTabs& tabs,
                                         BkCounters& ctr) {
    return bk_apply_t<false>(G, tile, finish, lane, tabs, ctr, nullptr);
}

// idx-th (0-based, ascending board order = window row-major order) tile of a narrowed legal set; warp-uniform
// arithmetic on the window words, no collectives.  The sets are small (a handful of tiles), so the k-th set bit
// is found by clearing the k lower ones.
__device__ __forceinline__ int bk_narrow_count(const BkNarrow& nw) { return __popc(nw.w0) + __popc(nw.w1) + __popc(nw.w2); }
__device__ __forceinline__ int bk_narrow_select(const BkNarrow& nw, int idx) {
    const int c0 = __popc(nw.w0), c1 = __popc(nw.w1);
    uint32_t w;
    int base;
    if (idx < c0) { w = nw.w0; base = 0; }
    else if (idx < c0 + c1) { w = nw.w1; base = 27; idx -= c0; }
    else { w = nw.w2; base = 54; idx -= c0 + c1; }
    for (int i = 0; i < idx; ++i) w &= w - 1u;
    const int bit = base + __ffs(w) - 1;
    const int wr = bit / 9, wc = bit - 9 * wr;
    return (nw.tr - 4 + wr) * 20 + (nw.tc - 4 + wc);
}

// Game::reset (game.rs:102-114)
__device__ __forceinline__ void bk_reset(BkRegs& G, int lane, BkCounters& ctr) {
    G.o0 = G.o1 = G.o2 = G.o3 = 0u;
    G.pc0 = G.pc1 = G.pc2 = G.pc3 = (1u << BK_NUM_PIECES) - 1u;
    G.meta = 0u; G.lastlens = 0u; G.t01 = 0u; G.t23 = 0u; G.ply = 0u;
    G.smask = 0u; G.alive = 0u; G.tw0 = G.tw1 = G.tw2 = 0u;
    G.legal = bk_movegen_start(G, 0, lane, ctr);
}

__device__ __forceinline__ int bk_warp_sum(int v) {
    return int(__reduce_add_sync(BK_FULL, unsigned(v)));
}

// Board::get_scores (board.rs:155-181)
__device__ __forceinline__ void bk_scores(const BkRegs& G, int (&sc)[4]) {
    sc[0] = bk_warp_sum(__popc(G.o0)); sc[1] = bk_warp_sum(__popc(G.o1));
    sc[2] = bk_warp_sum(__popc(G.o2)); sc[3] = bk_warp_sum(__popc(G.o3));
    const uint32_t pcs[4] = {G.pc0, G.pc1, G.pc2, G.pc3};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        sc[p] -= 89;
        if (pcs[p] == 0u) {
            sc[p] += 15;
            if (((G.lastlens >> (8 * p)) & 0xFFu) == 1u) sc[p] += 5;
        }
    }
}

// Game::get_payoff (game.rs:252-272)
__device__ __forceinline__ void bk_payoff(const BkRegs& G, float (&pay)[4]) {
    int sc[4];
    bk_scores(G, sc);
    int hi = sc[0];
#pragma unroll
    for (int p = 1; p < 4; ++p) hi = sc[p] > hi ? sc[p] : hi;
    int k = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) k += (sc[p] == hi) ? 1 : 0;
    const float share = __fdiv_rn(1.0f, float(k));
#pragma unroll
    for (int p = 0; p < 4; ++p) pay[p] = (sc[p] == hi) ? share : 0.0f;
}

__device__ __forceinline__ uint64_t bk_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Order-independent digest of the full game state (same word list as oracle/oracle_capi.cpp
// state_digest): own rows, legal rows, remaining-piece masks, last piece lengths, seat + eliminated.
__device__ __forceinline__ uint64_t bk_digest(const BkRegs& G, int lane) {
    uint64_t s = 0ull;
    if (lane < 20) {
        s += bk_splitmix64((uint64_t(0 * 32 + lane) << 32) | G.o0);
        s += bk_splitmix64((uint64_t(1 * 32 + lane) << 32) | G.o1);
        s += bk_splitmix64((uint64_t(2 * 32 + lane) << 32) | G.o2);
        s += bk_splitmix64((uint64_t(3 * 32 + lane) << 32) | G.o3);
        s += bk_splitmix64((uint64_t(4 * 32 + lane) << 32) | G.legal);
    }
    if (lane < 4) {
        s += bk_splitmix64((uint64_t(5 * 32 + lane) << 32) | bk_sel4(lane, G.pc0, G.pc1, G.pc2, G.pc3));
        s += bk_splitmix64((uint64_t(6 * 32 + lane) << 32) | ((G.lastlens >> (8 * lane)) & 0xFFu));
    }
    if (lane == 0) s += bk_splitmix64((uint64_t(7 * 32) << 32) | (uint32_t(bk_cur(G)) | (bk_elim(G) << 4)));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(BK_FULL, s, d);
    return s;
}

// number of legal tiles and the idx-th one in ascending tile order
__device__ __forceinline__ int bk_legal_count(uint32_t legal) { return bk_warp_sum(__popc(legal)); }

__device__ __forceinline__ void bk_legal_select_rc(uint32_t legal, int idx, int lane, int& r, int& c) {
    const int cnt = __popc(legal);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(BK_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int excl = incl - cnt;
    const unsigned who = __ballot_sync(BK_FULL, (idx >= excl) && (idx < incl));
    const int src = who ? (__ffs(who) - 1) : 0;                 // the row that holds the idx-th legal tile
    const uint32_t row = __shfl_sync(BK_FULL, legal, src);
    const int k = idx - __shfl_sync(BK_FULL, excl, src);
    // lane j asks: is column j the k-th set bit of that row?  (all lanes cooperate instead of one lane searching)
    const bool hit = ((row >> lane) & 1u) && (__popc(row & ((1u << lane) - 1u)) == k);
    const unsigned col = __ballot_sync(BK_FULL, hit);
    r = src;
    c = col ? (__ffs(col) - 1) : 0;
}

__device__ __forceinline__ int bk_legal_select(uint32_t legal, int idx, int lane) {
    int r, c;
    bk_legal_select_rc(legal, idx, lane, r, c);
    return r * 20 + c;
}
