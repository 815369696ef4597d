// bk_mcts_kernels.cuh — warp-per-game MCTS self-play (self_play/src/simulation.rs:37-296,
// self_play/src/node.rs:8-41) for sm_100a.
//
// Tree layout (per game, in HBM).  A reference `Node` is one CHILD ENTRY of its parent; entries of one
// parent are contiguous and in ascending tile order (the canonical stand-in for HashMap iteration,
// SURVEY.md Appendix D):
//     S[cap] = { N u32 visits, Q f32 = value_sum/visits (0 if unvisited), P f32 prior,
//                TN u32 = tile | n_children << 9 | to_play << 18 | expanded << 20 }      (select stream)
//     X[cap] = { W f32 value_sum, child block offset u32, node id u32, - }               (backup / descent stream)
// two coalesced 16-byte vector streams: lane i of a select issues S[off+i] and X[off+i] together, so one
// tree level costs ONE dependent memory round trip (the winner's child block offset, size, seat and node
// id arrive with its statistics; nothing is chased through the node table).  Q is refreshed by the backup,
// which removes one IEEE division per child per select.
// An entry that has been expanded points (node id) into the node table, whose records are full game states
// (BkState) plus the child block offset/count.  Keeping the state of every expanded node in HBM replaces
// the reference's "clone the game and replay the path" (simulation.rs:196-203) by ONE Game::apply per
// simulation — same states, same results, depth-times less rules work; 180 GB of HBM pays for it.
//
// Exactness (parity mode, the only mode here): one simulation in flight per game, simulations of a game
// in order; f32 UCB arithmetic with explicit round-to-nearest ops in the reference's source order
// (simulation.rs:88-98), the ln/sqrt factor from a host-built table indexed by parent visits, ties
// resolved to the highest tile (`>=` over ascending order, simulation.rs:141).
#pragma once
#include "bk_env_kernels.cuh"

#define BK_PATH_CAP 128
#define BK_NODE_NONE 0xFFFFFFFFu
#define BK_TN_TILE(tn) ((tn) & 0x1FFu)
#define BK_TN_NCHILD(tn) (((tn) >> 9) & 0x1FFu)
#define BK_TN_TOPLAY(tn) (((tn) >> 18) & 3u)
#define BK_TN_EXPANDED(tn) (((tn) >> 20) & 1u)
#define BK_TN_PENDING(tn) (((tn) >> 21) & 1u)   // throughput mode: leaf selected, evaluator answer outstanding

// opt-in throughput modes; 0 = the reference's exact behaviour
#define BK_MODE_SKIP_FORCED_FLAG 1u   // a root position with exactly one legal tile is not searched
#define BK_MODE_FORCE_VL_FLAG 2u      // run the multi-leaf code path even with one leaf per round (tests)
#define BK_MODE_TREE_REUSE_FLAG 4u    // the subtree of the played child becomes the next ply's tree
#define BK_MAX_LEAVES_PER_ROUND 32

#define BK_SP_ERR_ENTRY_CAP 1u
#define BK_SP_ERR_PATH_CAP 2u
#define BK_SP_ERR_NO_CHILD 4u   // select found no child with score >= 0 (reference would unwrap None)
#define BK_SP_ERR_POLICY_CAP 8u
#define BK_SP_ERR_APPLY 16u

// Lane src's value for every lane of the one-warp search, as a warp reduction instead of a shuffle: REDUX writes a UNIFORM
// register, so what depends on the value — the select loop's exit, the next block's address and size, the leaf's parent —
// stays warp-uniform for the compiler.  With every such value uniform (and the opt-in modes compiled out of the exact
// kernel) ptxas needs no convergence check (BRA.DIV) before any collective of the simulation loop, keeps the scalar
// state in uniform registers, and the 20-games-per-SM build fits its 96 registers without spilling.
#define BK_BCAST(v, src) __reduce_or_sync(BK_FULL, lane == (src) ? (v) : 0u)

struct BkSearchCfg {
    uint32_t sims, sample_moves;
    float frac, alpha;
    uint64_t seed;
    uint32_t first_game_id;
    uint32_t max_nodes, entry_cap, policy_cap;
    uint32_t mode, leaves_per_round;
    const float* ucb_tab;    // [sims + 2]: (ln((N + c_base + 1)/c_base) + c_init) * sqrt(N), host libm
    const float* rcp_tab;    // [sims + 2]: RN(1/d), or NULL — see bk_ucb_div
    const float* prior_tab;  // [401]: stub prior for n children = e / (e + e + ... n times), f32 sequential
    float stub_value;        // 0.25
};

struct BkTree {
    uint4* S;         // [entry_cap] {N, Q, P, TN}
    uint4* X;         // [entry_cap] {W, child_off, node id, -}
    BkState* nodes;   // [max_nodes]; pad[0] = child offset, pad[1] = child count
    double* scratch;  // [400]
    uint32_t* remap;  // [2 * max_nodes], tree-reuse mode only: new node id / new child-block offset per old node
};

// persisted per game between launches
struct BkSearchHdr {
    uint32_t n_nodes, n_entries, root_visits, sims_done;
    uint32_t pend_kind, pend_depth, pend_parent, pend_tile;  // 0 none, 1 root, 2 leaf awaiting evaluator
    uint32_t err, pol_count, plies_searched, pend_entry;
    uint32_t pend_count, forced_plies, reused, rsv1;  // multi-leaf: leaves outstanding; plies skipped as forced; tree kept
    uint32_t path[BK_PATH_CAP];
    uint8_t path_tp[BK_PATH_CAP];
};

// one outstanding leaf of the multi-leaf (virtual loss) mode
struct BkPend {
    uint32_t depth, parent, tile, entry, slot, rsv[3];
    uint32_t path[BK_PATH_CAP];
    uint8_t path_tp[BK_PATH_CAP];
};

struct BkSpCounters {  // lane-0 partial sums
    uint32_t sims, applies, entries, nodes;
};

// per-warp shared scratch.  E_CAP: room for the children's exp(policy) — 400 where an evaluator's policy arrives, 1 in the
// fixed-prior stub kernels, which never touch it (1.6 KB less shared memory per game: 28 resident games per SM instead of 24)
template <int E_CAP>
struct BkWarpSmemT {
    float e[E_CAP];
    uint16_t tile[400];
    uint32_t path[BK_PATH_CAP];
    uint32_t path_n[BK_PATH_CAP];   // visits / value_sum of the path entries as the select read them
    uint32_t path_w[BK_PATH_CAP];   // (valid only inside one kernel: the backup then needs no reload)
    uint8_t path_tp[BK_PATH_CAP];
};
typedef BkWarpSmemT<400> BkWarpSmem;
typedef BkWarpSmemT<1> BkWarpSmemStub;

// Optional streaming (evict-first) stores for what an expansion writes (-DBK_STREAM_STORES): of the ~10 k
// child entries and 800 node states a move tree creates only a few are ever read again.  Measured on B200
// (1024 games, 800 sims, complete games): 1297 ms with, 1283 ms without — within noise, so it is off.
#if defined(BK_WARP_EMU) || !defined(BK_STREAM_STORES)
#define BK_ST_STREAM(ptr, v) (*(ptr) = (v))
#else
#define BK_ST_STREAM(ptr, v) __stcs((ptr), (v))
#endif

__device__ __forceinline__ void bk_store_stream(BkState* __restrict__ s, int lane, const BkRegs& G) {
    if (lane < 20) {
        BK_ST_STREAM(reinterpret_cast<uint4*>(s->own) + lane, make_uint4(G.o0, G.o1, G.o2, G.o3));
        BK_ST_STREAM(&s->legal[lane], G.legal);
    }
    if (lane == 0) {
        BK_ST_STREAM(reinterpret_cast<uint4*>(s->pieces), make_uint4(G.pc0, G.pc1, G.pc2, G.pc3));
        BK_ST_STREAM(reinterpret_cast<uint4*>(&s->meta), make_uint4(G.meta, G.lastlens, G.t01, G.t23));
        BK_ST_STREAM(&s->ply, G.ply);
        BK_ST_STREAM(reinterpret_cast<uint4*>(&s->alive), make_uint4(G.alive, G.tw0, G.tw1, G.tw2));
    }
    BK_ST_STREAM(&s->smask[lane], (unsigned short)(G.smask));
}

// simulation.rs:92-94: u = F / (1 + N), an IEEE f32 division on the select's critical path (one per child per level).
// N + 1 = d is a small integer and F is one of the sims + 2 values of ucb_tab, so there are only (sims + 2)^2 possible
// divisions: the host checks ALL of them (bk_selfplay_create) against the three-operation form
//     q0 = RN(F * r), r = RN(1/d);   rem = fma(-d, q0, F) (exact);   q = fma(rem, r, q0)
// and passes rcp_tab only if every quotient equals RN(F / d) bit for bit; otherwise rcp_tab is NULL and the IEEE
// division is used.  Same result, a third of the latency (no MUFU.RCP + Newton + range fix-up).
__device__ __forceinline__ float bk_ucb_div(const BkSearchCfg& cfg, float F, uint32_t n_visits) {
    if (cfg.rcp_tab) {
        const float r = cfg.rcp_tab[n_visits + 1u];
        const float d = float(n_visits + 1u);
        const float q0 = __fmul_rn(F, r);
        return __fmaf_rn(__fmaf_rn(-d, q0, F), r, q0);
    }
    return __fdiv_rn(F, __fadd_rn(1.0f, float(n_visits)));
}

// f32 exp as the oracle defines it: exp in f64, rounded once to f32 (see oracle/mcts_oracle.hpp exp_f32)
__device__ __forceinline__ float bk_exp_f32(float x) { return float(exp(double(x))); }

__device__ __forceinline__ float bk_sel4f(int p, float a, float b, float c, float d) {
    return p == 0 ? a : (p == 1 ? b : (p == 2 ? c : d));
}

// index of abs tile (r, c) inside the mover-frame policy vector (inverse of game.rs:77-89 rotation,
// i.e. what simulation.rs:25-34 undoes `cur` times)
__device__ __forceinline__ int bk_frame_index(int r, int c, int cur) {
    if (cur == 0) return r * 20 + c;
    if (cur == 1) return (19 - c) * 20 + r;
    if (cur == 2) return (19 - r) * 20 + (19 - c);
    return c * 20 + (19 - r);
}

// evaluate()'s expansion half (simulation.rs:66-81): children for the legal tiles of state L (with
// policy > 0 when a policy is given), priors exp(p)/sum exp(p) summed sequentially in ascending order.
// Writes state L to node slot `hdr.n_nodes` and returns that node id, or BK_NODE_NONE when the state
// yields no child (the node then stays unexpanded, as in the reference) or a pool overflowed.
struct BkBlock { uint32_t off, n; };   // child block of the node an expansion created

template <class SM>
__device__ __forceinline__ uint32_t bk_tree_expand(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg,
                                                   const BkRegs& L, const float* policy, int lane, SM& sm,
                                                   BkSpCounters& ctr, BkBlock& blk, uint32_t slot = BK_NODE_NONE,
                                                   bool store_state = true) {
    blk.off = 0u; blk.n = 0u;
    const bool own_slot = slot == BK_NODE_NONE;        // exact mode: the next free node; multi-leaf: reserved
    const uint32_t id = own_slot ? hd.n_nodes : slot;
    const int cur = bk_cur(L);
    // per-lane pass over this row's legal tiles
    uint32_t m = L.legal;
    int cnt = 0;
    uint32_t keep = 0u;
    if (policy) {
        uint32_t mm = m;
        while (mm) {
            const int c = __ffs(mm) - 1;
            mm &= mm - 1u;
            const float p = policy[bk_frame_index(lane, c, cur)];
            if (p > 0.0f) { keep |= 1u << c; }
        }
        m = keep;
    }
    cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(BK_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int n = int(__reduce_add_sync(BK_FULL, unsigned(cnt)));      // (a reduction, not lane 31's scan value: uniform for the compiler)
    int pos = incl - cnt;
    {
        uint32_t mm = m;
        while (mm) {
            const int c = __ffs(mm) - 1;
            mm &= mm - 1u;
            sm.tile[pos] = uint16_t(lane * 20 + c);
            if (policy) sm.e[pos] = bk_exp_f32(policy[bk_frame_index(lane, c, cur)]);
            ++pos;
        }
    }
    __syncwarp();
    if (n == 0) return BK_NODE_NONE;
    if (hd.n_entries + uint32_t(n) > cfg.entry_cap || id >= cfg.max_nodes) {
        hd.err |= BK_SP_ERR_ENTRY_CAP;
        return BK_NODE_NONE;
    }
    float total = 0.0f;
    if (policy) {
        for (int i = 0; i < n; ++i) total = __fadd_rn(total, sm.e[i]);  // simulation.rs:74, ascending order
    }
    const uint32_t off = hd.n_entries;
    const float stub_prior = cfg.prior_tab[n];
    for (int i = lane; i < n; i += 32) {
        const float pr = policy ? __fdiv_rn(sm.e[i], total) : stub_prior;
        BK_ST_STREAM(&tr.S[off + i], make_uint4(0u, 0u, __float_as_uint(pr), uint32_t(sm.tile[i])));
        BK_ST_STREAM(&tr.X[off + i], make_uint4(0u, 0u, BK_NODE_NONE, 0u));
    }
    BkState* ns = &tr.nodes[id];
    if (store_state) bk_store_stream(ns, lane, L);
    if (lane == 0) { BK_ST_STREAM(&ns->pad[0], off); BK_ST_STREAM(&ns->pad[1], uint32_t(n)); }
    hd.n_entries += uint32_t(n);
    if (own_slot) hd.n_nodes += 1u;
    if (lane == 0) { ctr.entries += uint32_t(n); ctr.nodes += 1u; }
    blk.off = off; blk.n = uint32_t(n);
    __syncwarp();
    return id;
}

// after evaluate(): the leaf entry remembers the seat to move at it (node.to_play, simulation.rs:78) and,
// if it got children, its child block and node id
__device__ __forceinline__ void bk_tree_link(const BkTree& tr, uint32_t entry, int tile, uint32_t id, int to_play,
                                             const BkBlock& blk, int lane) {
    if (lane == 0) {
        uint32_t tn = uint32_t(tile) | (uint32_t(to_play) << 18);
        if (id != BK_NODE_NONE) {
            tn |= (blk.n << 9) | (1u << 20);
            tr.X[entry].y = blk.off;
            tr.X[entry].z = id;
        }
        tr.S[entry].w = tn;
    }
    __syncwarp();
}

// add_exploration_noise (simulation.rs:101-114) on the root's child block
__device__ __forceinline__ void bk_tree_noise(const BkTree& tr, const BkSearchCfg& cfg, uint32_t game_id, uint32_t ply,
                                              int lane) {
    const uint32_t off = tr.nodes[0].pad[0];
    const int n = int(tr.nodes[0].pad[1]);
    if (n <= 1) return;
    double mx = -1.0e300;
    for (int i = lane; i < n; i += 32) {
        const double lg = bk_log_gamma_draw(cfg.seed, game_id, ply, uint32_t(i), double(cfg.alpha));
        tr.scratch[i] = lg;
        mx = lg > mx ? lg : mx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double o = __shfl_xor_sync(BK_FULL, mx, d);
        mx = o > mx ? o : mx;
    }
    for (int i = lane; i < n; i += 32) tr.scratch[i] = bk_det_exp(__dadd_rn(tr.scratch[i], -mx));
    __syncwarp();
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum = __dadd_rn(sum, tr.scratch[i]);
    const float keepf = __fsub_rn(1.0f, cfg.frac);
    for (int i = lane; i < n; i += 32) {
        const float noise = float(__ddiv_rn(tr.scratch[i], sum));
        const float pr = __uint_as_float(tr.S[off + i].z);
        tr.S[off + i].z = __float_as_uint(__fadd_rn(__fmul_rn(pr, keepf), __fmul_rn(noise, cfg.frac)));
    }
    __syncwarp();
}

__device__ __forceinline__ void bk_prefetch_state(const BkState* s, int lane) {
#if !defined(BK_WARP_EMU) && !defined(BK_NO_PREFETCH)
    if (lane < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(s) + 128 * lane));
#else
    (void)s; (void)lane;
#endif
}

struct BkLeaf {
    uint32_t parent;  // node whose child entry is the leaf
    uint32_t entry;   // entry index of the leaf
    int tile;
    int depth;        // path length (>= 1)
    bool ok;
    bool collide;     // multi-leaf mode: the leaf is already waiting for the evaluator
};

// Optional (-DBK_CHILD_PREFETCH): ask L2 for the child blocks of every expanded child of this level while
// the scores are being computed.  Measured on B200: no gain (the kernel is bound by dependent issue
// latency, not by L2 misses: profiles/r01_ab_mcts_variants.log), so it is off.
__device__ __forceinline__ void bk_prefetch_block(const BkTree& tr, uint32_t tn, uint32_t off) {
#if !defined(BK_WARP_EMU) && defined(BK_CHILD_PREFETCH)
    if (BK_TN_EXPANDED(tn)) {
#if BK_CHILD_PREFETCH == 3
        uint32_t d0, d1;    // real loads whose results are never read: the lines land in L1, nothing waits for them
        asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(d0) : "l"(tr.S + off));
        asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(d1) : "l"(tr.X + off));
#elif BK_CHILD_PREFETCH == 2
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tr.S + off));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tr.X + off));
#else
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tr.S + off));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tr.X + off));
#endif
    }
#else
    (void)tr; (void)tn; (void)off;
#endif
}

// the selection loop of mcts() (simulation.rs:198-203) with select_child/ucb_score (:88-98,:135-147).
// `root` is the root's child block (kept in registers by the caller: no per-simulation reload).
// VL (multi-leaf throughput mode): every entry on the way down gets its visit at once — a visit worth 0
// until the value arrives, i.e. a virtual loss (Q = W / (N + 1)) that steers the next selections of the
// same round elsewhere; the backup then only adds the value.  One warp owns the tree, so no atomics.
// PF: ask L2 for the node state the leaf step will load, one level ahead.  It pays when a few games per SM are bound by
// latency (1024 games: the one-warp kernel), and costs 3 % when many resident games are bound by L2 / DRAM traffic (8192 games:
// the prefetches are half of the DRAM reads), so the high-residency instantiations switch it off.
template <bool VL, bool PF = true, class SM>
__device__ __forceinline__ BkLeaf bk_tree_select(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg,
                                                 const BkBlock& root, int lane, SM& sm) {
    uint32_t node = 0u;
    uint32_t off = root.off;
    int n = int(root.n);
    uint32_t Np = hd.root_visits;
    int depth = 0;
    uint32_t e = 0u, tn = 0u;
    bool no_child = false;          // the loop's only exits: this flag, or an unexpanded winner (the leaf)
    for (;;) {
        // If this level's winner turns out to be a leaf, the state of `node` is what the leaf step loads
        // next: ask L2 for its 5 lines now, one level of latency ahead.
        if (PF) bk_prefetch_state(&tr.nodes[node], lane);
        const float F = cfg.ucb_tab[Np];
        uint32_t wi, b_tn = 0u, b_n = 0u, b_w = 0u, b_off = 0u, b_node = 0u;
        if (n <= 32) {
            // one child per lane (the common case): score it, the last maximal child is the highest
            // lane holding the maximum (`>=` over ascending order, simulation.rs:141)
            uint32_t key = 0u;
            if (lane < n) {
                const uint4 sv = tr.S[off + lane];
                const uint4 xv = tr.X[off + lane];
                bk_prefetch_block(tr, sv.w, xv.y);
                const float u = bk_ucb_div(cfg, F, sv.x);                         // simulation.rs:92-94
                const float sc = __fadd_rn(__fmul_rn(u, __uint_as_float(sv.z)), __uint_as_float(sv.y));  // :95-97, Q cached
                if (sc >= 0.0f) key = __float_as_uint(sc) + 1u;                                     // NaN / negative never wins
                b_tn = sv.w; b_n = sv.x; b_w = xv.x; b_off = xv.y; b_node = xv.z;
            }
            const uint32_t kmax = __reduce_max_sync(BK_FULL, key);
            if (kmax == 0u) { no_child = true; break; }
            wi = 31u - uint32_t(__clz(int(__ballot_sync(BK_FULL, key == kmax))));
        } else {
            float best = 0.0f;
            int bi = -1;
            for (int i = lane; i < n; i += 32) {
                const uint4 sv = tr.S[off + i];
                const uint4 xv = tr.X[off + i];
                const float u = bk_ucb_div(cfg, F, sv.x);
                const float sc = __fadd_rn(__fmul_rn(u, __uint_as_float(sv.z)), __uint_as_float(sv.y));
                if (sc >= best) { best = sc; bi = i; b_tn = sv.w; b_n = sv.x; b_w = xv.x; b_off = xv.y; b_node = xv.z; }
            }
            const uint32_t key = bi >= 0 ? __float_as_uint(best) + 1u : 0u;
            const uint32_t kmax = __reduce_max_sync(BK_FULL, key);
            if (kmax == 0u) { no_child = true; break; }
            wi = __reduce_max_sync(BK_FULL, key == kmax ? uint32_t(bi) + 1u : 0u) - 1u;
        }
        const int src = int(wi & 31u);                     // child i lives on lane i % 32
        tn = BK_BCAST(b_tn, src);
        e = off + wi;
        if (lane == src) {                                 // the winner's lane holds everything the backup needs
            const int slot = depth & (BK_PATH_CAP - 1);    // a path longer than the cap is reported after the loop
            sm.path[slot] = e; sm.path_n[slot] = b_n; sm.path_w[slot] = b_w;
            sm.path_tp[slot] = uint8_t(BK_TN_TOPLAY(tn));
            if (VL && !BK_TN_PENDING(tn)) {
                const uint32_t nv = b_n + 1u;
                *reinterpret_cast<uint2*>(&tr.S[e]) =
                    make_uint2(nv, __float_as_uint(__fdiv_rn(__uint_as_float(b_w), float(nv))));
            }
        }
        ++depth;
        if (!BK_TN_EXPANDED(tn)) break;                    // the leaf
        Np = BK_BCAST(b_n, src);
        off = BK_BCAST(b_off, src);
        node = BK_BCAST(b_node, src);
        n = int(BK_TN_NCHILD(tn));
    }
    BkLeaf lf;
    lf.parent = node;
    lf.entry = e;
    lf.tile = int(BK_TN_TILE(tn));
    lf.collide = VL && !no_child && BK_TN_PENDING(tn);
    lf.depth = depth;
    lf.ok = true;
    if (no_child) { hd.err |= BK_SP_ERR_NO_CHILD; lf.ok = false; }
    if (depth > BK_PATH_CAP) { hd.err |= BK_SP_ERR_PATH_CAP; lf.ok = false; }
    __syncwarp();
    return lf;
}

// backpropagate (simulation.rs:164-171): every entry on the path gets +1 visit and the value of the seat
// to move AT that node (0 for a terminal / unexpanded leaf, node.rs:20).  The visit count and value sum
// the select read are still current (one simulation in flight per game), so nothing is reloaded.
template <class SM>
__device__ __forceinline__ void bk_tree_backup(const BkTree& tr, int depth, const float (&val)[4], int lane,
                                               const SM& sm) {
    for (int d = lane; d < depth; d += 32) {
        const uint32_t e = sm.path[d];
        const int tp = int(sm.path_tp[d]);
        const uint32_t nv = sm.path_n[d] + 1u;                                                       // visits += 1
        const float w = __fadd_rn(__uint_as_float(sm.path_w[d]), bk_sel4f(tp, val[0], val[1], val[2], val[3]));
        tr.X[e].x = __float_as_uint(w);                                                              // value_sum +=
        *reinterpret_cast<uint2*>(&tr.S[e]) = make_uint2(nv, __float_as_uint(__fdiv_rn(w, float(nv))));  // N, Q
    }
    __syncwarp();
}

// the tail of mcts() + select_action (simulation.rs:213-229, :118-130, :150-161): record the root's visit
// distribution, then pick the tile to play.
__device__ __forceinline__ int bk_tree_finish_ply(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg,
                                                  uint32_t game_id, uint32_t ply, uint32_t* pol_off, uint16_t* pol_tile,
                                                  uint32_t* pol_visits, int lane, uint32_t* picked_entry = nullptr) {
    const uint32_t off = tr.nodes[0].pad[0];
    const int n = int(tr.nodes[0].pad[1]);
    const uint32_t k = hd.plies_searched;
    // policy record
    if (hd.pol_count + uint32_t(n) > cfg.policy_cap || k >= BK_HIST_CAP) {
        hd.err |= BK_SP_ERR_POLICY_CAP;
    } else {
        for (int i = lane; i < n; i += 32) {
            const uint4 sv = tr.S[off + i];
            pol_tile[hd.pol_count + i] = uint16_t(BK_TN_TILE(sv.w));
            pol_visits[hd.pol_count + i] = sv.x;
        }
        if (lane == 0) { pol_off[k] = hd.pol_count; pol_off[k + 1] = hd.pol_count + uint32_t(n); }
        hd.pol_count += uint32_t(n);
    }
    hd.plies_searched = k + 1u;
    int pick = -1;
    if (hd.plies_searched < cfg.sample_moves) {
        uint32_t total = 0u;
        for (int i = lane; i < n; i += 32) total += tr.S[off + i].x;
        total = __reduce_add_sync(BK_FULL, total);
        const float u = bk_action_uniform(cfg.seed, game_id, ply);
        const float ft = float(total);
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) {                                            // simulation.rs:123-128
            sum = __fadd_rn(sum, __fdiv_rn(float(tr.S[off + i].x), ft));
            if (sum > u) { pick = i; break; }
        }
        if (pick < 0) pick = n - 1;                                              // simulation.rs:129
    } else {
        uint32_t bestv = 0u;
        int bi = -1;
        for (int i = lane; i < n; i += 32) {
            const uint32_t v = tr.S[off + i].x;
            if (v >= bestv) { bestv = v; bi = i; }                               // max_by keeps the last maximum
        }
        const uint32_t key = bi >= 0 ? bestv + 1u : 0u;
        const uint32_t kmax = __reduce_max_sync(BK_FULL, key);
        pick = int(__reduce_max_sync(BK_FULL, key == kmax ? uint32_t(bi) + 1u : 0u)) - 1;
    }
    if (picked_entry) *picked_entry = off + uint32_t(pick);
    return int(BK_TN_TILE(tr.S[off + uint32_t(pick)].w));
}

// write the single-child root of a forced ply: finish_ply then emits [(tile, sims)] and plays the tile
__device__ __forceinline__ bool bk_forced_root(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg, const BkRegs& G,
                                               int lane) {
    if (!(cfg.mode & BK_MODE_SKIP_FORCED_FLAG)) return false;
    if (bk_legal_count(G.legal) != 1) return false;
    const unsigned who = __ballot_sync(BK_FULL, G.legal != 0u);
    const int row = __ffs(who) - 1;
    const int tile = row * 20 + (__ffs(__shfl_sync(BK_FULL, G.legal, row)) - 1);
    if (lane == 0) {
        tr.S[0] = make_uint4(cfg.sims, 0u, __float_as_uint(1.0f), uint32_t(tile));
        tr.X[0] = make_uint4(0u, 0u, BK_NODE_NONE, 0u);
        tr.nodes[0].pad[0] = 0u; tr.nodes[0].pad[1] = 1u;
    }
    hd.n_nodes = 1u; hd.n_entries = 1u; hd.root_visits = cfg.sims; hd.sims_done = cfg.sims;
    hd.forced_plies += 1u;
    __syncwarp();
    return true;
}

// ---- tree reuse (opt-in: BK_MODE_TREE_REUSE) ---------------------------------------
// After the action is played, the subtree below the chosen root child is compacted IN PLACE to the front of
// the game's pools and becomes the next ply's tree; the next search then only tops the root up to
// sims_per_move visits.  (The reference builds a new tree every ply, simulation.rs:183 — visit counts differ.)
// Node ids and child blocks are allocated in expansion order and a child is always expanded after its parent,
// so walking the old ids upwards (a) sees a node's membership decided before the node itself and (b) copies
// every record to an index <= its old one: a forward copy never overwrites what is still to be read.
// Returns false (fresh tree next ply) when the played child was never expanded.
__device__ __forceinline__ bool bk_tree_reroot(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg,
                                               uint32_t played_entry, int lane) {
    const uint4 sa = tr.S[played_entry];
    if (!BK_TN_EXPANDED(sa.w)) return false;
    const uint32_t r = tr.X[played_entry].z;
    const uint32_t nn = hd.n_nodes;
    uint32_t* newid = tr.remap;                      // 0 = not in the subtree, else new id + 1
    uint32_t* newoff = tr.remap + cfg.max_nodes;
    for (uint32_t i = lane; i < nn; i += 32) newid[i] = 0u;
    __syncwarp();
    if (lane == 0) newid[r] = 1u;
    __syncwarp();
    // pass 1: membership, top-down in ascending id order
    for (uint32_t id = r; id < nn; ++id) {
        if (newid[id] == 0u) continue;                                   // warp-uniform (same address on every lane)
        const uint32_t off = tr.nodes[id].pad[0], n = tr.nodes[id].pad[1];
        for (uint32_t i = lane; i < n; i += 32)
            if (BK_TN_EXPANDED(tr.S[off + i].w)) newid[tr.X[off + i].z] = 1u;
        __syncwarp();
    }
    // pass 2: new ids / new child-block offsets = ranks by old id (chunked warp scans)
    uint32_t base_id = 0u, base_off = 0u;
    for (uint32_t c = r; c < nn; c += 32) {
        const uint32_t id = c + uint32_t(lane);
        const uint32_t m = (id < nn && newid[id]) ? 1u : 0u;
        const uint32_t cnt = m ? tr.nodes[id].pad[1] : 0u;
        uint32_t im = m, ic = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t vm = __shfl_up_sync(BK_FULL, im, d), vc = __shfl_up_sync(BK_FULL, ic, d);
            if (lane >= d) { im += vm; ic += vc; }
        }
        if (m) { newid[id] = base_id + (im - m) + 1u; newoff[id] = base_off + (ic - cnt); }
        base_id += __reduce_add_sync(BK_FULL, m);          // (= lane 31's scan values, as reductions: uniform for the compiler)
        base_off += __reduce_add_sync(BK_FULL, cnt);
    }
    __syncwarp();
    // pass 3: forward copy with pointer rewrite
    for (uint32_t id = r; id < nn; ++id) {
        const uint32_t ni1 = newid[id];
        if (ni1 == 0u) continue;
        const uint32_t ni = ni1 - 1u, noff = newoff[id];
        const uint32_t off = tr.nodes[id].pad[0], n = tr.nodes[id].pad[1];
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t i = i0 + uint32_t(lane);
            uint4 sv = make_uint4(0u, 0u, 0u, 0u), xv = sv;
            if (i < n) {
                sv = tr.S[off + i]; xv = tr.X[off + i];
                if (BK_TN_EXPANDED(sv.w)) { const uint32_t c = xv.z; xv.y = newoff[c]; xv.z = newid[c] - 1u; }
            }
            __syncwarp();                                                 // every lane has read before any lane writes
            if (i < n) { tr.S[noff + i] = sv; tr.X[noff + i] = xv; }
            __syncwarp();
        }
        if (ni != id) {
            BkRegs T;
            bk_load(&tr.nodes[id], lane, T);
            __syncwarp();
            bk_store(&tr.nodes[ni], lane, T);
        }
        if (lane == 0) { tr.nodes[ni].pad[0] = noff; tr.nodes[ni].pad[1] = n; }
        __syncwarp();
    }
    hd.n_nodes = base_id;
    hd.n_entries = base_off;
    hd.reused = 1u;
    return true;
}

// start of a ply on a kept tree: the root's visits so far are its children's (each pass through the node chose one)
__device__ __forceinline__ BkBlock bk_tree_resume(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg, int lane) {
    BkBlock root;
    root.off = tr.nodes[0].pad[0]; root.n = tr.nodes[0].pad[1];
    uint32_t tot = 0u;
    for (uint32_t i = lane; i < root.n; i += 32) tot += tr.S[root.off + i].x;
    tot = __reduce_add_sync(BK_FULL, tot);
    if (tot > cfg.sims) tot = cfg.sims;                  // (cannot exceed: the child had at most sims visits)
    hd.root_visits = tot;
    hd.sims_done = tot;
    if ((cfg.mode & BK_MODE_SKIP_FORCED_FLAG) && root.n == 1u) {     // forced ply on a kept tree: all visits to the one child
        if (lane == 0) tr.S[root.off].x = cfg.sims;
        hd.root_visits = cfg.sims; hd.sims_done = cfg.sims;
        hd.forced_plies += 1u;
        __syncwarp();
    }
    return root;
}

// One simulation's leaf step for the fixed-prior stub: apply the leaf tile to the parent's state,
// evaluate (terminal payoff, or stub value + expansion), back up.
template <bool PF, class SM>
__device__ __forceinline__ void bk_sim_stub(const BkTree& tr, BkSearchHdr& hd, const BkSearchCfg& cfg,
                                            const BkBlock& root, int lane, const BkTabs& tabs, SM& sm,
                                            BkCounters& gctr, BkSpCounters& ctr) {
    hd.root_visits += 1u;                                                        // simulation.rs:194
    const BkLeaf lf = bk_tree_select<false, PF>(tr, hd, cfg, root, lane, sm);
    if (!lf.ok) return;
    BkRegs L;
    bk_load(&tr.nodes[lf.parent], lane, L);
    if (!bk_apply(L, lf.tile, -1, lane, tabs, gctr)) { hd.err |= BK_SP_ERR_APPLY; return; }
    if (lane == 0) ctr.applies += 1u;
    float val[4];
    int tp = 0;
    if (bk_terminal(L)) {
        bk_payoff(L, val);                                                       // simulation.rs:45-47
    } else {
        BkBlock blk;
        const uint32_t id = bk_tree_expand(tr, hd, cfg, L, nullptr, lane, sm, ctr, blk);
        tp = bk_cur(L);                                                          // simulation.rs:78
        bk_tree_link(tr, lf.entry, lf.tile, id, tp, blk, lane);
        val[0] = val[1] = val[2] = val[3] = cfg.stub_value;
    }
    if (lane == 0) sm.path_tp[lf.depth - 1] = uint8_t(tp);
    __syncwarp();
    bk_tree_backup(tr, lf.depth, val, lane, sm);
    hd.sims_done += 1u;
    if (lane == 0) ctr.sims += 1u;
}

// training_game() (simulation.rs:267-296) with the stub evaluator, up to max_plies plies, on one warp.
// MODES = false: the exact reference behaviour only — the opt-in modes (forced-ply shortcut, tree reuse) are compiled out,
// cfg.mode must be 0 (then no header has `reused` set: bk_selfplay_set_mode clears it when the mode is left).
template <bool PF, bool MODES, class SM>
__device__ __forceinline__ void kb_selfplay_stub(const BkSearchCfg& cfg, BkState* __restrict__ states,
                                                 uint16_t* __restrict__ hist, const BkTree& tr, BkSearchHdr* hdr_g,
                                                 uint32_t* pol_off, uint16_t* pol_tile, uint32_t* pol_visits,
                                                 int max_plies, unsigned long long* counters, int g, int lane,
                                                 const BkTabs& tabs, SM& sm) {
    BkRegs G;
    bk_load(&states[g], lane, G);
    BkSearchHdr hd;
    hd.err = hdr_g->err;
    hd.pol_count = hdr_g->pol_count;
    hd.plies_searched = hdr_g->plies_searched;
    hd.forced_plies = hdr_g->forced_plies;
    BkCounters gctr = {0u, 0u};
    BkSpCounters ctr = {0u, 0u, 0u, 0u};
    const uint32_t game_id = cfg.first_game_id + uint32_t(g);
    int plies = 0;
    hd.reused = hdr_g->reused;
    hd.n_nodes = hdr_g->n_nodes; hd.n_entries = hdr_g->n_entries;            // (meaningful only with a kept tree)
    while (!bk_terminal(G) && (max_plies < 0 || plies < max_plies) && hd.err == 0u) {
        BkBlock root;
        bool search = true;
        if (MODES && hd.reused) {                                                      // (opt-in tree reuse)
            root = bk_tree_resume(tr, hd, cfg, lane);
        } else {
            hd.n_nodes = 0u; hd.n_entries = 0u; hd.root_visits = 0u; hd.sims_done = 0u;   // fresh tree, :183
            search = !(MODES && bk_forced_root(tr, hd, cfg, G, lane));                 // (opt-in shortcut, off by default)
            if (search) bk_tree_expand(tr, hd, cfg, G, nullptr, lane, sm, ctr, root);  // evaluate(root), :184
        }
        if (search) {
            bk_tree_noise(tr, cfg, game_id, G.ply, lane);                              // :190
            while (hd.sims_done < cfg.sims && hd.err == 0u) bk_sim_stub<PF>(tr, hd, cfg, root, lane, tabs, sm, gctr, ctr);
        }
        if (hd.err) break;
        uint32_t played = 0u;
        const int action = int(__reduce_max_sync(BK_FULL, unsigned(bk_tree_finish_ply(tr, hd, cfg, game_id, G.ply, pol_off, pol_tile,
                                                                                        pol_visits, lane, &played))));   // (uniform for the compiler)
        const int p = bk_cur(G);
        const uint32_t ply = G.ply;
        if (!bk_apply(G, action, -1, lane, tabs, gctr)) { hd.err |= BK_SP_ERR_APPLY; break; }   // :288
        if (lane == 0 && ply < BK_HIST_CAP) hist[size_t(g) * BK_HIST_CAP + ply] = uint16_t(action | (p << 9));
        ++plies;
        hd.reused = 0u;
        if (MODES && (cfg.mode & BK_MODE_TREE_REUSE_FLAG) && !bk_terminal(G))
            bk_tree_reroot(tr, hd, cfg, __reduce_max_sync(BK_FULL, played), lane);
    }
    bk_store(&states[g], lane, G);
    if (lane == 0) {
        hdr_g->err = hd.err;
        hdr_g->pol_count = hd.pol_count;
        hdr_g->plies_searched = hd.plies_searched;
        hdr_g->n_nodes = hd.n_nodes;
        hdr_g->n_entries = hd.n_entries;
        hdr_g->root_visits = hd.root_visits;
        hdr_g->sims_done = hd.sims_done;
        hdr_g->forced_plies = hd.forced_plies;
        hdr_g->reused = hd.reused;
        hdr_g->pend_kind = 0u;
    }
    const unsigned crem = __reduce_add_sync(BK_FULL, gctr.crem);
    if (lane == 0 && counters) {
        atomicAdd(&counters[0], (unsigned long long)ctr.sims);
        atomicAdd(&counters[1], (unsigned long long)ctr.applies + (unsigned long long)plies);
        atomicAdd(&counters[2], (unsigned long long)gctr.movegens);
        atomicAdd(&counters[3], 120ull * (unsigned long long)crem);
        atomicAdd(&counters[4], (unsigned long long)ctr.entries);
        atomicAdd(&counters[5], (unsigned long long)ctr.nodes);
    }
}

// ---- external-evaluator protocol (one leaf per live game per round) ---------------------------------------
// pend_kind values
#define BK_PEND_NONE 0u   // no ply in progress (or the game is over)
#define BK_PEND_ROOT 1u   // root position waiting for the evaluator          (simulation.rs:183-189)
#define BK_PEND_LEAF 2u   // a leaf position waiting for the evaluator         (simulation.rs:206)
#define BK_PEND_DONE 3u   // all simulations of this ply are done; waiting for bk_selfplay_end_ply
#define BK_PEND_RESUME 4u // tree-reuse mode: the root is already expanded; the next step call runs on (no answer to consume)

__device__ __forceinline__ void bk_hdr_load(const BkSearchHdr* h, BkSearchHdr& hd, const BkTree& tr, int lane,
                                            BkWarpSmem& sm) {
    hd.n_nodes = h->n_nodes; hd.n_entries = h->n_entries; hd.root_visits = h->root_visits; hd.sims_done = h->sims_done;
    hd.pend_kind = h->pend_kind; hd.pend_depth = h->pend_depth; hd.pend_parent = h->pend_parent; hd.pend_tile = h->pend_tile;
    hd.err = h->err; hd.pol_count = h->pol_count; hd.plies_searched = h->plies_searched; hd.pend_entry = h->pend_entry;
    hd.pend_count = h->pend_count; hd.forced_plies = h->forced_plies; hd.reused = h->reused;
    for (int d = lane; d < int(hd.pend_depth) && d < BK_PATH_CAP; d += 32) {
        const uint32_t e = h->path[d];
        sm.path[d] = e; sm.path_tp[d] = h->path_tp[d];
        sm.path_n[d] = tr.S[e].x; sm.path_w[d] = tr.X[e].x;    // what the select had read (nothing ran in between)
    }
    __syncwarp();
}

__device__ __forceinline__ void bk_hdr_store(BkSearchHdr* h, const BkSearchHdr& hd, int lane, const BkWarpSmem& sm) {
    __syncwarp();
    for (int d = lane; d < int(hd.pend_depth) && d < BK_PATH_CAP; d += 32) { h->path[d] = sm.path[d]; h->path_tp[d] = sm.path_tp[d]; }
    if (lane == 0) {
        h->n_nodes = hd.n_nodes; h->n_entries = hd.n_entries; h->root_visits = hd.root_visits; h->sims_done = hd.sims_done;
        h->pend_kind = hd.pend_kind; h->pend_depth = hd.pend_depth; h->pend_parent = hd.pend_parent; h->pend_tile = hd.pend_tile;
        h->err = hd.err; h->pol_count = hd.pol_count; h->plies_searched = hd.plies_searched; h->pend_entry = hd.pend_entry;
        h->pend_count = hd.pend_count; h->forced_plies = hd.forced_plies; h->reused = hd.reused;
    }
}

// start mcts() for one game: fresh tree, the root position becomes the pending leaf (tentative node 0)
__device__ __forceinline__ void kb_sp_begin(const BkSearchCfg& cfg, const BkState* __restrict__ states, const BkTree& tr,
                                            BkSearchHdr* hdr_g, int g, int lane) {
    BkRegs G;
    bk_load(&states[g], lane, G);
    const bool live = !bk_terminal(G) && hdr_g->err == 0u;
    BkSearchHdr hd;
    hd.forced_plies = hdr_g->forced_plies;
    hd.reused = live ? hdr_g->reused : 0u;
    uint32_t kind = live ? BK_PEND_ROOT : BK_PEND_NONE;
    if (hd.reused) {                                                       // opt-in tree reuse: the root is expanded already
        hd.n_nodes = hdr_g->n_nodes; hd.n_entries = hdr_g->n_entries;
        bk_tree_resume(tr, hd, cfg, lane);
        bk_tree_noise(tr, cfg, cfg.first_game_id + uint32_t(g), G.ply, lane);
        kind = hd.sims_done >= cfg.sims ? BK_PEND_DONE : BK_PEND_RESUME;
    } else {
        hd.n_nodes = 0u; hd.n_entries = 0u; hd.root_visits = 0u; hd.sims_done = 0u;
        const bool forced = live && bk_forced_root(tr, hd, cfg, G, lane);  // opt-in shortcut, off by default
        if (live && !forced) bk_store(&tr.nodes[0], lane, G);
        if (forced) kind = BK_PEND_DONE;
    }
    __syncwarp();                          // every lane has read the header before lane 0 rewrites it
    if (lane == 0) {
        hdr_g->n_nodes = hd.n_nodes; hdr_g->n_entries = hd.n_entries; hdr_g->root_visits = hd.root_visits;
        hdr_g->sims_done = hd.sims_done; hdr_g->forced_plies = hd.forced_plies; hdr_g->reused = hd.reused;
        hdr_g->pend_depth = 0u; hdr_g->pend_parent = 0u; hdr_g->pend_tile = 0u; hdr_g->pend_entry = 0u;
        hdr_g->pend_count = 0u;
        hdr_g->pend_kind = kind;
    }
}

// consume the evaluator's answer for the pending position, then run simulations until the next
// non-terminal leaf needs the evaluator (or the ply's simulations are exhausted)
__device__ __forceinline__ void kb_sp_step(const BkSearchCfg& cfg, const BkTree& tr, BkSearchHdr* hdr_g,
                                           const float* __restrict__ policy, const float* __restrict__ value,
                                           unsigned long long* counters, int g, int lane, const BkTabs& tabs,
                                           BkWarpSmem& sm) {
    BkSearchHdr hd;
    bk_hdr_load(hdr_g, hd, tr, lane, sm);
    if (hd.pend_kind != BK_PEND_ROOT && hd.pend_kind != BK_PEND_LEAF && hd.pend_kind != BK_PEND_RESUME) return;
    BkCounters gctr = {0u, 0u};
    BkSpCounters ctr = {0u, 0u, 0u, 0u};
    const uint32_t game_id = cfg.first_game_id + uint32_t(g);
    const float* pol = policy;                           // the caller passes this game's dense evaluator row
    BkRegs L;
    BkBlock root;
    if (hd.pend_kind == BK_PEND_RESUME) {                // kept tree: nothing to consume, run on
        root.off = tr.nodes[0].pad[0]; root.n = tr.nodes[0].pad[1];
    } else if (hd.pend_kind == BK_PEND_ROOT) {
        bk_load(&tr.nodes[hd.n_nodes], lane, L);        // the pending position (tentative node slot)
        bk_tree_expand(tr, hd, cfg, L, pol, lane, sm, ctr, root);                      // evaluate(root), value dropped
        if (hd.n_nodes == 0u) { hd.err |= BK_SP_ERR_NO_CHILD; }                         // reference: unwrap on None
        else bk_tree_noise(tr, cfg, game_id, L.ply, lane);
    } else {
        bk_load(&tr.nodes[hd.n_nodes], lane, L);
        root.off = tr.nodes[0].pad[0]; root.n = tr.nodes[0].pad[1];
        BkBlock blk;
        const uint32_t id = bk_tree_expand(tr, hd, cfg, L, pol, lane, sm, ctr, blk);
        const int cur = bk_cur(L);
        bk_tree_link(tr, hd.pend_entry, int(hd.pend_tile), id, cur, blk, lane);
        float val[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) val[i] = value[(i + 4 - cur) & 3];                     // value.rotate_right(cur)
        if (lane == 0) sm.path_tp[hd.pend_depth - 1] = uint8_t(cur);
        __syncwarp();
        bk_tree_backup(tr, int(hd.pend_depth), val, lane, sm);
        hd.sims_done += 1u;
        if (lane == 0) ctr.sims += 1u;
    }
    hd.pend_kind = BK_PEND_DONE;
    hd.pend_depth = 0u;
    while (hd.sims_done < cfg.sims && hd.err == 0u) {
        hd.root_visits += 1u;
        const BkLeaf lf = bk_tree_select<false>(tr, hd, cfg, root, lane, sm);
        if (!lf.ok) break;
        bk_load(&tr.nodes[lf.parent], lane, L);
        if (!bk_apply(L, lf.tile, -1, lane, tabs, gctr)) { hd.err |= BK_SP_ERR_APPLY; break; }
        if (lane == 0) ctr.applies += 1u;
        if (bk_terminal(L)) {
            float val[4];
            bk_payoff(L, val);
            if (lane == 0) sm.path_tp[lf.depth - 1] = 0;
            __syncwarp();
            bk_tree_backup(tr, lf.depth, val, lane, sm);
            hd.sims_done += 1u;
            if (lane == 0) ctr.sims += 1u;
            continue;
        }
        if (hd.n_nodes >= cfg.max_nodes) { hd.err |= BK_SP_ERR_ENTRY_CAP; break; }
        bk_store(&tr.nodes[hd.n_nodes], lane, L);       // tentative: becomes node n_nodes if it gets children
        hd.pend_kind = BK_PEND_LEAF;
        hd.pend_depth = uint32_t(lf.depth);
        hd.pend_parent = lf.parent;
        hd.pend_tile = uint32_t(lf.tile);
        hd.pend_entry = lf.entry;
        break;
    }
    bk_hdr_store(hdr_g, hd, lane, sm);
    const unsigned crem = __reduce_add_sync(BK_FULL, gctr.crem);
    if (lane == 0 && counters) {
        atomicAdd(&counters[0], (unsigned long long)ctr.sims);
        atomicAdd(&counters[1], (unsigned long long)ctr.applies);
        atomicAdd(&counters[2], (unsigned long long)gctr.movegens);
        atomicAdd(&counters[3], 120ull * (unsigned long long)crem);
        atomicAdd(&counters[4], (unsigned long long)ctr.entries);
        atomicAdd(&counters[5], (unsigned long long)ctr.nodes);
    }
}

// the tail of mcts() and the game.apply of training_game() for games whose simulations are done
__device__ __forceinline__ void kb_sp_end(const BkSearchCfg& cfg, BkState* __restrict__ states, uint16_t* __restrict__ hist,
                                          const BkTree& tr, BkSearchHdr* hdr_g, uint32_t* pol_off, uint16_t* pol_tile,
                                          uint32_t* pol_visits, unsigned long long* counters, int g, int lane,
                                          const BkTabs& tabs, BkWarpSmem& sm) {
    BkSearchHdr hd;
    bk_hdr_load(hdr_g, hd, tr, lane, sm);
    if (hd.pend_kind != BK_PEND_DONE) return;
    if (hd.err != 0u) {
        // A search that stopped on an error (no selectable root child, a full pool, a path over the cap) has no
        // trustworthy root block: nothing is recorded or played; the host reports hd.err (sp_check_errors).
        __syncwarp();
        if (lane == 0) hdr_g->pend_kind = BK_PEND_NONE;
        return;
    }
    BkRegs G;
    bk_load(&states[g], lane, G);
    BkCounters gctr = {0u, 0u};
    const uint32_t game_id = cfg.first_game_id + uint32_t(g);
    uint32_t played = 0u;
    const int action = bk_tree_finish_ply(tr, hd, cfg, game_id, G.ply, pol_off, pol_tile, pol_visits, lane, &played);
    const int p = bk_cur(G);
    const uint32_t ply = G.ply;
    hd.reused = 0u;
    if (!bk_apply(G, action, -1, lane, tabs, gctr)) hd.err |= BK_SP_ERR_APPLY;
    else {
        if (lane == 0 && ply < BK_HIST_CAP) hist[size_t(g) * BK_HIST_CAP + ply] = uint16_t(action | (p << 9));
        bk_store(&states[g], lane, G);
        if ((cfg.mode & BK_MODE_TREE_REUSE_FLAG) && !bk_terminal(G)) bk_tree_reroot(tr, hd, cfg, played, lane);
    }
    hd.pend_kind = BK_PEND_NONE;
    hd.pend_depth = 0u;
    bk_hdr_store(hdr_g, hd, lane, sm);
    const unsigned crem = __reduce_add_sync(BK_FULL, gctr.crem);
    if (lane == 0 && counters) {
        atomicAdd(&counters[1], 1ull);
        atomicAdd(&counters[2], (unsigned long long)gctr.movegens);
        atomicAdd(&counters[3], 120ull * (unsigned long long)crem);
    }
}


// ---- throughput modes (opt-in, NOT visit-count comparable with the reference) ----
//
// (1) Forced plies: a root position with exactly one legal tile has one possible policy record,
//     [(tile, 1.0)], and one possible action, whatever the simulations do (simulation.rs:213-229); with
//     BK_MODE_SKIP_FORCED the 800 simulations are not run and the record is written directly.  The returned
//     training tuple is IDENTICAL to the exact mode's (tests/test_emu_mcts.py, tests/test_gpu_mcts.py).
// (2) Multi-leaf rounds with virtual loss: up to leaves_per_round simulations of a game are in flight at once,
//     so one evaluator round serves n_games * leaves_per_round positions (a small batch of games can fill the
//     tensor cores).  Selections of one round see each other's visits as losses (bk_tree_select<true>).  A
//     selection that arrives at a leaf already waiting for the evaluator is rolled back and ends the round's
//     collection.  Invariants kept: every simulation is one evaluated (or terminal) leaf; sum of the root's
//     child visits == sims_per_move; with one leaf per round the results equal the exact mode bit for bit.

// backup of the multi-leaf mode: the visits were given by the select; add the value, refresh Q
__device__ __forceinline__ void bk_tree_backup_vl(const BkTree& tr, int depth, const float (&val)[4], int lane,
                                                  const BkWarpSmem& sm) {
    for (int d = lane; d < depth; d += 32) {
        const uint32_t e = sm.path[d];
        const int tp = int(sm.path_tp[d]);
        const uint32_t nv = tr.S[e].x;
        const float w = __fadd_rn(__uint_as_float(tr.X[e].x), bk_sel4f(tp, val[0], val[1], val[2], val[3]));
        tr.X[e].x = __float_as_uint(w);
        tr.S[e].y = __float_as_uint(__fdiv_rn(w, float(nv)));
    }
    __syncwarp();
}

// take back the visits a rolled-back selection gave to the interior entries of its path
__device__ __forceinline__ void bk_tree_unvisit(const BkTree& tr, int depth, int lane, const BkWarpSmem& sm) {
    for (int d = lane; d < depth; d += 32) {
        const uint32_t e = sm.path[d];
        const uint32_t nv = tr.S[e].x - 1u;
        const float q = nv ? __fdiv_rn(__uint_as_float(tr.X[e].x), float(nv)) : 0.0f;
        *reinterpret_cast<uint2*>(&tr.S[e]) = make_uint2(nv, __float_as_uint(q));
    }
    __syncwarp();
}

// multi-leaf counterpart of kb_sp_step: policy/value hold leaves_per_round slots per game
__device__ __forceinline__ void kb_sp_step_vl(const BkSearchCfg& cfg, const BkTree& tr, BkSearchHdr* hdr_g, BkPend* pend_g,
                                              const float* __restrict__ policy, const float* __restrict__ value,
                                              unsigned long long* counters, int g, int lane, const BkTabs& tabs,
                                              BkWarpSmem& sm) {
    BkSearchHdr hd;
    hd.n_nodes = hdr_g->n_nodes; hd.n_entries = hdr_g->n_entries; hd.root_visits = hdr_g->root_visits;
    hd.sims_done = hdr_g->sims_done; hd.pend_kind = hdr_g->pend_kind; hd.pend_count = hdr_g->pend_count;
    hd.err = hdr_g->err; hd.pol_count = hdr_g->pol_count; hd.plies_searched = hdr_g->plies_searched;
    hd.forced_plies = hdr_g->forced_plies;
    if (hd.pend_kind != BK_PEND_ROOT && hd.pend_kind != BK_PEND_LEAF && hd.pend_kind != BK_PEND_RESUME) return;
    BkCounters gctr = {0u, 0u};
    BkSpCounters ctr = {0u, 0u, 0u, 0u};
    const uint32_t game_id = cfg.first_game_id + uint32_t(g);
    const uint32_t K = cfg.leaves_per_round;
    const float* pol = policy;          // the caller passes this game's first dense evaluator row
    const float* vals = value;
    BkRegs L;
    BkBlock root;
    if (hd.pend_kind == BK_PEND_RESUME) {                // kept tree: nothing to consume, run on
        root.off = tr.nodes[0].pad[0]; root.n = tr.nodes[0].pad[1];
    } else if (hd.pend_kind == BK_PEND_ROOT) {
        bk_load(&tr.nodes[0], lane, L);
        const uint32_t id = bk_tree_expand(tr, hd, cfg, L, pol, lane, sm, ctr, root, 0u, false);   // evaluate(root)
        hd.n_nodes = 1u;
        if (id == BK_NODE_NONE) hd.err |= BK_SP_ERR_NO_CHILD;
        else bk_tree_noise(tr, cfg, game_id, L.ply, lane);
    } else {
        root.off = tr.nodes[0].pad[0]; root.n = tr.nodes[0].pad[1];
        for (uint32_t j = 0; j < hd.pend_count && hd.err == 0u; ++j) {       // answers are consumed in selection order
            const BkPend* P = &pend_g[j];
            const uint32_t slot = P->slot, entry = P->entry;
            const int depth = int(P->depth);
            bk_load(&tr.nodes[slot], lane, L);
            BkBlock blk;
            const uint32_t id = bk_tree_expand(tr, hd, cfg, L, pol + size_t(j) * 400, lane, sm, ctr, blk, slot, false);
            const int cur = bk_cur(L);
            bk_tree_link(tr, entry, int(P->tile), id, cur, blk, lane);      // rewrites TN: the pending mark goes
            float val[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) val[i] = vals[size_t(j) * 4 + ((i + 4 - cur) & 3)];   // value.rotate_right(cur)
            for (int d = lane; d < depth; d += 32) { sm.path[d] = P->path[d]; sm.path_tp[d] = P->path_tp[d]; }
            __syncwarp();
            if (lane == 0) sm.path_tp[depth - 1] = uint8_t(cur);
            __syncwarp();
            bk_tree_backup_vl(tr, depth, val, lane, sm);
            hd.sims_done += 1u;
            if (lane == 0) ctr.sims += 1u;
        }
    }
    hd.pend_count = 0u;
    while (hd.sims_done + hd.pend_count < cfg.sims && hd.pend_count < K && hd.err == 0u) {
        hd.root_visits += 1u;
        const BkLeaf lf = bk_tree_select<true>(tr, hd, cfg, root, lane, sm);
        if (!lf.ok) break;
        if (lf.collide) {                                  // leaf already outstanding: roll back, close the round
            bk_tree_unvisit(tr, lf.depth - 1, lane, sm);
            hd.root_visits -= 1u;
            break;
        }
        bk_load(&tr.nodes[lf.parent], lane, L);
        if (!bk_apply(L, lf.tile, -1, lane, tabs, gctr)) { hd.err |= BK_SP_ERR_APPLY; break; }
        if (lane == 0) ctr.applies += 1u;
        if (bk_terminal(L)) {
            float val[4];
            bk_payoff(L, val);
            if (lane == 0) sm.path_tp[lf.depth - 1] = 0;
            __syncwarp();
            bk_tree_backup_vl(tr, lf.depth, val, lane, sm);
            hd.sims_done += 1u;
            if (lane == 0) ctr.sims += 1u;
            continue;
        }
        if (hd.n_nodes >= cfg.max_nodes) { hd.err |= BK_SP_ERR_ENTRY_CAP; break; }
        const uint32_t slot = hd.n_nodes;
        hd.n_nodes += 1u;                                   // reserved; stays unused if the leaf gets no child
        bk_store(&tr.nodes[slot], lane, L);
        BkPend* P = &pend_g[hd.pend_count];
        for (int d = lane; d < lf.depth; d += 32) { P->path[d] = sm.path[d]; P->path_tp[d] = sm.path_tp[d]; }
        if (lane == 0) {
            P->depth = uint32_t(lf.depth); P->parent = lf.parent; P->tile = uint32_t(lf.tile); P->entry = lf.entry;
            P->slot = slot;
            tr.S[lf.entry].w |= 1u << 21;                   // BK_TN_PENDING
        }
        __syncwarp();
        hd.pend_count += 1u;
    }
    hd.pend_kind = hd.pend_count ? BK_PEND_LEAF : BK_PEND_DONE;
    if (lane == 0) {
        hdr_g->n_nodes = hd.n_nodes; hdr_g->n_entries = hd.n_entries; hdr_g->root_visits = hd.root_visits;
        hdr_g->sims_done = hd.sims_done; hdr_g->pend_kind = hd.pend_kind; hdr_g->pend_count = hd.pend_count;
        hdr_g->pend_depth = 0u; hdr_g->err = hd.err; hdr_g->pol_count = hd.pol_count;
        hdr_g->plies_searched = hd.plies_searched; hdr_g->forced_plies = hd.forced_plies;
    }
    const unsigned crem = __reduce_add_sync(BK_FULL, gctr.crem);
    if (lane == 0 && counters) {
        atomicAdd(&counters[0], (unsigned long long)ctr.sims);
        atomicAdd(&counters[1], (unsigned long long)ctr.applies);
        atomicAdd(&counters[2], (unsigned long long)gctr.movegens);
        atomicAdd(&counters[3], 120ull * (unsigned long long)crem);
        atomicAdd(&counters[4], (unsigned long long)ctr.entries);
        atomicAdd(&counters[5], (unsigned long long)ctr.nodes);
    }
}

// ---- training tensors (the consumer side, model/training.py:70-119 `save()`), built on the device -------------
// For ply i of a game with mover p: planes 0..3 = squares laid BEFORE ply i by seats p, p+1, p+2, p+3; plane 4 =
// the tiles of the recorded policy (the root's children); all turned p quarter turns (torch.rot90(k=p)); the dense
// policy (visits / total visits, f32) turned the same way; the game's payoff repeated.  One CTA per game walks its
// plies in order with the running bitboards in shared memory; every thread writes a coalesced slice per ply.
__device__ __forceinline__ void kb_training_tensors(const uint16_t* __restrict__ hist, const uint32_t* __restrict__ pol_off,
                                                    const uint16_t* __restrict__ pol_tile,
                                                    const uint32_t* __restrict__ pol_visits, int plies,
                                                    const float* __restrict__ payoff4, float* __restrict__ states,
                                                    float* __restrict__ policies, float* __restrict__ values,
                                                    uint32_t* own /* shared [4][20] */, float* dense /* shared [400] */,
                                                    uint32_t* legal /* shared [20] */, int tid, int nthreads) {
    for (int i = tid; i < 80; i += nthreads) own[i] = 0u;
    __syncthreads();
    for (int ply = 0; ply < plies; ++ply) {
        const uint32_t h = hist[ply];
        const int p = int(h >> 9), tile = int(h & 0x1FFu);
        const uint32_t a = pol_off[ply], b = pol_off[ply + 1];
        for (int i = tid; i < 400; i += nthreads) dense[i] = 0.0f;
        for (int i = tid; i < 20; i += nthreads) legal[i] = 0u;
        __syncthreads();
        uint32_t total = 0u;
        for (uint32_t e = a; e < b; ++e) total += pol_visits[e];           // every thread: <= 400 cached loads
        const float ft = float(total);
        for (uint32_t e = a + uint32_t(tid); e < b; e += uint32_t(nthreads)) {
            const int t = int(pol_tile[e]);
            dense[t] = __fdiv_rn(float(pol_visits[e]), ft);                // simulation.rs:222
            atomicOr(&legal[t / 20], 1u << (t % 20));
        }
        __syncthreads();
        float* st = states + size_t(ply) * 2000;
        float* po = policies + size_t(ply) * 400;
        for (int e = tid; e < 2000; e += nthreads) {
            const int plane = e / 400, j = (e % 400) / 20, k = e % 20;
            int r, c;
            if (p == 0) { r = j; c = k; }
            else if (p == 1) { r = k; c = 19 - j; }
            else if (p == 2) { r = 19 - j; c = 19 - k; }
            else { r = 19 - k; c = j; }
            const uint32_t row = plane < 4 ? own[((plane + p) & 3) * 20 + r] : legal[r];
            st[e] = float((row >> c) & 1u);
            if (plane == 0) po[e] = dense[r * 20 + c];
        }
        for (int e = tid; e < 4; e += nthreads) values[size_t(ply) * 4 + e] = payoff4[e];
        __syncthreads();
        if (tid == 0) own[p * 20 + tile / 20] |= 1u << (tile % 20);       // "make the move that was made"
        __syncthreads();
    }
}
