// bk_mcts_pipe.cuh — the fixed-prior self-play search as a TWO-WARP PIPELINE per game (exact mode, small batches).
//
// One simulation of mcts() (simulation.rs:192-210) is a dependent chain: select down the tree (one memory round
// trip + a scored reduction per level), then Game::apply + move generation + expansion at the leaf, then the
// backup.  With 1024 games on 148 SMs there are 7 warps per SM and every cycle of that chain is exposed.  The two
// halves need each other less than it seems:
//
//   * the backup of simulation k needs the leaf's VALUE, and with the fixed-prior stub that is the constant 0.25 for
//     every seat unless the leaf is terminal — so it can be written as soon as the select has found the leaf;
//   * the select of simulation k+1 needs the leaf's EXPANSION only if it walks through that very leaf.
//
// So warp A (select + backup) runs one simulation ahead of warp B (apply + move generation + expansion):
// A hands leaf k to B, writes backup k with the stub value, and selects k+1 while B works.  Before A USES select
// k+1 it waits for B's verdict on leaf k; the speculation is void — and k+1 is selected again after the backup has
// been rewritten with the payoff — only if leaf k turned out terminal.  A select that arrives AT the leaf B is
// working on waits for B and starts over.  Results are bit-identical to the sequential kernel: every select reads
// exactly the statistics the sequential order would have shown it (tests/test_golden_configs.py, every ply of
// complete 800-sim games).
//
// A terminal leaf is remembered: B marks the entry (TN bit 22) and stores the winners' mask, so later visits —
// frequent near the end of a game — are backed up by A alone, with the payoff rebuilt from the mask (the reference
// clones, replays and re-scores the same position every time, simulation.rs:196-203; same numbers).
#pragma once
#include "bk_mcts_kernels.cuh"

#define BK_TN_TERMINAL(tn) (((tn) >> 22) & 1u)
#ifndef BK_PIPE_TAB_CAP
#define BK_PIPE_TAB_CAP 1024      // UCB factor tables up to this many entries are staged in shared memory
#endif

// -DBK_PIPE_STATS (probe builds only, tools/probe_mcts_pipe.py): cycles each warp spends waiting for the other go to
// counters[6] (A) / [7] (B), a game's total cycles to [8]; finer sums over all games to g_pipe_stats:
// 0 select cycles, 1 levels walked, 2 backup cycles, 3 B: state load + apply, 4 B: expansion + link, 5 selects, 6 voided selects
#if defined(BK_PIPE_STATS) && !defined(BK_WARP_EMU)
#define BK_PIPE_T0() const long long _t0 = clock64()
#define BK_PIPE_T1(acc) (acc) += (unsigned long long)(clock64() - _t0)
static __device__ unsigned long long g_pipe_stats[32];
#define BK_STAT_T0(name) const long long name = clock64()
#define BK_STAT_T1(name, slot) do { if (lane == 0) atomicAdd(&g_pipe_stats[slot], (unsigned long long)(clock64() - name)); } while (0)
#define BK_STAT_ADD(slot, v) do { if (lane == 0) atomicAdd(&g_pipe_stats[slot], (unsigned long long)(v)); } while (0)
#else
#define BK_PIPE_T0() do {} while (0)
#define BK_PIPE_T1(acc) do {} while (0)
#define BK_STAT_T0(name) do {} while (0)
#define BK_STAT_T1(name, slot) do {} while (0)
#define BK_STAT_ADD(slot, v) do {} while (0)
#endif

struct BkPathBuf {
    uint32_t e[BK_PATH_CAP];
    uint32_t n[BK_PATH_CAP];      // visits / value sum as the select read them (the backup's inputs, and the rollback's)
    uint32_t w[BK_PATH_CAP];
    uint8_t tp[BK_PATH_CAP];
};

struct BkPipeShared {
    volatile uint32_t req_seq, done_seq, quit, go;
    volatile uint32_t req_parent, req_entry, req_tile;
    volatile uint32_t res_kind, res_mask;          // 0 expanded, 1 terminal (res_mask = winners), 2 error
    volatile uint32_t err;
    volatile uint32_t n_nodes, n_entries;          // pool cursors, handed over at the phase barriers
};
#define BK_PIPE_OK 0u
#define BK_PIPE_TERMINAL 1u
#define BK_PIPE_ERROR 2u

// The tests' CPU emulator runs every lane as an OS thread: a polling warp must give its cores away there.
#ifdef BK_WARP_EMU
#include <chrono>
#include <thread>
#define BK_PIPE_POLL_PAUSE() std::this_thread::yield()
#else
#define BK_PIPE_POLL_PAUSE() do {} while (0)
#endif

__device__ __forceinline__ uint32_t bk_pipe_read(const volatile uint32_t* p, int lane) {   // warp-uniform read of a flag
    uint32_t v = 0u;
    if (lane == 0) v = *p;
    return __shfl_sync(BK_FULL, v, 0);
}

__device__ __forceinline__ void bk_payoff_from_mask(uint32_t mask, float (&pay)[4]) {       // game.rs:252-272
    const float share = __fdiv_rn(1.0f, float(__popc(mask & 0xFu)));
#pragma unroll
    for (int p = 0; p < 4; ++p) pay[p] = ((mask >> p) & 1u) ? share : 0.0f;
}

__device__ __forceinline__ uint32_t bk_winners_mask(const BkRegs& G) {
    int sc[4];
    bk_scores(G, sc);
    int hi = sc[0];
#pragma unroll
    for (int p = 1; p < 4; ++p) hi = sc[p] > hi ? sc[p] : hi;
    uint32_t m = 0u;
#pragma unroll
    for (int p = 0; p < 4; ++p) m |= (sc[p] == hi) ? (1u << p) : 0u;
    return m;
}

#define BK_SEL_LEAF 0
#define BK_SEL_KNOWN_TERMINAL 1
#define BK_SEL_HIT_PENDING 2
#define BK_SEL_ERROR 3

struct BkPipeLeaf {
    uint32_t parent, entry, mask;
    int tile, depth, kind;
};

// select_child loop (simulation.rs:198-203, :88-98, :135-147); same arithmetic as bk_tree_select<false>.
// `pending` = the entry warp B is expanding right now (BK_NODE_NONE if none): arriving there ends the walk.
// ucb / rcp: the two small tables of the UCB factor (cfg.ucb_tab, cfg.rcp_tab) — copies in shared memory when they fit
// (BK_PIPE_TAB_CAP): both loads sit on the select's dependent chain, and this kernel's L1 is too contended to serve them.
__device__ __forceinline__ float bk_pipe_u(const float* __restrict__ rcp, float F, uint32_t n_visits) {      // see bk_ucb_div
    if (rcp) {
        const float r = rcp[n_visits + 1u];
        const float d = float(n_visits + 1u);
        const float q0 = __fmul_rn(F, r);
        return __fmaf_rn(__fmaf_rn(-d, q0, F), r, q0);
    }
    return __fdiv_rn(F, __fadd_rn(1.0f, float(n_visits)));
}

__device__ __forceinline__ BkPipeLeaf bk_pipe_select(const BkTree& tr, const float* __restrict__ ucb, const float* __restrict__ rcp,
                                                     const BkBlock& root, uint32_t root_visits,
                                                     uint32_t pending, int lane, BkPathBuf& pb, uint32_t& err) {
    uint32_t node = 0u, off = root.off, Np = root_visits, e = 0u, tn = 0u, mask = 0u;
    int n = int(root.n), depth = 0, kind = BK_SEL_LEAF;
    for (;;) {
        const float F = ucb[Np];
        uint32_t wi, b_tn = 0u, b_n = 0u, b_w = 0u, b_off = 0u, b_node = 0u;
        if (n <= 32) {
            uint32_t key = 0u;
            if (lane < n) {
                const uint4 sv = tr.S[off + lane];
                const uint4 xv = tr.X[off + lane];
                const float u = bk_pipe_u(rcp, F, sv.x);
                const float sc = __fadd_rn(__fmul_rn(u, __uint_as_float(sv.z)), __uint_as_float(sv.y));
                if (sc >= 0.0f) key = __float_as_uint(sc) + 1u;
                b_tn = sv.w; b_n = sv.x; b_w = xv.x; b_off = xv.y; b_node = xv.z;
            }
            const uint32_t kmax = __reduce_max_sync(BK_FULL, key);
            if (kmax == 0u) { err |= BK_SP_ERR_NO_CHILD; kind = BK_SEL_ERROR; break; }
            wi = 31u - uint32_t(__clz(int(__ballot_sync(BK_FULL, key == kmax))));
        } else {
            float best = 0.0f;
            int bi = -1;
            for (int i = lane; i < n; i += 32) {
                const uint4 sv = tr.S[off + i];
                const float u = bk_pipe_u(rcp, F, sv.x);
                const float sc = __fadd_rn(__fmul_rn(u, __uint_as_float(sv.z)), __uint_as_float(sv.y));
                if (sc >= best) { best = sc; bi = i; b_tn = sv.w; b_n = sv.x; }
            }
            if (bi >= 0) {                       // the descent stream of this lane's best only (in flight while the warp reduces)
                const uint4 xv = tr.X[off + bi];
                b_w = xv.x; b_off = xv.y; b_node = xv.z;
            }
            const uint32_t key = bi >= 0 ? __float_as_uint(best) + 1u : 0u;
            const uint32_t kmax = __reduce_max_sync(BK_FULL, key);
            if (kmax == 0u) { err |= BK_SP_ERR_NO_CHILD; kind = BK_SEL_ERROR; break; }
            wi = __reduce_max_sync(BK_FULL, key == kmax ? uint32_t(bi) + 1u : 0u) - 1u;
        }
        const int src = int(wi & 31u);
        e = off + wi;
        if (e == pending) { kind = BK_SEL_HIT_PENDING; break; }      // B has not finished this leaf: its flags are not to be trusted
        tn = __shfl_sync(BK_FULL, b_tn, src);                       // the four broadcasts are issued back to back: the next
        const uint32_t w_n = __shfl_sync(BK_FULL, b_n, src);         // level's loads wait for one shuffle latency, not two
        const uint32_t w_off = __shfl_sync(BK_FULL, b_off, src);
        const uint32_t w_node = __shfl_sync(BK_FULL, b_node, src);
        if (lane == src) {
            const int slot = depth & (BK_PATH_CAP - 1);
            pb.e[slot] = e; pb.n[slot] = b_n; pb.w[slot] = b_w;
            pb.tp[slot] = uint8_t(BK_TN_TOPLAY(tn));
        }
        ++depth;
        if (!BK_TN_EXPANDED(tn)) {
            if (BK_TN_TERMINAL(tn)) { kind = BK_SEL_KNOWN_TERMINAL; mask = w_off; }
            break;
        }
        Np = w_n; off = w_off; node = w_node;
        n = int(BK_TN_NCHILD(tn));
    }
    if (depth > BK_PATH_CAP) { err |= BK_SP_ERR_PATH_CAP; kind = BK_SEL_ERROR; }
    __syncwarp();
    BkPipeLeaf lf;
    lf.parent = node; lf.entry = e; lf.mask = mask; lf.tile = int(BK_TN_TILE(tn)); lf.depth = depth; lf.kind = kind;
    return lf;
}

// backpropagate (simulation.rs:164-171) from the values the select recorded.  leaf_tp: seat whose value the LEAF entry
// receives (node.to_play: the mover at an expanded leaf, 0 at a terminal one — irrelevant when all four values are equal).
__device__ __forceinline__ void bk_pipe_backup(const BkTree& tr, int depth, const float (&val)[4], int leaf_tp, int lane,
                                               const BkPathBuf& pb) {
    for (int d = lane; d < depth; d += 32) {
        const uint32_t e = pb.e[d];
        const int tp = d == depth - 1 ? leaf_tp : int(pb.tp[d]);
        const uint32_t nv = pb.n[d] + 1u;
        const float w = __fadd_rn(__uint_as_float(pb.w[d]), bk_sel4f(tp, val[0], val[1], val[2], val[3]));
        tr.X[e].x = __float_as_uint(w);
        *reinterpret_cast<uint2*>(&tr.S[e]) = make_uint2(nv, __float_as_uint(__fdiv_rn(w, float(nv))));
    }
    __syncwarp();
}

// ---- warp B: the leaf half of a simulation --------------------------------------------------------------------------------
template <class SM>
__device__ __forceinline__ void bk_pipe_leaf_worker(const BkSearchCfg& cfg, const BkTree& tr, BkPipeShared& ps, int lane,
                                                    const BkTabs& tabs, SM& sm, BkCounters& gctr, BkSpCounters& ctr,
                                                    unsigned long long& waited) {
    BkSearchHdr hd;
    hd.n_nodes = ps.n_nodes; hd.n_entries = ps.n_entries; hd.err = 0u;
    uint32_t seq = 0u;
    for (;;) {
        uint32_t have;
        {
            BK_PIPE_T0();
            for (;;) {
                have = bk_pipe_read(&ps.req_seq, lane);
                if (have != seq) break;
                if (bk_pipe_read(&ps.quit, lane)) break;
                BK_PIPE_POLL_PAUSE();
            }
            BK_PIPE_T1(waited);
        }
        if (have == seq) break;                                     // quit, nothing outstanding
        __threadfence_block();
        const uint32_t parent = ps.req_parent, entry = ps.req_entry;
        const int tile = int(ps.req_tile);
        BkRegs L;
        BK_STAT_T0(tb0);
        bk_load(&tr.nodes[parent], lane, L);
        uint32_t kind = BK_PIPE_OK, mask = 0u;
        const bool applied = bk_apply(L, tile, -1, lane, tabs, gctr);
        BK_STAT_T1(tb0, 3);
        BK_STAT_T0(tb1);
        if (!applied) { hd.err |= BK_SP_ERR_APPLY; kind = BK_PIPE_ERROR; }
        else {
            if (lane == 0) ctr.applies += 1u;
            if (bk_terminal(L)) {                                   // simulation.rs:45-47; remembered in the entry
                mask = bk_winners_mask(L);
                if (lane == 0) { tr.X[entry].y = mask; tr.S[entry].w |= 1u << 22; }
                kind = BK_PIPE_TERMINAL;
            } else {
                BkBlock blk;
                const uint32_t id = bk_tree_expand(tr, hd, cfg, L, nullptr, lane, sm, ctr, blk);
                bk_tree_link(tr, entry, tile, id, bk_cur(L), blk, lane);                    // simulation.rs:78
                if (hd.err) kind = BK_PIPE_ERROR;
            }
        }
        __syncwarp();
        BK_STAT_T1(tb1, 4);
        __threadfence_block();                                      // the tree writes above are visible before the verdict
        seq += 1u;
        if (lane == 0) { ps.res_kind = kind; ps.res_mask = mask; if (hd.err) ps.err = ps.err | hd.err; __threadfence_block(); ps.done_seq = seq; }
        __syncwarp();
        if (kind == BK_PIPE_ERROR) {                                // keep answering so that A never waits in vain
            for (;;) {
                const uint32_t h2 = bk_pipe_read(&ps.req_seq, lane);
                if (h2 != seq) { seq = h2; if (lane == 0) { ps.res_kind = BK_PIPE_ERROR; __threadfence_block(); ps.done_seq = seq; } __syncwarp(); }
                if (bk_pipe_read(&ps.quit, lane)) break;
            }
            break;
        }
    }
    if (lane == 0) { ps.n_nodes = hd.n_nodes; ps.n_entries = hd.n_entries; }
}

// ---- warp A: select + backup, one simulation ahead of B -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t bk_pipe_wait(BkPipeShared& ps, uint32_t seq, int lane, unsigned long long& waited) {
    BK_PIPE_T0();
    while (bk_pipe_read(&ps.done_seq, lane) != seq) { BK_PIPE_POLL_PAUSE(); }
    BK_PIPE_T1(waited);
    __threadfence_block();
    return bk_pipe_read(&ps.res_kind, lane);
}

// runs cfg.sims simulations of one ply; returns the error flags raised on this side
__device__ __forceinline__ uint32_t bk_pipe_search(const BkSearchCfg& cfg, const BkTree& tr, const BkBlock& root, BkPipeShared& ps,
                                                   int lane, BkPathBuf (&pbs)[2], BkSpCounters& ctr, unsigned long long& waited,
                                                   const float* __restrict__ ucb, const float* __restrict__ rcp) {
    uint32_t err = 0u, started = 0u, seq = 0u, pending = BK_NODE_NONE;
    int cur = 0, pend_depth = 0;
    bool inflight = false;
    float stub[4] = {cfg.stub_value, cfg.stub_value, cfg.stub_value, cfg.stub_value};
    // B's verdict on the leaf in flight; a terminal leaf voids the speculative backup: rewrite it with the payoff
    auto settle = [&](uint32_t kind) {
        inflight = false;
        pending = BK_NODE_NONE;
        if (kind == BK_PIPE_TERMINAL) {
            float pay[4];
            bk_payoff_from_mask(bk_pipe_read(&ps.res_mask, lane), pay);
            bk_pipe_backup(tr, pend_depth, pay, 0, lane, pbs[cur ^ 1]);
        }
    };
    for (;;) {
        if (started == cfg.sims) {
            if (inflight) { const uint32_t k = bk_pipe_wait(ps, seq, lane, waited); settle(k); if (k == BK_PIPE_ERROR) err |= BK_SP_ERR_ENTRY_CAP; }
            break;
        }
        BK_STAT_T0(ts0);
        const BkPipeLeaf lf = bk_pipe_select(tr, ucb, rcp, root, started + 1u, pending, lane, pbs[cur], err);   // root visits: :194
        BK_STAT_T1(ts0, 0);
        BK_STAT_ADD(1, lf.depth);
        BK_STAT_ADD(5, 1);
        if (lf.kind == BK_SEL_HIT_PENDING) {                        // walked into B's leaf: wait for it, select again
            const uint32_t k = bk_pipe_wait(ps, seq, lane, waited);
            settle(k);
            if (k == BK_PIPE_ERROR) { err |= BK_SP_ERR_ENTRY_CAP; break; }
            BK_STAT_ADD(6, 1);
            continue;
        }
        if (inflight) {
            const uint32_t k = bk_pipe_wait(ps, seq, lane, waited);
            settle(k);
            if (k == BK_PIPE_ERROR) { err |= BK_SP_ERR_ENTRY_CAP; break; }
            if (k == BK_PIPE_TERMINAL) { BK_STAT_ADD(6, 1); continue; }   // the select above saw a backup that never happened
        }
        if (lf.kind == BK_SEL_ERROR) break;
        if (lf.kind == BK_SEL_KNOWN_TERMINAL) {                     // payoff already known: no work for B
            float pay[4];
            bk_payoff_from_mask(lf.mask, pay);
            bk_pipe_backup(tr, lf.depth, pay, 0, lane, pbs[cur]);
            started += 1u;
            continue;
        }
        if (lane == 0) {
            ps.req_parent = lf.parent; ps.req_entry = lf.entry; ps.req_tile = uint32_t(lf.tile);
            __threadfence_block();
            ps.req_seq = seq + 1u;
        }
        seq += 1u;
        inflight = true;
        pending = lf.entry;
        pend_depth = lf.depth;
        BK_STAT_T0(tk0);
        bk_pipe_backup(tr, lf.depth, stub, 0, lane, pbs[cur]);      // speculative: right unless the leaf is terminal
        BK_STAT_T1(tk0, 2);
        started += 1u;
        cur ^= 1;
    }
    if (lane == 0) { ctr.sims += started; ps.quit = 1u; }
    __syncwarp();
    return err;
}

// training_game() (simulation.rs:267-296) with the stub evaluator on the two-warp pipeline.  Exact mode only
// (the throughput modes keep the one-warp kernel).  Called by all 64 threads of the game's CTA.
template <class SM>
__device__ __forceinline__ void kb_selfplay_stub_pipe(const BkSearchCfg& cfg, BkState* __restrict__ states, uint16_t* __restrict__ hist,
                                                      const BkTree& tr, BkSearchHdr* hdr_g, uint32_t* pol_off, uint16_t* pol_tile,
                                                      uint32_t* pol_visits, int max_plies, unsigned long long* counters, int g,
                                                      int warp, int lane, const BkTabs& tabs, SM& sm,
                                                      BkPathBuf (&pbs)[2], BkPipeShared& ps, const float* __restrict__ ucb,
                                                      const float* __restrict__ rcp) {
    BkRegs G;
    BkSearchHdr hd;
    BkCounters gctr = {0u, 0u};
    BkSpCounters ctr = {0u, 0u, 0u, 0u};
    unsigned long long waited = 0ull;
    const uint32_t game_id = cfg.first_game_id + uint32_t(g);
    int plies = 0;
    BkBlock root;
    root.off = 0u; root.n = 0u;
    BK_PIPE_T0();
    if (warp == 0) {
        bk_load(&states[g], lane, G);
        hd.err = hdr_g->err; hd.pol_count = hdr_g->pol_count; hd.plies_searched = hdr_g->plies_searched;
        hd.forced_plies = hdr_g->forced_plies; hd.reused = 0u;
        hd.n_nodes = 0u; hd.n_entries = 0u; hd.root_visits = 0u; hd.sims_done = 0u;
        if (lane == 0) ps.err = 0u;
    }
    for (;;) {
        if (warp == 0) {
            const bool go = !bk_terminal(G) && (max_plies < 0 || plies < max_plies) && hd.err == 0u;
            if (go) {
                hd.n_nodes = 0u; hd.n_entries = 0u;                                             // fresh tree, simulation.rs:183
                bk_tree_expand(tr, hd, cfg, G, nullptr, lane, sm, ctr, root);                   // evaluate(root), :184 (B idles at the barrier: its scratch is free)
                bk_tree_noise(tr, cfg, game_id, G.ply, lane);                                   // :190
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {
                ps.n_nodes = hd.n_nodes; ps.n_entries = hd.n_entries;
                ps.req_seq = 0u; ps.done_seq = 0u; ps.quit = 0u; ps.res_kind = 0u;
                ps.go = (go && hd.err == 0u) ? 1u : 0u;
            }
        }
        __syncthreads();
        if (!ps.go) break;
        if (warp == 0) {
            hd.err |= bk_pipe_search(cfg, tr, root, ps, lane, pbs, ctr, waited, ucb, rcp);
        } else {
            bk_pipe_leaf_worker(cfg, tr, ps, lane, tabs, sm, gctr, ctr, waited);
        }
        __syncthreads();
        if (warp == 0) {
            hd.err |= ps.err;
            hd.n_nodes = ps.n_nodes; hd.n_entries = ps.n_entries;
            hd.root_visits = cfg.sims; hd.sims_done = cfg.sims;
            if (hd.err == 0u) {
                const int action = bk_tree_finish_ply(tr, hd, cfg, game_id, G.ply, pol_off, pol_tile, pol_visits, lane, nullptr);
                const int p = bk_cur(G);
                const uint32_t ply = G.ply;
                if (!bk_apply(G, action, -1, lane, tabs, gctr)) hd.err |= BK_SP_ERR_APPLY;            // :288
                else if (lane == 0 && ply < BK_HIST_CAP) hist[size_t(g) * BK_HIST_CAP + ply] = uint16_t(action | (p << 9));
                ++plies;
            }
        }
    }
    if (warp == 0) {
        bk_store(&states[g], lane, G);
        if (lane == 0) {
            hdr_g->err = hd.err; hdr_g->pol_count = hd.pol_count; hdr_g->plies_searched = hd.plies_searched;
            hdr_g->n_nodes = hd.n_nodes; hdr_g->n_entries = hd.n_entries; hdr_g->root_visits = hd.root_visits;
            hdr_g->sims_done = hd.sims_done; hdr_g->forced_plies = hd.forced_plies; hdr_g->reused = 0u;
            hdr_g->pend_kind = 0u;
        }
    }
    const unsigned crem = __reduce_add_sync(BK_FULL, gctr.crem);
    if (lane == 0 && counters) {
        atomicAdd(&counters[0], (unsigned long long)ctr.sims);
        atomicAdd(&counters[1], (unsigned long long)ctr.applies + (unsigned long long)plies);
        atomicAdd(&counters[2], (unsigned long long)gctr.movegens);
        atomicAdd(&counters[3], 120ull * (unsigned long long)crem);
        atomicAdd(&counters[4], (unsigned long long)ctr.entries);
        atomicAdd(&counters[5], (unsigned long long)ctr.nodes);
#if defined(BK_PIPE_STATS) && !defined(BK_WARP_EMU)
        unsigned long long total = 0ull;
        BK_PIPE_T1(total);
        atomicAdd(&counters[6 + warp], waited);
        if (warp == 0) atomicAdd(&counters[8], total);
#endif
    }
}
