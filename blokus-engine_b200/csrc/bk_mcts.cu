// bk_mcts.cu — MCTS self-play clients: kernel wrappers and the bk_selfplay_* C ABI
// (include/blokus_b200.h).  Host mirror of self_play/src/lib.rs:9-32 + simulation.rs:267-296, batched.
#include <math.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <vector>

#include "bk_host.h"
#include "bk_mcts_kernels.cuh"
#include "bk_mcts_pipe.cuh"
#include "bk_eval_kernels.cuh"

struct bk_selfplay {
    int n = 0;
    int device = 0;
    bk_config cfg{};
    uint32_t first_id = 0;
    bk_env* env = nullptr;
    BkSearchCfg dcfg{};
    uint4* d_S = nullptr;
    uint4* d_X = nullptr;
    BkState* d_nodes = nullptr;
    double* d_scratch = nullptr;
    BkSearchHdr* d_hdr = nullptr;
    BkPend* d_pend = nullptr;         // [n][pend_cap], multi-leaf mode only
    uint32_t pend_cap = 0;            // leaves per round d_pend was allocated for
    uint32_t* d_remap = nullptr;      // [n][2 * max_nodes], tree-reuse mode only
    uint32_t* d_slot_base = nullptr;  // [n + 1]: first dense evaluator row of each game's outstanding positions (+ total)
    bool use_vl = false;
    int num_sms = 148;
    int pipe_resident = 7;            // CTAs of k_selfplay_stub_pipe an SM holds (occupancy API at create)
    int full_resident = 8;            // same for the all-registers one-warp instantiations k_selfplay_stub<1, *>
    int stub_min_blocks = 0;          // 0 = choose by batch size; BK_STUB_MIN_BLOCKS in the environment overrides (probes)
    int stub_pipe = -1;               // -1 = two-warp pipeline for small exact-mode batches; BK_STUB_PIPE=0/1 forces it off/on
    uint32_t* d_pol_off = nullptr;    // [n][BK_HIST_CAP + 1]
    uint16_t* d_pol_tile = nullptr;   // [n][policy_cap]
    uint32_t* d_pol_visits = nullptr; // [n][policy_cap]
    float* d_ucb = nullptr;
    float* d_rcp = nullptr;           // RN(1/d) table of bk_ucb_div
    bool short_div = false;           // the short division was verified exhaustively for this configuration
    float* d_prior = nullptr;
    unsigned long long* d_counters = nullptr;  // [16], cumulative ([6..] are used by probe builds only)
    uint8_t* d_stage = nullptr;       // [n][400 * 16] gather staging for last_root
    int32_t* d_round = nullptr;       // [4] per-round scalars of the fused network loop
    int64_t* d_ply_off = nullptr;     // [n + 1] prefix of plies per game (training tensors)
    int64_t* d_pack_off = nullptr;    // [2][n + 1] prefixes of plies / policy entries per game (packed results)
    int64_t* d_pack_ptr = nullptr;    // [pack_plies + 1]
    uint16_t* d_pack_tile = nullptr;  // [pack_entries]
    uint32_t* d_pack_visits = nullptr;
    int64_t pack_plies = -1, pack_entries = -1, pack_cap_plies = 0, pack_cap_entries = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.0f;
};

// ---- kernels ------------------------------------------------------------------------------------------
__global__ void k_sp_summary(const BkState* __restrict__ states, BkSummary* __restrict__ out, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_summary(states, out, g, lane);
}

struct BkPools {
    uint4* S; uint4* X; BkState* nodes; double* scratch; BkSearchHdr* hdr; BkPend* pend; uint32_t* remap; uint32_t* slot_base;
    uint32_t* pol_off; uint16_t* pol_tile; uint32_t* pol_visits;
};

__device__ __forceinline__ BkTree bk_tree_of(const BkPools& pl, const BkSearchCfg& cfg, int g) {
    BkTree t;
    const size_t eo = size_t(g) * cfg.entry_cap;
    t.S = pl.S + eo; t.X = pl.X + eo;
    t.nodes = pl.nodes + size_t(g) * cfg.max_nodes;
    t.scratch = pl.scratch + size_t(g) * 400;
    t.remap = pl.remap ? pl.remap + size_t(g) * 2 * cfg.max_nodes : nullptr;
    return t;
}

// MINB = CTAs (= games) the register allocation must let an SM hold.  One simulation is a long dependent
// chain, so a small batch (<= 12 games per SM) runs fastest with all the registers ptxas wants (166), while a
// large batch gains more from residency: 126 registers at 16 games/SM, 96 at 20 (measured on B200, 8192 games:
// 2.55e8 -> 3.84e8 -> 4.31e8 sims/s; 1024 games: 1.79e8 -> 1.69e8 -> 1.47e8; 24 and 32 per SM spill and are slower:
// profiles/r01_ab_mcts_occupancy.log).  The host picks per batch size.
template <int MINB, bool MODES>
__global__ void __launch_bounds__(32, MINB)
k_selfplay_stub(BkSearchCfg cfg, BkPools pl, BkState* states, uint16_t* hist, int n, int max_plies,
                unsigned long long* counters) {
    // candidate window masks read in place (global memory through L1): the 5 KB per-CTA copy capped residency at 24 games
    // per SM; without it 28 fit and the 8192-game batch runs 30 % faster (profiles/r02_ab_mcts_stub_global_cands.log)
    __shared__ uint16_t cand_scratch[32];
    __shared__ BkWarpSmemStub wsm;
    const BkTabs tabs = bk_global_tables(cand_scratch);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x;
    if (g >= n) return;
    const BkTree tr = bk_tree_of(pl, cfg, g);
    kb_selfplay_stub<(MINB < 16), MODES>(cfg, states, hist, tr, &pl.hdr[g], pl.pol_off + size_t(g) * (BK_HIST_CAP + 1),
                     pl.pol_tile + size_t(g) * cfg.policy_cap, pl.pol_visits + size_t(g) * cfg.policy_cap, max_plies,
                     counters, g, lane, tabs, wsm);
}

// Exact mode, small batches: two warps per game — warp 0 selects and backs up one simulation ahead of warp 1, which applies,
// generates moves and expands (bk_mcts_pipe.cuh).  Same results as k_selfplay_stub bit for bit.
// 7 resident games per SM (1024 games on 148 SMs are 6.9 per SM): 144 registers per thread.  With the 195 the compiler would
// like, only 5 CTAs fit an SM and 1024 games run in two waves (measured: 1.55 s against 1.26 s for the one-warp kernel).
__global__ void __launch_bounds__(64, 7)
k_selfplay_stub_pipe(BkSearchCfg cfg, BkPools pl, BkState* states, uint16_t* hist, int n, int max_plies, unsigned long long* counters) {
    // candidate window masks straight from global memory (L1): without the 5 KB per-CTA copy the 1024-game run is 2 %
    // faster (profiles/r02_ab_mcts_global_cands.log)
    __shared__ uint16_t cand_scratch[32 * 2];
    __shared__ BkWarpSmemStub wsm;
    __shared__ BkPathBuf pbs[2];
    __shared__ BkPipeShared ps;
    __shared__ float s_ucb[BK_PIPE_TAB_CAP], s_rcp[BK_PIPE_TAB_CAP];
    const BkTabs tabs = bk_global_tables(cand_scratch);
    // the UCB factor tables (sims + 2 and sims + 3 floats) in shared memory when they fit
    const bool tabs_fit = cfg.sims + 3u <= BK_PIPE_TAB_CAP;
    if (tabs_fit)
        for (uint32_t i = threadIdx.x; i < cfg.sims + 3u; i += blockDim.x) {
            s_ucb[i] = i < cfg.sims + 2u ? cfg.ucb_tab[i] : 0.0f;
            s_rcp[i] = cfg.rcp_tab ? cfg.rcp_tab[i] : 0.0f;
        }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    if (g >= n) return;
    const BkTree tr = bk_tree_of(pl, cfg, g);
    const float* ucb = tabs_fit ? s_ucb : cfg.ucb_tab;
    const float* rcp = cfg.rcp_tab ? (tabs_fit ? s_rcp : cfg.rcp_tab) : nullptr;
    kb_selfplay_stub_pipe(cfg, states, hist, tr, &pl.hdr[g], pl.pol_off + size_t(g) * (BK_HIST_CAP + 1),
                          pl.pol_tile + size_t(g) * cfg.policy_cap, pl.pol_visits + size_t(g) * cfg.policy_cap, max_plies, counters, g,
                          warp, lane, tabs, wsm, pbs, ps, ucb, rcp);
}

__global__ void __launch_bounds__(32 * 4) k_sp_begin(BkSearchCfg cfg, BkPools pl, const BkState* states, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_sp_begin(cfg, states, bk_tree_of(pl, cfg, g), &pl.hdr[g], g, lane);
}

// The evaluator batch is DENSE in every mode — game g's outstanding positions (one in the exact mode, up to
// leaves_per_round in the multi-leaf mode) occupy rows slot_base[g] .. slot_base[g + 1) in (game, slot) order,
// slot_base[n] = rows to evaluate — so the evaluator never works on finished games or empty slots.  One CTA, chunked scan.
__global__ void __launch_bounds__(256) k_sp_slot_scan(BkPools pl, int n, int vl, int32_t* __restrict__ round_out) {
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t carry;
    __shared__ uint32_t waiting;                     // games that need the next expand_backup call (incl. kept trees resuming)
    if (threadIdx.x == 0) { carry = 0u; waiting = 0u; }
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 256) {
        const int g = c0 + int(threadIdx.x);
        uint32_t cnt = 0u;
        if (g < n) {
            const uint32_t k = pl.hdr[g].pend_kind;
            cnt = k == BK_PEND_ROOT ? 1u : (k == BK_PEND_LEAF ? (vl ? pl.hdr[g].pend_count : 1u) : 0u);
            if (k == BK_PEND_ROOT || k == BK_PEND_LEAF || k == BK_PEND_RESUME) atomicAdd(&waiting, 1u);
        }
        uint32_t incl = cnt;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        uint32_t before = carry;
        for (int i = 0; i < w; ++i) before += warp_tot[i];
        if (g < n) pl.slot_base[g] = before + incl - cnt;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pl.slot_base[n] = carry;
        if (round_out) { round_out[0] = int32_t(carry); round_out[1] = int32_t(waiting); }   // rows to evaluate, games waiting
    }
}

// the pending positions' planes written straight into the evaluator's first-layer input (padded NHWC bf16), dense rows
__global__ void __launch_bounds__(256) k_sp_planes_nhwc(BkSearchCfg cfg, BkPools pl, int n, int vl, uint16_t* __restrict__ x64) {
    const int K = int(cfg.leaves_per_round);
    const int g = blockIdx.x / K, j = blockIdx.x % K;
    if (g >= n) return;
    const BkSearchHdr* h = &pl.hdr[g];
    bool pend = h->pend_kind == BK_PEND_ROOT || h->pend_kind == BK_PEND_LEAF;
    uint32_t slot = h->n_nodes;
    if (vl) {
        if (h->pend_kind == BK_PEND_ROOT) { pend = j == 0; slot = 0u; }
        else if (h->pend_kind == BK_PEND_LEAF) {
            pend = uint32_t(j) < h->pend_count;
            if (pend) slot = pl.pend[size_t(g) * cfg.leaves_per_round + j].slot;
        }
    }
    if (!pend) return;
    const size_t row = size_t(pl.slot_base[g]) + size_t(j);
    kb_planes_nhwc(&bk_tree_of(pl, cfg, g).nodes[slot], reinterpret_cast<uint4*>(x64 + row * BK_PAD_IMAGE * BK_EVAL_IN_CH), threadIdx.x,
                   blockDim.x);
}

// planes of every game's pending position, float32 [n][5][20][20]; counts pending games
__global__ void k_sp_planes(BkSearchCfg cfg, BkPools pl, int n, int vl, float* __restrict__ out,
                            int32_t* __restrict__ pending) {
    const int K = int(cfg.leaves_per_round);
    const int g = blockIdx.x / K, j = blockIdx.x % K;   // game, leaf slot of the round (K = 1 in the exact mode)
    if (g >= n) return;
    const BkSearchHdr* h = &pl.hdr[g];
    bool pend = h->pend_kind == BK_PEND_ROOT || h->pend_kind == BK_PEND_LEAF;
    const bool resume = h->pend_kind == BK_PEND_RESUME;   // kept tree: needs a step call but has no position to evaluate
    uint32_t slot = h->n_nodes;                      // exact mode: the tentative node slot
    if (vl) {
        if (h->pend_kind == BK_PEND_ROOT) { pend = j == 0; slot = 0u; }
        else if (h->pend_kind == BK_PEND_LEAF) {
            pend = uint32_t(j) < h->pend_count;
            if (pend) slot = pl.pend[size_t(g) * cfg.leaves_per_round + j].slot;
        }
    }
    // dense rows: nothing is written for games / slots without a position waiting
    if (pend) kb_planes<float>(&bk_tree_of(pl, cfg, g).nodes[slot], out + size_t(pl.slot_base[g] + uint32_t(j)) * 2000, threadIdx.x, blockDim.x);
    if ((pend || resume) && threadIdx.x == 0 && j == 0) atomicAdd(pending, 1);
}

__global__ void k_sp_count(BkPools pl, int n, int32_t* __restrict__ counts) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint32_t k = pl.hdr[g].pend_kind;
    if (k == BK_PEND_ROOT || k == BK_PEND_LEAF || k == BK_PEND_RESUME) atomicAdd(&counts[0], 1);
    if (k == BK_PEND_DONE) atomicAdd(&counts[1], 1);
}

__global__ void __launch_bounds__(32)
k_sp_step(BkSearchCfg cfg, BkPools pl, int n, const float* policy, const float* value, unsigned long long* counters) {
    __shared__ uint16_t cand_scratch[32];
    __shared__ BkWarpSmem wsm;
    const BkTabs tabs = bk_global_tables(cand_scratch);     // one simulation per launch: a 5 KB table copy per game would cost more than it saves
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x;
    if (g >= n) return;
    const uint32_t row = pl.slot_base[g];            // this game's dense evaluator row (valid while its position is out)
    kb_sp_step(cfg, bk_tree_of(pl, cfg, g), &pl.hdr[g], policy + size_t(row) * 400, value + size_t(row) * 4, counters, g, lane,
               tabs, wsm);
}

__global__ void __launch_bounds__(32)
k_sp_step_vl(BkSearchCfg cfg, BkPools pl, int n, const float* policy, const float* value, unsigned long long* counters) {
    __shared__ uint16_t cand_scratch[32];
    __shared__ BkWarpSmem wsm;
    const BkTabs tabs = bk_global_tables(cand_scratch);     // one simulation per launch: a 5 KB table copy per game would cost more than it saves
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x;
    if (g >= n) return;
    const uint32_t row0 = pl.slot_base[g];           // this game's first dense evaluator row (valid while it has leaves out)
    kb_sp_step_vl(cfg, bk_tree_of(pl, cfg, g), &pl.hdr[g], pl.pend + size_t(g) * cfg.leaves_per_round, policy + size_t(row0) * 400,
                  value + size_t(row0) * 4, counters, g, lane, tabs, wsm);
}

__global__ void __launch_bounds__(32)
k_sp_end(BkSearchCfg cfg, BkPools pl, BkState* states, uint16_t* hist, int n, unsigned long long* counters) {
    __shared__ uint16_t cand_scratch[32];
    __shared__ BkWarpSmem wsm;
    const BkTabs tabs = bk_global_tables(cand_scratch);     // one simulation per launch: a 5 KB table copy per game would cost more than it saves
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x;
    if (g >= n) return;
    kb_sp_end(cfg, states, hist, bk_tree_of(pl, cfg, g), &pl.hdr[g], pl.pol_off + size_t(g) * (BK_HIST_CAP + 1),
              pl.pol_tile + size_t(g) * cfg.policy_cap, pl.pol_visits + size_t(g) * cfg.policy_cap, counters, g, lane, tabs,
              wsm);
}

__global__ void __launch_bounds__(256)
k_training_tensors(BkSearchCfg cfg, BkPools pl, const uint16_t* hist, const BkSummary* summary, const int64_t* ply_off,
                   int n, float* states, float* policies, float* values) {
    __shared__ uint32_t own[80];
    __shared__ float dense[400];
    __shared__ uint32_t legal[20];
    const int g = blockIdx.x;
    if (g >= n) return;
    const int64_t o = ply_off[g];
    kb_training_tensors(hist + size_t(g) * BK_HIST_CAP, pl.pol_off + size_t(g) * (BK_HIST_CAP + 1),
                        pl.pol_tile + size_t(g) * cfg.policy_cap, pl.pol_visits + size_t(g) * cfg.policy_cap,
                        int(ply_off[g + 1] - o), summary[g].payoff, states + o * 2000, policies + o * 400, values + o * 4,
                        own, dense, legal, threadIdx.x, blockDim.x);
}

// pack every game's policy records into one CSR: ply_ptr[ply_off[g] + k] = first entry of ply k of game g in the
// packed arrays (plus one closing entry at ply_ptr[total_plies]), tiles / visits copied game after game
__global__ void k_pack_results(BkSearchCfg cfg, BkPools pl, int n, const int64_t* __restrict__ ply_off,
                               const int64_t* __restrict__ ent_off, int64_t* __restrict__ ply_ptr,
                               uint16_t* __restrict__ tile, uint32_t* __restrict__ visits) {
    const int g = blockIdx.x;
    if (g >= n) return;
    const uint32_t plies = pl.hdr[g].plies_searched, cnt = pl.hdr[g].pol_count;
    const uint32_t* off = pl.pol_off + size_t(g) * (BK_HIST_CAP + 1);
    const uint16_t* st = pl.pol_tile + size_t(g) * cfg.policy_cap;
    const uint32_t* sv = pl.pol_visits + size_t(g) * cfg.policy_cap;
    const int64_t p0 = ply_off[g], e0 = ent_off[g];
    for (uint32_t k = threadIdx.x; k < plies; k += blockDim.x) ply_ptr[p0 + k] = e0 + int64_t(off[k]);
    for (uint32_t e = threadIdx.x; e < cnt; e += blockDim.x) { tile[e0 + e] = st[e]; visits[e0 + e] = sv[e]; }
    if (g == n - 1 && threadIdx.x == 0) ply_ptr[ply_off[n]] = ent_off[n];
}

// gather the root's child block (tile, visits, value_sum, prior) of every game: out[g][400] x 4 arrays
__global__ void k_last_root(BkSearchCfg cfg, BkPools pl, int n, int32_t* counts, int16_t* tile, uint32_t* visits,
                            float* wsum, float* prior) {
    const int g = blockIdx.x;
    if (g >= n) return;
    const BkTree tr = bk_tree_of(pl, cfg, g);
    const bool have = pl.hdr[g].n_nodes > 0u;
    const uint32_t off = have ? tr.nodes[0].pad[0] : 0u;
    const int cnt = have ? int(tr.nodes[0].pad[1]) : 0;
    if (threadIdx.x == 0) counts[g] = cnt;
    for (int i = threadIdx.x; i < 400; i += blockDim.x) {
        const bool in = i < cnt;
        const uint4 sv = in ? tr.S[off + i] : make_uint4(0u, 0u, 0u, 0u);
        tile[size_t(g) * 400 + i] = in ? int16_t(BK_TN_TILE(sv.w)) : int16_t(-1);
        visits[size_t(g) * 400 + i] = sv.x;
        wsum[size_t(g) * 400 + i] = in ? __uint_as_float(tr.X[off + i].x) : 0.0f;
        prior[size_t(g) * 400 + i] = __uint_as_float(sv.z);
    }
}

// ---- host side ------------------------------------------------------------------------------------------
static float host_exp_f32(float x) { return float(std::exp(double(x))); }  // same definition as bk_exp_f32

static BkPools pools_of(const bk_selfplay* sp) {
    BkPools p;
    p.S = sp->d_S; p.X = sp->d_X; p.nodes = sp->d_nodes; p.scratch = sp->d_scratch;
    p.hdr = sp->d_hdr; p.pend = sp->d_pend; p.remap = sp->d_remap; p.slot_base = sp->d_slot_base; p.pol_off = sp->d_pol_off; p.pol_tile = sp->d_pol_tile; p.pol_visits = sp->d_pol_visits;
    return p;
}

static int sp_use(const bk_selfplay* sp) {
    if (!sp) return bk_fail(BK_ERR_INVALID_ARG, "null bk_selfplay handle");
    BK_CUDA(cudaSetDevice(sp->device));
    return BK_OK;
}

static int sp_check_errors(bk_selfplay* sp) {
    std::vector<BkSearchHdr> h(size_t(sp->n));
    BK_CUDA(cudaMemcpyAsync(h.data(), sp->d_hdr, sizeof(BkSearchHdr) * h.size(), cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    for (int g = 0; g < sp->n; ++g) {
        const uint32_t e = h[size_t(g)].err;
        if (!e) continue;
        char buf[160];
        snprintf(buf, sizeof buf, "self-play game %d stopped: flags 0x%x (%s%s%s%s%s)", g, e,
                 (e & BK_SP_ERR_ENTRY_CAP) ? "child/node pool full " : "", (e & BK_SP_ERR_PATH_CAP) ? "path too deep " : "",
                 (e & BK_SP_ERR_NO_CHILD) ? "no selectable child " : "", (e & BK_SP_ERR_POLICY_CAP) ? "policy pool full " : "",
                 (e & BK_SP_ERR_APPLY) ? "illegal tile " : "");
        return bk_fail((e & (BK_SP_ERR_ENTRY_CAP | BK_SP_ERR_POLICY_CAP | BK_SP_ERR_PATH_CAP)) ? BK_ERR_CAPACITY : BK_ERR_STATE, buf);
    }
    return BK_OK;
}

static int selfplay_alloc(bk_selfplay* sp, const bk_config* cfg, uint32_t first_game_id, uint32_t max_children_per_game);

extern "C" {

int bk_selfplay_create(int n_games, int device, const bk_config* cfg, uint32_t first_game_id,
                       uint32_t max_children_per_game, bk_selfplay** out) {
    if (!cfg || !out || n_games <= 0) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_create: bad argument");
    if (cfg->sims_per_move == 0 || cfg->sims_per_move > 1000000u)
        return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_create: sims_per_move must be in 1..1000000");
    if (!(cfg->c_base > 0.0f)) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_create: c_base must be > 0");
    if (!(cfg->dirichlet_alpha > 0.0f)) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_create: dirichlet_alpha must be > 0");
    bk_selfplay* sp = new bk_selfplay();
    sp->n = n_games;
    sp->device = device;
    sp->cfg = *cfg;
    sp->first_id = first_game_id;
#ifndef BK_WARP_EMU
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0) sp->num_sms = prop.multiProcessorCount;
        if (const char* e = getenv("BK_STUB_MIN_BLOCKS")) sp->stub_min_blocks = atoi(e);
        int resident = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_selfplay_stub_pipe, 64, 0) == cudaSuccess && resident > 0)
            sp->pipe_resident = resident;
        int r0 = 0, r1 = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r0, k_selfplay_stub<1, false>, 32, 0) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r1, k_selfplay_stub<1, true>, 32, 0) == cudaSuccess && r0 > 0 && r1 > 0)
            sp->full_resident = r0 < r1 ? r0 : r1;
    }
#endif
#ifdef BK_WARP_EMU
    sp->stub_pipe = 0;       // the CPU emulator pays a futex per mailbox poll: the pipeline runs there only when a test asks for it
#endif
    if (const char* e = getenv("BK_STUB_PIPE")) sp->stub_pipe = atoi(e);     // probes / tests: 0 = one-warp kernel, 1 = pipeline
    int rc = bk_env_create(n_games, device, &sp->env);
    if (rc) { delete sp; return rc; }
    rc = selfplay_alloc(sp, cfg, first_game_id, max_children_per_game);
    if (rc) { bk_selfplay_destroy(sp); return rc; }  // e.g. cudaMalloc of the pools failed: release what exists
    *out = sp;
    return BK_OK;
}

}  // extern "C"

static int selfplay_alloc(bk_selfplay* sp, const bk_config* cfg, uint32_t first_game_id, uint32_t max_children_per_game) {
    const int n_games = sp->n;
    BkSearchCfg& d = sp->dcfg;
    d.sims = cfg->sims_per_move;
    d.sample_moves = cfg->sample_moves;
    d.frac = cfg->exploration_fraction;
    d.alpha = cfg->dirichlet_alpha;
    d.seed = cfg->seed;
    d.first_game_id = first_game_id;
    d.max_nodes = cfg->sims_per_move + 2u;
    d.entry_cap = max_children_per_game ? max_children_per_game : (cfg->sims_per_move + 2u) * 128u;
    d.policy_cap = 32768u;
    d.mode = 0u;
    d.leaves_per_round = 1u;
    d.stub_value = 0.25f;
    const size_t ne = size_t(n_games) * d.entry_cap;
    BK_CUDA(cudaMalloc(&sp->d_S, sizeof(uint4) * ne));
    BK_CUDA(cudaMalloc(&sp->d_X, sizeof(uint4) * ne));
    BK_CUDA(cudaMalloc(&sp->d_nodes, sizeof(BkState) * size_t(n_games) * d.max_nodes));
    BK_CUDA(cudaMalloc(&sp->d_scratch, sizeof(double) * 400 * size_t(n_games)));
    BK_CUDA(cudaMalloc(&sp->d_hdr, sizeof(BkSearchHdr) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&sp->d_pol_off, sizeof(uint32_t) * (BK_HIST_CAP + 1) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&sp->d_pol_tile, sizeof(uint16_t) * size_t(d.policy_cap) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&sp->d_pol_visits, sizeof(uint32_t) * size_t(d.policy_cap) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&sp->d_counters, sizeof(unsigned long long) * 16));
    BK_CUDA(cudaMalloc(&sp->d_stage, size_t(n_games) * 400 * 16));
    BK_CUDA(cudaMalloc(&sp->d_round, sizeof(int32_t) * 4));
    BK_CUDA(cudaMalloc(&sp->d_slot_base, sizeof(uint32_t) * (size_t(n_games) + 1)));
    BK_CUDA(cudaMemsetAsync(sp->d_slot_base, 0, sizeof(uint32_t) * (size_t(n_games) + 1), sp->env->stream));
    BK_CUDA(cudaMemsetAsync(sp->d_hdr, 0, sizeof(BkSearchHdr) * size_t(n_games), sp->env->stream));
    BK_CUDA(cudaMemsetAsync(sp->d_pol_off, 0, sizeof(uint32_t) * (BK_HIST_CAP + 1) * size_t(n_games), sp->env->stream));
    BK_CUDA(cudaMemsetAsync(sp->d_counters, 0, sizeof(unsigned long long) * 16, sp->env->stream));
    // simulation.rs:91-93 — the factor of ucb_score that depends only on the parent's visit count,
    // evaluated on the host with the platform libm exactly as the reference's f32 expression reads
    std::vector<float> ucb(size_t(cfg->sims_per_move) + 2);
    for (size_t i = 0; i < ucb.size(); ++i) {
        const float pv = float(i);
        ucb[i] = (std::log((pv + cfg->c_base + 1.0f) / cfg->c_base) + cfg->c_init) * std::sqrt(pv);
    }
    // simulation.rs:67-80 for the stub's constant policy 1.0: prior = e / (sequential f32 sum of n e's)
    std::vector<float> prior(401, 0.0f);
    const float e1 = host_exp_f32(1.0f);
    float total = 0.0f;
    for (int k = 1; k <= 400; ++k) { total += e1; prior[size_t(k)] = e1 / total; }
    // bk_ucb_div: exhaustive check of the short division over every (parent visits, child visits) pair this handle can see
    std::vector<float> rcp(ucb.size() + 1, 0.0f);
    for (size_t d = 1; d < rcp.size(); ++d) rcp[d] = 1.0f / float(d);
    bool short_div_exact = getenv("BK_NO_SHORT_DIV") == nullptr && cfg->sims_per_move <= 4096u;   // (sims + 2)^2 checks: bounded
    for (size_t i = 0; i < ucb.size() && short_div_exact; ++i)
        for (size_t d = 1; d < rcp.size(); ++d) {
            const float F = ucb[i], fd = float(d), q0 = F * rcp[d];
            const float q = std::fmaf(std::fmaf(-fd, q0, F), rcp[d], q0);
            const float want = F / fd;
            uint32_t a, b;
            memcpy(&a, &q, 4); memcpy(&b, &want, 4);
            if (a != b) { short_div_exact = false; break; }
        }
    sp->short_div = short_div_exact;
    BK_CUDA(cudaMalloc(&sp->d_rcp, sizeof(float) * rcp.size()));
    BK_CUDA(cudaMemcpy(sp->d_rcp, rcp.data(), sizeof(float) * rcp.size(), cudaMemcpyHostToDevice));
    d.rcp_tab = short_div_exact ? sp->d_rcp : nullptr;
    BK_CUDA(cudaMalloc(&sp->d_ucb, sizeof(float) * ucb.size()));
    BK_CUDA(cudaMalloc(&sp->d_prior, sizeof(float) * prior.size()));
    BK_CUDA(cudaMemcpy(sp->d_ucb, ucb.data(), sizeof(float) * ucb.size(), cudaMemcpyHostToDevice));
    BK_CUDA(cudaMemcpy(sp->d_prior, prior.data(), sizeof(float) * prior.size(), cudaMemcpyHostToDevice));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    d.ucb_tab = sp->d_ucb;
    d.prior_tab = sp->d_prior;
    BK_CUDA(cudaEventCreate(&sp->ev0));
    BK_CUDA(cudaEventCreate(&sp->ev1));
    return BK_OK;
}

extern "C" {

void bk_selfplay_destroy(bk_selfplay* sp) {
    if (!sp) return;
    cudaSetDevice(sp->device);
    cudaFree(sp->d_S); cudaFree(sp->d_X); cudaFree(sp->d_nodes);
    cudaFree(sp->d_scratch); cudaFree(sp->d_hdr); cudaFree(sp->d_pol_off); cudaFree(sp->d_pol_tile);
    cudaFree(sp->d_pol_visits); cudaFree(sp->d_ucb); cudaFree(sp->d_rcp); cudaFree(sp->d_prior); cudaFree(sp->d_counters);
    cudaFree(sp->d_stage);
    cudaFree(sp->d_round);
    cudaFree(sp->d_ply_off);
    cudaFree(sp->d_pend);
    cudaFree(sp->d_remap);
    cudaFree(sp->d_slot_base);
    cudaFree(sp->d_pack_off); cudaFree(sp->d_pack_ptr); cudaFree(sp->d_pack_tile); cudaFree(sp->d_pack_visits);
    if (sp->ev0) cudaEventDestroy(sp->ev0);
    if (sp->ev1) cudaEventDestroy(sp->ev1);
    bk_env_destroy(sp->env);
    delete sp;
}

int bk_selfplay_reset(bk_selfplay* sp, uint32_t first_game_id) {
    int rc = sp_use(sp);
    if (rc) return rc;
    rc = bk_env_reset(sp->env);
    if (rc) return rc;
    sp->first_id = first_game_id;
    sp->dcfg.first_game_id = first_game_id;
    BK_CUDA(cudaMemsetAsync(sp->d_hdr, 0, sizeof(BkSearchHdr) * size_t(sp->n), sp->env->stream));
    BK_CUDA(cudaMemsetAsync(sp->d_pol_off, 0, sizeof(uint32_t) * (BK_HIST_CAP + 1) * size_t(sp->n), sp->env->stream));
    return BK_OK;
}

int bk_selfplay_set_mode(bk_selfplay* sp, uint32_t flags, int leaves_per_round) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (flags & ~(BK_MODE_SKIP_FORCED | BK_MODE_FORCE_MULTI_LEAF | BK_MODE_TREE_REUSE))
        return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_set_mode: unknown flag");
    if (leaves_per_round < 1 || leaves_per_round > BK_MAX_LEAVES_PER_ROUND)
        return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_set_mode: leaves_per_round must be in 1..32");
    std::vector<BkSearchHdr> h(size_t(sp->n));
    BK_CUDA(cudaMemcpyAsync(h.data(), sp->d_hdr, sizeof(BkSearchHdr) * h.size(), cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    for (const BkSearchHdr& x : h)
        if (x.pend_kind != BK_PEND_NONE) return bk_fail(BK_ERR_STATE, "bk_selfplay_set_mode: a ply is in progress");
    const bool vl = leaves_per_round > 1 || (flags & BK_MODE_FORCE_MULTI_LEAF);
    if (vl && uint32_t(leaves_per_round) > sp->pend_cap) {       // grow-only; no other buffer depends on K
        cudaFree(sp->d_pend);
        sp->d_pend = nullptr;
        sp->pend_cap = 0u;
        BK_CUDA(cudaMalloc(&sp->d_pend, sizeof(BkPend) * size_t(sp->n) * size_t(leaves_per_round)));
        sp->pend_cap = uint32_t(leaves_per_round);
    }
    if ((flags & BK_MODE_TREE_REUSE) && !sp->d_remap)
        BK_CUDA(cudaMalloc(&sp->d_remap, sizeof(uint32_t) * 2 * size_t(sp->dcfg.max_nodes) * size_t(sp->n)));
    if (!(flags & BK_MODE_TREE_REUSE) && (sp->dcfg.mode & BK_MODE_TREE_REUSE_FLAG)) {   // leaving the mode: drop kept trees
        for (BkSearchHdr& x : h) x.reused = 0u;
        BK_CUDA(cudaMemcpyAsync(sp->d_hdr, h.data(), sizeof(BkSearchHdr) * h.size(), cudaMemcpyHostToDevice, sp->env->stream));
        BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    }
    sp->use_vl = vl;
    sp->dcfg.mode = flags;
    sp->dcfg.leaves_per_round = uint32_t(leaves_per_round);
    return BK_OK;
}

int bk_selfplay_run_stub(bk_selfplay* sp, int max_plies) {
    int rc = sp_use(sp);
    if (rc) return rc;
    cudaStream_t st = sp->env->stream;
    BK_CUDA(cudaEventRecord(sp->ev0, st));
    const int per_sm = (sp->n + sp->num_sms - 1) / sp->num_sms;     // games an SM would hold if all were resident
    // (a batch one game per SM slot too large for an instantiation runs in two waves: 1300 games took 78 ms for 12 plies on
    // <1> — 8 resident per SM — against 47 ms on <12>, profiles/r02_ab_mcts_pipe_threshold.log)
    const int minb = sp->stub_min_blocks ? sp->stub_min_blocks : (per_sm <= (sp->full_resident < 9 ? sp->full_resident : 9) ? 1 : (per_sm <= 12 ? 12 : (per_sm <= 16 ? 16 : (per_sm <= 20 ? 20 : 28))));
    // exact mode and a batch small enough to be latency bound (<= 9 games per SM): the two-warp pipeline
    // ... and only while every game is resident at once (one CTA per game: a second wave would double the run)
    const bool pipe = sp->dcfg.mode == 0u && (sp->stub_pipe > 0 || (sp->stub_pipe < 0 && per_sm <= 9 && per_sm <= sp->pipe_resident &&
                                                                    !sp->stub_min_blocks));
    if (pipe)
        BK_LAUNCH(k_selfplay_stub_pipe, sp->n, 64, st, sp->dcfg, pools_of(sp), sp->env->d_states, sp->env->d_hist, sp->n, max_plies,
                  sp->d_counters);
    else {
        // exact mode gets the kernel without the opt-in modes' code (see kb_selfplay_stub)
#define BK_STUB_LAUNCH(MINB, MODES) BK_LAUNCH((k_selfplay_stub<MINB, MODES>), sp->n, 32, st, sp->dcfg, pools_of(sp), sp->env->d_states, \
                                              sp->env->d_hist, sp->n, max_plies, sp->d_counters)
        const bool modes = sp->dcfg.mode != 0u;
        if (minb >= 28) { if (modes) BK_STUB_LAUNCH(28, true); else BK_STUB_LAUNCH(28, false); }
        else if (minb >= 20) { if (modes) BK_STUB_LAUNCH(20, true); else BK_STUB_LAUNCH(20, false); }
        else if (minb >= 16) { if (modes) BK_STUB_LAUNCH(16, true); else BK_STUB_LAUNCH(16, false); }
        else if (minb >= 12) { if (modes) BK_STUB_LAUNCH(12, true); else BK_STUB_LAUNCH(12, false); }
        else { if (modes) BK_STUB_LAUNCH(1, true); else BK_STUB_LAUNCH(1, false); }
#undef BK_STUB_LAUNCH
    }
    BK_CUDA(cudaEventRecord(sp->ev1, st));
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(st));
    BK_CUDA(cudaEventElapsedTime(&sp->last_ms, sp->ev0, sp->ev1));
    return sp_check_errors(sp);
}

static int sp_counts(bk_selfplay* sp, int32_t* pending, int32_t* done) {
    int32_t* d = sp->env->d_i32;
    cudaStream_t st = sp->env->stream;
    BK_CUDA(cudaMemsetAsync(d, 0, sizeof(int32_t) * 2, st));
    BK_LAUNCH(k_sp_count, (sp->n + 127) / 128, 128, st, pools_of(sp), sp->n, d);
    BK_CUDA(cudaGetLastError());
    int32_t h[2];
    BK_CUDA(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, st));
    BK_CUDA(cudaStreamSynchronize(st));
    if (pending) *pending = h[0];
    if (done) *done = h[1];
    return BK_OK;
}

int bk_selfplay_begin_ply(bk_selfplay* sp) {
    int rc = sp_use(sp);
    if (rc) return rc;
    BK_LAUNCH(k_sp_begin, (sp->n + 3) / 4, 128, sp->env->stream, sp->dcfg, pools_of(sp), sp->env->d_states, sp->n);
    BK_CUDA(cudaGetLastError());
    return BK_OK;
}

int bk_selfplay_leaf_planes(bk_selfplay* sp, float* dev_planes, int32_t* pending_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (!dev_planes) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_leaf_planes: dev_planes is null");
    cudaStream_t st = sp->env->stream;
    int32_t* d_cnt = sp->env->d_i32 + 2;
    BK_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int32_t), st));
    BK_LAUNCH(k_sp_slot_scan, 1, 256, st, pools_of(sp), sp->n, sp->use_vl ? 1 : 0, static_cast<int32_t*>(nullptr));
    BK_LAUNCH(k_sp_planes, sp->n * int(sp->dcfg.leaves_per_round), 256, st, sp->dcfg, pools_of(sp), sp->n,
              sp->use_vl ? 1 : 0, dev_planes, d_cnt);
    BK_CUDA(cudaGetLastError());
    if (pending_out) {
        BK_CUDA(cudaMemcpyAsync(pending_out, d_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        BK_CUDA(cudaStreamSynchronize(st));
    }
    return BK_OK;
}

int bk_selfplay_leaf_rows(bk_selfplay* sp, int32_t* rows_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (!rows_out) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_leaf_rows: null argument");
    uint32_t rows = 0;
    BK_CUDA(cudaMemcpyAsync(&rows, sp->d_slot_base + sp->n, sizeof(uint32_t), cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    *rows_out = int32_t(rows);
    return BK_OK;
}

int bk_selfplay_expand_backup(bk_selfplay* sp, const float* dev_policy, const float* dev_value, int32_t* pending_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (!dev_policy || !dev_value) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_expand_backup: null evaluator output");
    cudaStream_t st = sp->env->stream;
    BK_CUDA(cudaEventRecord(sp->ev0, st));
    if (sp->use_vl)
        BK_LAUNCH(k_sp_step_vl, sp->n, 32, st, sp->dcfg, pools_of(sp), sp->n, dev_policy, dev_value, sp->d_counters);
    else
        BK_LAUNCH(k_sp_step, sp->n, 32, st, sp->dcfg, pools_of(sp), sp->n, dev_policy, dev_value, sp->d_counters);
    BK_CUDA(cudaEventRecord(sp->ev1, st));
    BK_CUDA(cudaGetLastError());
    if (pending_out) {
        rc = sp_counts(sp, pending_out, nullptr);
        if (rc) return rc;
        BK_CUDA(cudaEventElapsedTime(&sp->last_ms, sp->ev0, sp->ev1));
    }
    return BK_OK;
}

int bk_selfplay_end_ply(bk_selfplay* sp) {
    int rc = sp_use(sp);
    if (rc) return rc;
    int32_t pending = 0;
    rc = sp_counts(sp, &pending, nullptr);
    if (rc) return rc;
    if (pending) return bk_fail(BK_ERR_STATE, "bk_selfplay_end_ply: some games still wait for the evaluator");
    BK_LAUNCH(k_sp_end, sp->n, 32, sp->env->stream, sp->dcfg, pools_of(sp), sp->env->d_states, sp->env->d_hist, sp->n,
              sp->d_counters);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    return sp_check_errors(sp);
}

int bk_selfplay_run_network(bk_selfplay* sp, bk_evaluator* ev, int max_plies, int64_t* rounds_out, int64_t* evals_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (!ev) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_run_network: null evaluator");
    if (ev->device != sp->device) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_run_network: evaluator lives on another device");
    if (int64_t(ev->cap_rows) < int64_t(sp->n) * int64_t(sp->dcfg.leaves_per_round))
        return bk_fail(BK_ERR_CAPACITY, "bk_selfplay_run_network: evaluator max_rows < n_games * leaves_per_round");
    cudaStream_t st = sp->env->stream;
    int32_t* d_round = sp->d_round;                  // {rows, games waiting}, written by the slot scan
    const int vl = sp->use_vl ? 1 : 0;
    const int slots = sp->n * int(sp->dcfg.leaves_per_round);
    int64_t rounds = 0, evals = 0;
    // enqueue: dense row assignment + the waiting positions' planes into the network's input; then one 8-byte read
    auto collect = [&](int32_t (&h)[2]) -> int {
        BK_LAUNCH(k_sp_slot_scan, 1, 256, st, pools_of(sp), sp->n, vl, d_round);
        BK_LAUNCH(k_sp_planes_nhwc, slots, 256, st, sp->dcfg, pools_of(sp), sp->n, vl, ev->d_x64);
        BK_CUDA(cudaGetLastError());
        BK_CUDA(cudaMemcpyAsync(h, d_round, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        BK_CUDA(cudaStreamSynchronize(st));
        return BK_OK;
    };
    BK_CUDA(cudaEventRecord(sp->ev0, st));
    for (int ply = 0; max_plies < 0 || ply < max_plies; ++ply) {
        int32_t live = 0;
        rc = bk_selfplay_live_games(sp, &live);
        if (rc) return rc;
        if (!live) break;
        rc = bk_selfplay_begin_ply(sp);
        if (rc) return rc;
        int32_t h[2] = {0, 0};
        rc = collect(h);
        if (rc) return rc;
        while (h[1] > 0) {
            if (h[0] > 0) {
                rc = bk_evaluator_forward_x64(ev, h[0], ev->d_policy, ev->d_value, nullptr, nullptr, st);
                if (rc) return rc;
                evals += h[0];
            }
            if (sp->use_vl)
                BK_LAUNCH(k_sp_step_vl, sp->n, 32, st, sp->dcfg, pools_of(sp), sp->n, ev->d_policy, ev->d_value, sp->d_counters);
            else
                BK_LAUNCH(k_sp_step, sp->n, 32, st, sp->dcfg, pools_of(sp), sp->n, ev->d_policy, ev->d_value, sp->d_counters);
            ++rounds;
            rc = collect(h);
            if (rc) return rc;
        }
        rc = bk_selfplay_end_ply(sp);
        if (rc) return rc;
    }
    BK_CUDA(cudaEventRecord(sp->ev1, st));
    BK_CUDA(cudaStreamSynchronize(st));
    BK_CUDA(cudaEventElapsedTime(&sp->last_ms, sp->ev0, sp->ev1));
    if (rounds_out) *rounds_out = rounds;
    if (evals_out) *evals_out = evals;
    return BK_OK;
}

int bk_selfplay_set_stream(bk_selfplay* sp, void* cuda_stream) {
    int rc = sp_use(sp);
    if (rc) return rc;
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    if (sp->env->stream && !sp->env->borrowed) cudaStreamDestroy(sp->env->stream);
    sp->env->stream = static_cast<cudaStream_t>(cuda_stream);
    sp->env->borrowed = true;
    return BK_OK;
}

int bk_selfplay_live_games(bk_selfplay* sp, int32_t* out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (!out) return bk_fail(BK_ERR_INVALID_ARG, "null argument");
    std::vector<int32_t> term(size_t(sp->n));
    rc = bk_env_is_terminal(sp->env, term.data());
    if (rc) return rc;
    int live = 0;
    for (int32_t t : term) live += t ? 0 : 1;
    *out = live;
    return BK_OK;
}

bk_env* bk_selfplay_env(bk_selfplay* sp) { return sp ? sp->env : nullptr; }

int bk_selfplay_results(bk_selfplay* sp, int32_t* plies_out, int32_t* policy_off_out, int32_t policy_cap,
                        int16_t* policy_tile_out, uint32_t* policy_visits_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    const size_t n = size_t(sp->n);
    std::vector<BkSearchHdr> h(n);
    BK_CUDA(cudaMemcpyAsync(h.data(), sp->d_hdr, sizeof(BkSearchHdr) * n, cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    std::vector<uint32_t> off(n * (BK_HIST_CAP + 1));
    BK_CUDA(cudaMemcpyAsync(off.data(), sp->d_pol_off, sizeof(uint32_t) * off.size(), cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    for (size_t g = 0; g < n; ++g) {
        const uint32_t plies = h[g].plies_searched;
        if (plies_out) plies_out[g] = int32_t(plies);
        if (policy_off_out)
            for (size_t k = 0; k <= BK_MAX_PLIES; ++k)
                policy_off_out[g * (BK_MAX_PLIES + 1) + k] = int32_t(k <= plies ? off[g * (BK_HIST_CAP + 1) + k] : h[g].pol_count);
        if ((policy_tile_out || policy_visits_out) && int64_t(h[g].pol_count) > int64_t(policy_cap))
            return bk_fail(BK_ERR_CAPACITY, "bk_selfplay_results: policy_cap too small");
        if (policy_tile_out && h[g].pol_count)
            BK_CUDA(cudaMemcpyAsync(policy_tile_out + g * size_t(policy_cap), sp->d_pol_tile + g * size_t(sp->dcfg.policy_cap),
                                    sizeof(uint16_t) * h[g].pol_count, cudaMemcpyDeviceToHost, sp->env->stream));
        if (policy_visits_out && h[g].pol_count)
            BK_CUDA(cudaMemcpyAsync(policy_visits_out + g * size_t(policy_cap), sp->d_pol_visits + g * size_t(sp->dcfg.policy_cap),
                                    sizeof(uint32_t) * h[g].pol_count, cudaMemcpyDeviceToHost, sp->env->stream));
    }
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    return BK_OK;
}

int bk_selfplay_results_sizes(bk_selfplay* sp, int64_t* total_plies_out, int64_t* total_entries_out, int64_t* ply_offset_out,
                              int64_t* entry_offset_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    const size_t n = size_t(sp->n);
    cudaStream_t st = sp->env->stream;
    std::vector<BkSearchHdr> h(n);
    BK_CUDA(cudaMemcpyAsync(h.data(), sp->d_hdr, sizeof(BkSearchHdr) * n, cudaMemcpyDeviceToHost, st));
    BK_CUDA(cudaStreamSynchronize(st));
    std::vector<int64_t> off(2 * (n + 1), 0);
    for (size_t g = 0; g < n; ++g) {
        off[g + 1] = off[g] + int64_t(h[g].plies_searched);
        off[n + 1 + g + 1] = off[n + 1 + g] + int64_t(h[g].pol_count);
    }
    if (!sp->d_pack_off) BK_CUDA(cudaMalloc(&sp->d_pack_off, sizeof(int64_t) * 2 * (n + 1)));
    BK_CUDA(cudaMemcpyAsync(sp->d_pack_off, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, st));
    BK_CUDA(cudaStreamSynchronize(st));
    sp->pack_plies = off[n];
    sp->pack_entries = off[2 * n + 1];
    if (total_plies_out) *total_plies_out = sp->pack_plies;
    if (total_entries_out) *total_entries_out = sp->pack_entries;
    if (ply_offset_out) for (size_t g = 0; g <= n; ++g) ply_offset_out[g] = off[g];
    if (entry_offset_out) for (size_t g = 0; g <= n; ++g) entry_offset_out[g] = off[n + 1 + g];
    return BK_OK;
}

int bk_selfplay_results_packed(bk_selfplay* sp, int64_t* ply_ptr_out, int16_t* tile_out, uint32_t* visits_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (sp->pack_plies < 0) return bk_fail(BK_ERR_STATE, "bk_selfplay_results_packed: call bk_selfplay_results_sizes first");
    if (!ply_ptr_out || !tile_out || !visits_out) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_results_packed: null output");
    cudaStream_t st = sp->env->stream;
    if (sp->pack_plies + 1 > sp->pack_cap_plies) {
        cudaFree(sp->d_pack_ptr); sp->d_pack_ptr = nullptr; sp->pack_cap_plies = 0;
        BK_CUDA(cudaMalloc(&sp->d_pack_ptr, sizeof(int64_t) * size_t(sp->pack_plies + 1)));
        sp->pack_cap_plies = sp->pack_plies + 1;
    }
    if (sp->pack_entries + 1 > sp->pack_cap_entries) {
        cudaFree(sp->d_pack_tile); cudaFree(sp->d_pack_visits); sp->d_pack_tile = nullptr; sp->d_pack_visits = nullptr;
        sp->pack_cap_entries = 0;
        BK_CUDA(cudaMalloc(&sp->d_pack_tile, sizeof(uint16_t) * size_t(sp->pack_entries + 1)));
        BK_CUDA(cudaMalloc(&sp->d_pack_visits, sizeof(uint32_t) * size_t(sp->pack_entries + 1)));
        sp->pack_cap_entries = sp->pack_entries + 1;
    }
    BK_LAUNCH(k_pack_results, sp->n, 256, st, sp->dcfg, pools_of(sp), sp->n, sp->d_pack_off, sp->d_pack_off + (sp->n + 1),
              sp->d_pack_ptr, sp->d_pack_tile, sp->d_pack_visits);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemcpyAsync(ply_ptr_out, sp->d_pack_ptr, sizeof(int64_t) * size_t(sp->pack_plies + 1), cudaMemcpyDeviceToHost, st));
    if (sp->pack_entries > 0) {
        BK_CUDA(cudaMemcpyAsync(tile_out, sp->d_pack_tile, sizeof(uint16_t) * size_t(sp->pack_entries), cudaMemcpyDeviceToHost, st));
        BK_CUDA(cudaMemcpyAsync(visits_out, sp->d_pack_visits, sizeof(uint32_t) * size_t(sp->pack_entries), cudaMemcpyDeviceToHost, st));
    }
    BK_CUDA(cudaStreamSynchronize(st));
    sp->pack_plies = -1;     // sizes are valid for one gather: the games may move on
    return BK_OK;
}

int bk_selfplay_last_root(bk_selfplay* sp, int32_t* counts_out, int16_t* tile_out, uint32_t* visits_out,
                          float* value_sum_out, float* prior_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    const size_t n = size_t(sp->n);
    int32_t* d_counts = sp->env->d_i32;
    uint8_t* base = sp->d_stage;
    uint32_t* d_vis = reinterpret_cast<uint32_t*>(base);
    float* d_w = reinterpret_cast<float*>(base + n * 400 * 4);
    float* d_p = reinterpret_cast<float*>(base + n * 400 * 8);
    int16_t* d_tile = reinterpret_cast<int16_t*>(base + n * 400 * 12);
    cudaStream_t st = sp->env->stream;
    BK_LAUNCH(k_last_root, sp->n, 128, st, sp->dcfg, pools_of(sp), sp->n, d_counts, d_tile, d_vis, d_w, d_p);
    BK_CUDA(cudaGetLastError());
    if (counts_out) BK_CUDA(cudaMemcpyAsync(counts_out, d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (tile_out) BK_CUDA(cudaMemcpyAsync(tile_out, d_tile, sizeof(int16_t) * 400 * n, cudaMemcpyDeviceToHost, st));
    if (visits_out) BK_CUDA(cudaMemcpyAsync(visits_out, d_vis, sizeof(uint32_t) * 400 * n, cudaMemcpyDeviceToHost, st));
    if (value_sum_out) BK_CUDA(cudaMemcpyAsync(value_sum_out, d_w, sizeof(float) * 400 * n, cudaMemcpyDeviceToHost, st));
    if (prior_out) BK_CUDA(cudaMemcpyAsync(prior_out, d_p, sizeof(float) * 400 * n, cudaMemcpyDeviceToHost, st));
    BK_CUDA(cudaStreamSynchronize(st));
    return BK_OK;
}

int bk_selfplay_training_sizes(bk_selfplay* sp, int64_t* total_plies_out, int64_t* ply_offset_out) {
    int rc = sp_use(sp);
    if (rc) return rc;
    const size_t n = size_t(sp->n);
    std::vector<BkSearchHdr> h(n);
    BK_CUDA(cudaMemcpyAsync(h.data(), sp->d_hdr, sizeof(BkSearchHdr) * n, cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    // The tensors pair history entry k with policy record k and replay the boards from an empty one, which is right only
    // if every ply of a game was searched by this handle (no bk_env_apply / playout before or between the searches).
    std::vector<BkSummary> sm(n);
    BK_LAUNCH(k_sp_summary, (sp->n + 3) / 4, 128, sp->env->stream, sp->env->d_states, sp->env->d_summary, sp->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemcpyAsync(sm.data(), sp->env->d_summary, sizeof(BkSummary) * n, cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    for (size_t g = 0; g < n; ++g)
        if (uint32_t(sm[g].ply) != h[g].plies_searched)
            return bk_fail(BK_ERR_STATE, "bk_selfplay_training_sizes: game " + std::to_string(g) + " has " + std::to_string(sm[g].ply) +
                                             " plies of history but " + std::to_string(h[g].plies_searched) +
                                             " searched plies: the games were advanced outside the search, the records do not align");
    std::vector<int64_t> off(n + 1, 0);
    for (size_t g = 0; g < n; ++g) off[g + 1] = off[g] + int64_t(h[g].plies_searched);
    if (!sp->d_ply_off) BK_CUDA(cudaMalloc(&sp->d_ply_off, sizeof(int64_t) * (n + 1)));
    BK_CUDA(cudaMemcpyAsync(sp->d_ply_off, off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    if (total_plies_out) *total_plies_out = off[n];
    if (ply_offset_out) for (size_t g = 0; g <= n; ++g) ply_offset_out[g] = off[g];
    return BK_OK;
}

int bk_selfplay_training_tensors(bk_selfplay* sp, float* dev_states, float* dev_policies, float* dev_values) {
    int rc = sp_use(sp);
    if (rc) return rc;
    if (!dev_states || !dev_policies || !dev_values) return bk_fail(BK_ERR_INVALID_ARG, "bk_selfplay_training_tensors: null output");
    if (!sp->d_ply_off) return bk_fail(BK_ERR_STATE, "bk_selfplay_training_tensors: call bk_selfplay_training_sizes first");
    cudaStream_t st = sp->env->stream;
    // payoffs come from the games' final states
    BK_LAUNCH(k_sp_summary, (sp->n + 3) / 4, 128, st, sp->env->d_states, sp->env->d_summary, sp->n);
    BK_CUDA(cudaEventRecord(sp->ev0, st));
    BK_LAUNCH(k_training_tensors, sp->n, 256, st, sp->dcfg, pools_of(sp), sp->env->d_hist, sp->env->d_summary, sp->d_ply_off,
              sp->n, dev_states, dev_policies, dev_values);
    BK_CUDA(cudaEventRecord(sp->ev1, st));
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(st));
    BK_CUDA(cudaEventElapsedTime(&sp->last_ms, sp->ev0, sp->ev1));
    return BK_OK;
}

int bk_selfplay_counters(bk_selfplay* sp, uint64_t out[6]) {
    int rc = sp_use(sp);
    if (rc) return rc;
    unsigned long long h[6];
    BK_CUDA(cudaMemcpyAsync(h, sp->d_counters, sizeof h, cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    for (int i = 0; i < 6; ++i) out[i] = h[i];
    return BK_OK;
}

int bk_selfplay_counters_raw(bk_selfplay* sp, uint64_t out[16]) {
    int rc = sp_use(sp);
    if (rc) return rc;
    unsigned long long h[16];
    BK_CUDA(cudaMemcpyAsync(h, sp->d_counters, sizeof h, cudaMemcpyDeviceToHost, sp->env->stream));
    BK_CUDA(cudaStreamSynchronize(sp->env->stream));
    for (int i = 0; i < 16; ++i) out[i] = h[i];
#if defined(BK_PIPE_STATS) && !defined(BK_WARP_EMU)
    unsigned long long g[32];
    BK_CUDA(cudaMemcpyFromSymbol(g, g_pipe_stats, sizeof g));
    for (int i = 0; i < 7; ++i) out[9 + i] = g[i];
#endif
    return BK_OK;
}

int bk_selfplay_probe_stats(bk_selfplay* sp, uint64_t out[32]) {
    int rc = sp_use(sp);
    if (rc) return rc;
    for (int i = 0; i < 32; ++i) out[i] = 0;
#if defined(BK_PIPE_STATS) && !defined(BK_WARP_EMU)
    unsigned long long g[32];
    BK_CUDA(cudaMemcpyFromSymbol(g, g_pipe_stats, sizeof g));
    for (int i = 0; i < 32; ++i) out[i] = g[i];
#endif
    return BK_OK;
}

int bk_selfplay_last_kernel_ms(bk_selfplay* sp, float* ms_out) {
    if (!sp || !ms_out) return bk_fail(BK_ERR_INVALID_ARG, "null argument");
    *ms_out = sp->last_ms;
    return BK_OK;
}

}  // extern "C"
