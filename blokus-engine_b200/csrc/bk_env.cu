// bk_env.cu — batch-of-games kernels and the bk_env_* C ABI (include/blokus_b200.h).
// Host mirror of blokus/src/game.rs `Game`: every accessor of game.rs:196-311 has an entry point.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "bk_host.h"

// ---- error plumbing -----------------------------------------------------------------------------
static thread_local std::string g_last_error;
void bk_set_error(const std::string& msg) { g_last_error = msg; }
int bk_fail(int code, const std::string& msg) { g_last_error = msg; return code; }

// ---- kernels (thin wrappers; bodies in bk_env_kernels.cuh) ------------------------------------------
#define BK_STEP_WARPS 4  // warps (games) per CTA for the per-call kernels

__global__ void __launch_bounds__(32 * BK_STEP_WARPS) k_reset(BkState* __restrict__ states, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_reset(states, g, lane);
}

__global__ void __launch_bounds__(32 * BK_STEP_WARPS)
k_apply(BkState* __restrict__ states, uint16_t* __restrict__ hist, const int32_t* __restrict__ tiles,
        const int32_t* __restrict__ finish, int32_t* __restrict__ status, int n, unsigned long long* counters) {
    __shared__ uint32_t smem[BK_TABS_SMEM_WORDS];
    const BkTabs tabs = bk_stage_tables(smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_apply(states, hist, tiles, finish, status, counters, g, lane, tabs);
}

__global__ void __launch_bounds__(32 * BK_STEP_WARPS)
k_place_piece(BkState* __restrict__ states, uint16_t* __restrict__ hist, const int32_t* __restrict__ pp,
              const int32_t* __restrict__ vv, const int32_t* __restrict__ oo, int32_t* __restrict__ status, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_place_piece(states, hist, pp, vv, oo, status, g, lane);
}

// One warp per CTA: the hardware CTA scheduler balances games of different length over the SMs.
// 32 CTAs (games) per SM is the hardware's CTA limit: 64 registers.  (A 28-per-SM build — 68 registers, enough for
// config 2's 27.7 games per SM — measured the same on one batch and 2 % slower with two batches in flight.)
template <bool HASH>
__global__ void __launch_bounds__(32, 32)
k_playout(BkState* __restrict__ states, uint16_t* __restrict__ hist, int n, uint64_t seed, uint32_t first_id,
          const uint32_t* __restrict__ ids, int max_plies, uint32_t flags, int32_t* __restrict__ steps_out, uint64_t* __restrict__ hash_out,
          unsigned long long* counters) {
    // the candidate window masks are copied to shared memory here: this kernel's first-tile scan gathers them on its critical
    // path, and reading them in place through L1 (bk_global_tables) measured 5.5 % slower (profiles/r02_ab_playout_global_cands.log)
    __shared__ uint32_t smem[BK_TABS_SMEM_WORDS];
    const BkTabs tabs = bk_stage_tables(smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x;                     // launched with one warp per CTA: the game index is CTA-uniform
    if (g >= n) return;
    const uint32_t game_id = __reduce_or_sync(BK_FULL, ids ? ids[g] : first_id + uint32_t(g));   // uniform for the compiler too
    kb_playout<HASH>(states, hist, seed, game_id, max_plies, flags, steps_out, hash_out, counters, g, lane, tabs);
}

__global__ void k_scores(const BkState* __restrict__ states, int32_t* __restrict__ plies, int32_t* __restrict__ scores, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_scores(states, plies, scores, g, lane);
}

// Integer-pipe peak probe: 8 independent LOP3/SHF chains per thread, exact op count by inline PTX.
#define BK_PROBE_ITERS 4096
__global__ void __launch_bounds__(256) k_int_probe(uint32_t* __restrict__ out) {
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const uint32_t k = blockIdx.x | 0x55aa00ffu;
#ifndef BK_WARP_EMU
#pragma unroll 1
    for (int i = 0; i < BK_PROBE_ITERS; ++i) {
#define BK_PROBE_STEP(x) asm volatile("lop3.b32 %0, %0, %1, %0, 0x96; shr.u32 %0, %0, 1;" : "+r"(x) : "r"(k));
        BK_PROBE_STEP(a0) BK_PROBE_STEP(a1) BK_PROBE_STEP(a2) BK_PROBE_STEP(a3)
        BK_PROBE_STEP(a4) BK_PROBE_STEP(a5) BK_PROBE_STEP(a6) BK_PROBE_STEP(a7)
        BK_PROBE_STEP(a0) BK_PROBE_STEP(a1) BK_PROBE_STEP(a2) BK_PROBE_STEP(a3)
        BK_PROBE_STEP(a4) BK_PROBE_STEP(a5) BK_PROBE_STEP(a6) BK_PROBE_STEP(a7)
    }
#endif
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ k;
}

__global__ void k_summary(const BkState* __restrict__ states, BkSummary* __restrict__ out, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    kb_summary(states, out, g, lane);
}

__global__ void k_cells(const BkState* __restrict__ states, uint8_t* __restrict__ out, int n, int what, int player) {
    const int g = blockIdx.x;
    if (g >= n) return;
    kb_cells(states, out, what, player, g, threadIdx.x, blockDim.x);
}

template <typename T>
__global__ void k_planes(const BkState* __restrict__ states, T* __restrict__ out, int n) {
    const int g = blockIdx.x;
    if (g >= n) return;
    kb_planes<T>(&states[g], out + size_t(g) * 2000, threadIdx.x, blockDim.x);
}

__global__ void k_legal_rows(const BkState* __restrict__ states, uint32_t* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * 20) out[i] = states[i / 20].legal[i % 20];
}

// Game::get_legal_tiles (game.rs:242-244) in list form: counts[g], then the tiles ascending in tiles[g][0 .. counts[g]).
// One warp per game: lane r owns row r, a shuffle scan of the row popcounts gives every row its first slot.
__global__ void k_legal_list(const BkState* __restrict__ states, int32_t* __restrict__ counts, int16_t* __restrict__ tiles, int n) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    uint32_t row = lane < 20 ? states[g].legal[lane] : 0u;
    const int cnt = __popc(row);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(BK_FULL, incl, d);
        if (lane >= d) incl += v;
    }
    int pos = incl - cnt;
    int16_t* out = tiles + size_t(g) * 400;
    for (; row; row &= row - 1u) out[pos++] = int16_t(lane * 20 + __ffs(row) - 1);
    if (lane == 31) counts[g] = incl;
}

// ---- host side --------------------------------------------------------------------------------------
static int grid_for(int n, int warps) { return (n + warps - 1) / warps; }

static int env_alloc_into(bk_env* e, int n_games, cudaStream_t stream) {
#ifndef BK_WARP_EMU
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device) == cudaSuccess && sms > 0) e->num_sms = sms;
    }
#endif
    if (stream) { e->stream = stream; e->borrowed = true; }
    else BK_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    BK_CUDA(cudaMalloc(&e->d_states, sizeof(BkState) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&e->d_hist, sizeof(uint16_t) * BK_HIST_CAP * size_t(n_games)));
    BK_CUDA(cudaMalloc(&e->d_i32, sizeof(int32_t) * 4 * size_t(n_games)));
    BK_CUDA(cudaMalloc(&e->d_hash, sizeof(uint64_t) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&e->d_counters, sizeof(unsigned long long) * 8));
    BK_CUDA(cudaMalloc(&e->d_summary, sizeof(BkSummary) * size_t(n_games)));
    BK_CUDA(cudaMalloc(&e->d_bytes, size_t(2000) * size_t(n_games)));
    BK_CUDA(cudaMemsetAsync(e->d_hist, 0, sizeof(uint16_t) * BK_HIST_CAP * size_t(n_games), e->stream));
    BK_CUDA(cudaMemsetAsync(e->d_i32, 0, sizeof(int32_t) * 4 * size_t(n_games), e->stream));
    BK_CUDA(cudaMemsetAsync(e->d_hash, 0, sizeof(uint64_t) * size_t(n_games), e->stream));
    BK_CUDA(cudaMemsetAsync(e->d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    BK_CUDA(cudaEventCreate(&e->ev0));
    BK_CUDA(cudaEventCreate(&e->ev1));
    BK_CUDA(cudaEventCreate(&e->uev[0]));
    BK_CUDA(cudaEventCreate(&e->uev[1]));
    return BK_OK;
}

int bk_env_alloc(int n_games, int device, cudaStream_t stream, bk_env** out) {
    if (n_games <= 0 || !out) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_create: n_games must be > 0");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return bk_fail(BK_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
    if (device < 0 || device >= count) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_create: bad device index");
    BK_CUDA(cudaSetDevice(device));
    bk_env* e = new bk_env();
    e->n = n_games;
    e->device = device;
    const int rc = env_alloc_into(e, n_games, stream);
    if (rc) { bk_env_destroy(e); return rc; }       // nothing leaks when an allocation fails half way
    *out = e;
    return BK_OK;
}

static int env_use(const bk_env* e) {
    if (!e) return bk_fail(BK_ERR_INVALID_ARG, "null bk_env handle");
    BK_CUDA(cudaSetDevice(e->device));
    return BK_OK;
}

// Close a timed kernel region.  sync = false leaves the work enqueued (the stream orders it before any
// later call of this handle; results are read by calls that synchronise); the elapsed time is then
// resolved lazily by bk_env_last_kernel_ms.
static int env_finish_timed(bk_env* e, bool sync) {
    BK_CUDA(cudaEventRecord(e->ev1, e->stream));
    BK_CUDA(cudaGetLastError());
    e->timing_pending = true;
    if (sync) {
        BK_CUDA(cudaStreamSynchronize(e->stream));
        BK_CUDA(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
        e->timing_pending = false;
    }
    return BK_OK;
}

static int env_summary(bk_env* e, std::vector<BkSummary>& host) {
    int rc = env_use(e);
    if (rc) return rc;
    BK_LAUNCH(k_summary, grid_for(e->n, BK_STEP_WARPS), 32 * BK_STEP_WARPS, e->stream, e->d_states, e->d_summary, e->n);
    BK_CUDA(cudaGetLastError());
    host.resize(size_t(e->n));
    BK_CUDA(cudaMemcpyAsync(host.data(), e->d_summary, sizeof(BkSummary) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

static int env_cells(bk_env* e, int what, int player, uint8_t* out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!out) return bk_fail(BK_ERR_INVALID_ARG, "null output buffer");
    BK_LAUNCH(k_cells, e->n, 128, e->stream, e->d_states, e->d_bytes, e->n, what, player);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemcpyAsync(out, e->d_bytes, size_t(400) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

static int env_fetch(bk_env* e, int32_t* plies_out, int32_t* scores_out, uint16_t* history_packed_out, bool wait);

extern "C" {

const char* bk_last_error(void) { return g_last_error.c_str(); }
#ifdef BK_WARP_EMU
const char* bk_version(void) { return "blokus-engine_b200 0.2 (CPU warp emulator of the kernel sources: tests only)"; }
#else
const char* bk_version(void) { return "blokus-engine_b200 0.2 (sm_100a)"; }
#endif
int bk_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

// ---- static tables (host copies of the generated header) ----
static const uint8_t h_piece_points[BK_NUM_PIECES] = BK_PIECE_POINTS_INIT;
static const uint8_t h_piece_first_variant[BK_NUM_PIECES + 1] = BK_PIECE_FIRST_VARIANT_INIT;
static const uint8_t h_variant_width[BK_NUM_VARIANTS] = BK_VARIANT_WIDTH_INIT;
static const uint8_t h_variant_height[BK_NUM_VARIANTS] = BK_VARIANT_HEIGHT_INIT;
static const uint8_t h_variant_ncells[BK_NUM_VARIANTS] = BK_VARIANT_NCELLS_INIT;
static const uint16_t h_variant_offsets[BK_NUM_VARIANTS][5] = BK_VARIANT_OFFSETS_INIT;

int bk_piece_points(int piece_id) {
    if (piece_id < 0 || piece_id >= BK_NUM_PIECES) return bk_fail(BK_ERR_INVALID_ARG, "bad piece id");
    return h_piece_points[piece_id];
}
int bk_piece_num_variants(int piece_id) {
    if (piece_id < 0 || piece_id >= BK_NUM_PIECES) return bk_fail(BK_ERR_INVALID_ARG, "bad piece id");
    return h_piece_first_variant[piece_id + 1] - h_piece_first_variant[piece_id];
}
int bk_piece_variant(int piece_id, int variant, int* width_out, int* len_out, int* offsets_out) {
    if (piece_id < 0 || piece_id >= BK_NUM_PIECES) return bk_fail(BK_ERR_INVALID_ARG, "bad piece id");
    const int gv = h_piece_first_variant[piece_id] + variant;
    if (variant < 0 || gv >= h_piece_first_variant[piece_id + 1]) return bk_fail(BK_ERR_INVALID_ARG, "bad variant index");
    if (width_out) *width_out = h_variant_width[gv];
    if (len_out) *len_out = (h_variant_height[gv] - 1) * 20 + h_variant_width[gv];
    for (int j = 0; j < h_variant_ncells[gv]; ++j)
        if (offsets_out) offsets_out[j] = h_variant_offsets[gv][j];
    return h_variant_ncells[gv];
}

int bk_env_create(int n_games, int device, bk_env** out) {
    int rc = bk_env_alloc(n_games, device, nullptr, out);
    if (rc) return rc;
    return bk_env_reset(*out);
}

void bk_env_destroy(bk_env* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_states); cudaFree(e->d_hist); cudaFree(e->d_i32); cudaFree(e->d_hash);
    cudaFree(e->d_counters); cudaFree(e->d_summary); cudaFree(e->d_bytes);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->uev[0]) cudaEventDestroy(e->uev[0]);
    if (e->uev[1]) cudaEventDestroy(e->uev[1]);
    if (e->stream && !e->borrowed) cudaStreamDestroy(e->stream);
    delete e;
}

int bk_env_num_games(const bk_env* e) { return e ? e->n : 0; }

int bk_env_reset(bk_env* e) {
    int rc = env_use(e);
    if (rc) return rc;
    BK_LAUNCH(k_reset, grid_for(e->n, BK_STEP_WARPS), 32 * BK_STEP_WARPS, e->stream, e->d_states, e->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemsetAsync(e->d_hist, 0, sizeof(uint16_t) * BK_HIST_CAP * size_t(e->n), e->stream));
    return BK_OK;   // enqueued; every reader synchronises the stream
}

int bk_env_clone(const bk_env* e, bk_env** out) {
    int rc = env_use(e);
    if (rc) return rc;
    rc = bk_env_alloc(e->n, e->device, nullptr, out);
    if (rc) return rc;
    bk_env* c = *out;
    BK_CUDA(cudaStreamSynchronize(e->stream));
    BK_CUDA(cudaMemcpyAsync(c->d_states, e->d_states, sizeof(BkState) * size_t(e->n), cudaMemcpyDeviceToDevice, c->stream));
    BK_CUDA(cudaMemcpyAsync(c->d_hist, e->d_hist, sizeof(uint16_t) * BK_HIST_CAP * size_t(e->n), cudaMemcpyDeviceToDevice, c->stream));
    BK_CUDA(cudaStreamSynchronize(c->stream));
    return BK_OK;
}

int bk_env_apply(bk_env* e, const int32_t* tiles, const int32_t* piece_to_finish, int32_t* status_out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!tiles) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_apply: tiles is null");
    const size_t nb = sizeof(int32_t) * size_t(e->n);
    int32_t* d_tiles = e->d_i32;
    int32_t* d_fin = e->d_i32 + e->n;
    int32_t* d_status = e->d_i32 + 2 * size_t(e->n);
    BK_CUDA(cudaMemcpyAsync(d_tiles, tiles, nb, cudaMemcpyHostToDevice, e->stream));
    if (piece_to_finish) BK_CUDA(cudaMemcpyAsync(d_fin, piece_to_finish, nb, cudaMemcpyHostToDevice, e->stream));
    BK_CUDA(cudaEventRecord(e->ev0, e->stream));
    BK_LAUNCH(k_apply, grid_for(e->n, BK_STEP_WARPS), 32 * BK_STEP_WARPS, e->stream, 
        e->d_states, e->d_hist, d_tiles, piece_to_finish ? d_fin : nullptr, d_status, e->n, nullptr);
    rc = env_finish_timed(e, true);
    if (rc) return rc;
    std::vector<int32_t> st(size_t(e->n));
    BK_CUDA(cudaMemcpyAsync(st.data(), d_status, nb, cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    int bad = -1;
    for (int g = 0; g < e->n; ++g) {
        if (status_out) status_out[g] = st[size_t(g)];
        if (st[size_t(g)] < 0 && bad < 0) bad = g;
    }
    if (bad >= 0) {
        char buf[96];
        snprintf(buf, sizeof buf, "Invalid move - game %d, Tile %d", bad, tiles[bad]);  // cf. game.rs:159-162
        return bk_fail(BK_ERR_ILLEGAL_MOVE, buf);
    }
    return BK_OK;
}

int bk_env_place_piece(bk_env* e, const int32_t* p, const int32_t* v, const int32_t* o, int32_t* status_out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!p || !v || !o) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_place_piece: null argument");
    const size_t nb = sizeof(int32_t) * size_t(e->n);
    int32_t* d_p = e->d_i32;
    int32_t* d_v = e->d_i32 + e->n;
    int32_t* d_o = e->d_i32 + 2 * size_t(e->n);
    int32_t* d_status = e->d_i32 + 3 * size_t(e->n);
    BK_CUDA(cudaMemcpyAsync(d_p, p, nb, cudaMemcpyHostToDevice, e->stream));
    BK_CUDA(cudaMemcpyAsync(d_v, v, nb, cudaMemcpyHostToDevice, e->stream));
    BK_CUDA(cudaMemcpyAsync(d_o, o, nb, cudaMemcpyHostToDevice, e->stream));
    BK_LAUNCH(k_place_piece, grid_for(e->n, BK_STEP_WARPS), 32 * BK_STEP_WARPS, e->stream, e->d_states, e->d_hist, d_p, d_v,
                                                                                        d_o, d_status, e->n);
    BK_CUDA(cudaGetLastError());
    std::vector<int32_t> st(size_t(e->n));
    BK_CUDA(cudaMemcpyAsync(st.data(), d_status, nb, cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    bool bad = false;
    for (int g = 0; g < e->n; ++g) {
        if (status_out) status_out[g] = st[size_t(g)];
        bad |= st[size_t(g)] < 0;
    }
    return bad ? bk_fail(BK_ERR_ILLEGAL_MOVE, "Invalid move") : BK_OK;  // game.rs:123
}

int bk_env_legal_mask(bk_env* e, uint8_t* out) { return env_cells(e, 0, 0, out); }
int bk_env_board(bk_env* e, uint8_t* out) { return env_cells(e, 1, 0, out); }
int bk_env_anchors(bk_env* e, int player, uint8_t* out) {
    if (player > 3) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_anchors: player must be < 4");
    return env_cells(e, 2, player, out);
}

int bk_env_legal_rows(bk_env* e, uint32_t* out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!out) return bk_fail(BK_ERR_INVALID_ARG, "null output buffer");
    uint32_t* d = reinterpret_cast<uint32_t*>(e->d_bytes);
    BK_LAUNCH(k_legal_rows, (e->n * 20 + 255) / 256, 256, e->stream, e->d_states, d, e->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemcpyAsync(out, d, sizeof(uint32_t) * 20 * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

int bk_env_legal_tiles(bk_env* e, int32_t* counts_out, int16_t* tiles_out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!counts_out || !tiles_out) return bk_fail(BK_ERR_INVALID_ARG, "null output buffer");
    int32_t* d_cnt = e->d_i32;
    int16_t* d_tiles = reinterpret_cast<int16_t*>(e->d_bytes);          // [n][400] int16 fits the [n][2000] byte staging
    BK_LAUNCH(k_legal_list, grid_for(e->n, 4), 128, e->stream, e->d_states, d_cnt, d_tiles, e->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemcpyAsync(counts_out, d_cnt, sizeof(int32_t) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaMemcpyAsync(tiles_out, d_tiles, sizeof(int16_t) * 400 * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

#define BK_SUMMARY_GETTER(name, type, width, expr)                      \
    int name(bk_env* e, type* out) {                                    \
        if (!out) return bk_fail(BK_ERR_INVALID_ARG, "null output buffer"); \
        std::vector<BkSummary> s;                                       \
        int rc = env_summary(e, s);                                     \
        if (rc) return rc;                                              \
        for (int g = 0; g < e->n; ++g)                                  \
            for (int k = 0; k < width; ++k) out[size_t(g) * width + k] = expr; \
        return BK_OK;                                                   \
    }
BK_SUMMARY_GETTER(bk_env_current_player, int32_t, 1, s[size_t(g)].cur)
BK_SUMMARY_GETTER(bk_env_is_terminal, int32_t, 1, s[size_t(g)].terminal)
BK_SUMMARY_GETTER(bk_env_is_player_active, int32_t, 4, s[size_t(g)].active[k])
BK_SUMMARY_GETTER(bk_env_scores, int32_t, 4, s[size_t(g)].scores[k])
BK_SUMMARY_GETTER(bk_env_payoff, float, 4, s[size_t(g)].payoff[k])
BK_SUMMARY_GETTER(bk_env_pieces, uint32_t, 4, s[size_t(g)].pieces[k])
BK_SUMMARY_GETTER(bk_env_last_piece_lens, int32_t, 4, s[size_t(g)].lastlens[k])
BK_SUMMARY_GETTER(bk_env_digest, uint64_t, 1, s[size_t(g)].digest)

int bk_env_board_state(bk_env* e, uint8_t* out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!out) return bk_fail(BK_ERR_INVALID_ARG, "null output buffer");
    BK_LAUNCH(k_planes<uint8_t>, e->n, 256, e->stream, e->d_states, e->d_bytes, e->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaMemcpyAsync(out, e->d_bytes, size_t(2000) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

int bk_env_board_state_dev_f32(bk_env* e, float* dev_out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!dev_out) return bk_fail(BK_ERR_INVALID_ARG, "null output buffer");
    BK_LAUNCH(k_planes<float>, e->n, 256, e->stream, e->d_states, dev_out, e->n);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

int bk_env_history(bk_env* e, int32_t* counts_out, int32_t* players_out, int32_t* tiles_out) {
    std::vector<BkSummary> s;
    int rc = env_summary(e, s);
    if (rc) return rc;
    std::vector<uint16_t> h(size_t(e->n) * BK_HIST_CAP);
    BK_CUDA(cudaMemcpyAsync(h.data(), e->d_hist, sizeof(uint16_t) * h.size(), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    for (int g = 0; g < e->n; ++g) {
        const int cnt = s[size_t(g)].ply;
        if (counts_out) counts_out[g] = cnt;
        for (int i = 0; i < BK_MAX_PLIES; ++i) {
            const uint16_t w = i < cnt ? h[size_t(g) * BK_HIST_CAP + i] : 0;
            if (players_out) players_out[size_t(g) * BK_MAX_PLIES + i] = i < cnt ? (w >> 9) : -1;
            if (tiles_out) tiles_out[size_t(g) * BK_MAX_PLIES + i] = i < cnt ? (w & 0x1FF) : -1;
        }
    }
    return BK_OK;
}

static int env_playout(bk_env* e, uint64_t seed, uint32_t first_game_id, const uint32_t* ids_host, int max_plies,
                       uint32_t flags) {
    int rc = env_use(e);
    if (rc) return rc;
    uint32_t* d_ids = nullptr;
    if (ids_host) {
        d_ids = reinterpret_cast<uint32_t*>(e->d_i32 + size_t(e->n));
        BK_CUDA(cudaMemcpyAsync(d_ids, ids_host, sizeof(uint32_t) * size_t(e->n), cudaMemcpyHostToDevice, e->stream));
    }
    BK_CUDA(cudaMemsetAsync(e->d_counters, 0, sizeof(unsigned long long) * 8, e->stream));
    BK_CUDA(cudaEventRecord(e->ev0, e->stream));
    // the per-ply state digest and the tests' seed-free policies are their own instantiation: the plain playout carries
    // neither their registers nor their code
    if (flags & (BK_PLAYOUT_HASH_FLAG | BK_PLAYOUT_MIN_TILE_FLAG | BK_PLAYOUT_MAX_TILE_FLAG))
        BK_LAUNCH(k_playout<true>, e->n, 32, e->stream, e->d_states, e->d_hist, e->n, seed, first_game_id, d_ids, max_plies,
                  flags, e->d_i32, e->d_hash, e->d_counters);
    else
        BK_LAUNCH(k_playout<false>, e->n, 32, e->stream, e->d_states, e->d_hist, e->n, seed, first_game_id, d_ids, max_plies,
                  flags, e->d_i32, e->d_hash, e->d_counters);
    return env_finish_timed(e, false);
}

int bk_env_playout(bk_env* e, uint64_t seed, uint32_t first_game_id, int max_plies, uint32_t flags) {
    return env_playout(e, seed, first_game_id, nullptr, max_plies, flags);
}

int bk_env_playout_ids(bk_env* e, uint64_t seed, const uint32_t* game_ids, int max_plies, uint32_t flags) {
    if (!game_ids) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_playout_ids: game_ids is null");
    return env_playout(e, seed, 0u, game_ids, max_plies, flags);
}

int bk_env_fetch(bk_env* e, int32_t* plies_out, int32_t* scores_out, uint16_t* history_packed_out) {
    return env_fetch(e, plies_out, scores_out, history_packed_out, true);
}

int bk_env_fetch_async(bk_env* e, int32_t* plies_out, int32_t* scores_out, uint16_t* history_packed_out) {
    return env_fetch(e, plies_out, scores_out, history_packed_out, false);
}

int bk_env_sync(bk_env* e) {
    int rc = env_use(e);
    if (rc) return rc;
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

}  // extern "C"

static int env_fetch(bk_env* e, int32_t* plies_out, int32_t* scores_out, uint16_t* history_packed_out, bool wait) {
    int rc = env_use(e);
    if (rc) return rc;
    int32_t* d_plies = e->d_i32 + 2 * size_t(e->n);
    int32_t* d_scores = reinterpret_cast<int32_t*>(e->d_bytes);
    if (plies_out || scores_out) {
        BK_LAUNCH(k_scores, grid_for(e->n, BK_STEP_WARPS), 32 * BK_STEP_WARPS, e->stream, e->d_states, d_plies, d_scores, e->n);
        BK_CUDA(cudaGetLastError());
    }
    if (plies_out) BK_CUDA(cudaMemcpyAsync(plies_out, d_plies, sizeof(int32_t) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    if (scores_out) BK_CUDA(cudaMemcpyAsync(scores_out, d_scores, sizeof(int32_t) * 4 * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    if (history_packed_out)
        BK_CUDA(cudaMemcpyAsync(history_packed_out, e->d_hist, sizeof(uint16_t) * BK_HIST_CAP * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    if (wait) BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

extern "C" {

int bk_probe_int_peak(int device, double* lane_ops_per_s_out, float* ms_out) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return bk_fail(BK_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
    if (device < 0 || device >= count || !lane_ops_per_s_out) return bk_fail(BK_ERR_INVALID_ARG, "bk_probe_int_peak: bad argument");
    BK_CUDA(cudaSetDevice(device));
    const int grid = 148 * 8, block = 256;
    uint32_t* d = nullptr;
    BK_CUDA(cudaMalloc(&d, sizeof(uint32_t) * size_t(grid) * block));
    cudaEvent_t a, b;
    BK_CUDA(cudaEventCreate(&a));
    BK_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        BK_CUDA(cudaEventRecord(a, nullptr));
        BK_LAUNCH(k_int_probe, grid, block, nullptr, d);
        BK_CUDA(cudaEventRecord(b, nullptr));
        BK_CUDA(cudaDeviceSynchronize());
        float ms = 0.f;
        BK_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    const double ops = double(grid) * block * double(BK_PROBE_ITERS) * 16.0 * 2.0;  // 16 steps x (lop3 + shr)
    *lane_ops_per_s_out = ops / (double(best) * 1e-3);
    if (ms_out) *ms_out = best;
    return BK_OK;
}

int bk_env_playout_results(bk_env* e, int32_t* steps_out, uint64_t* hash_out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (steps_out) BK_CUDA(cudaMemcpyAsync(steps_out, e->d_i32, sizeof(int32_t) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    if (hash_out) BK_CUDA(cudaMemcpyAsync(hash_out, e->d_hash, sizeof(uint64_t) * size_t(e->n), cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    return BK_OK;
}

int bk_env_last_kernel_ms(bk_env* e, float* ms_out) {
    if (!e || !ms_out) return bk_fail(BK_ERR_INVALID_ARG, "null argument");
    if (e->timing_pending) {
        BK_CUDA(cudaSetDevice(e->device));
        BK_CUDA(cudaEventSynchronize(e->ev1));
        BK_CUDA(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
        e->timing_pending = false;
    }
    *ms_out = e->last_ms;
    return BK_OK;
}

int bk_env_event_record(bk_env* e, int which) {
    int rc = env_use(e);
    if (rc) return rc;
    if (which < 0 || which > 1) return bk_fail(BK_ERR_INVALID_ARG, "bk_env_event_record: which must be 0 or 1");
    BK_CUDA(cudaEventRecord(e->uev[which], e->stream));
    return BK_OK;
}

int bk_env_event_elapsed(bk_env* e, float* ms_out) {
    int rc = env_use(e);
    if (rc) return rc;
    if (!ms_out) return bk_fail(BK_ERR_INVALID_ARG, "null argument");
    BK_CUDA(cudaStreamSynchronize(e->stream));
    BK_CUDA(cudaEventElapsedTime(ms_out, e->uev[0], e->uev[1]));
    return BK_OK;
}

int bk_env_playout_counters(bk_env* e, uint64_t out[3]) {
    int rc = env_use(e);
    if (rc) return rc;
    unsigned long long h[3];
    BK_CUDA(cudaMemcpyAsync(h, e->d_counters, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    BK_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < 3; ++i) out[i] = h[i];
    return BK_OK;
}

}  // extern "C"
