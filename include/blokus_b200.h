/* blokus_b200.h — C ABI of the B200-native Blokus self-play hot path.
 *
 * Drop-in boundary for the path SURVEY.md §8 names: the `blokus` crate's game-state API
 * (blokus/src/game.rs:91-312, blokus/src/board.rs:18-206) and the `self_play` crate's MCTS /
 * client interface (self_play/src/lib.rs:9-63, self_play/src/simulation.rs:14-296), batched over
 * thousands of games that live in B200 HBM.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - Every call returns 0 on success or a negative bk_status; bk_last_error() returns a
 *     thread-local UTF-8 message (the analogue of the reference's Err(String)).
 *   - A handle is bound to one CUDA device and one stream and is not thread-safe; distinct handles
 *     are independent (mirrors the reference's one-process-per-game model, model/training.py:204).
 *   - Unless a parameter is named dev_*, buffers are HOST memory allocated by the caller; the
 *     library performs the host<->device copies on the handle's stream and synchronises it.
 *   - Arrays are per game, game-major: out[g][...] for g in [0, n_games).
 *   - Tiles are row*20+col (0..399); players/seats are 0..3; piece ids are the PIECE_TYPES order of
 *     blokus/src/pieces.rs:30-52.
 *   - There is no CPU fallback: without a CUDA device every compute call fails with BK_ERR_CUDA.
 */
#ifndef BLOKUS_B200_H
#define BLOKUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_BOARD_DIM 20
#define BK_BOARD_TILES 400
#define BK_NUM_PLAYERS 4
#define BK_MAX_PLIES 360 /* >= 4 players * 89 squares */

typedef enum bk_status {
    BK_OK = 0,
    BK_ERR_INVALID_ARG = -1,
    BK_ERR_CUDA = -2,
    BK_ERR_ILLEGAL_MOVE = -3, /* some game rejected its tile; see the per-game status output */
    BK_ERR_CAPACITY = -4,     /* a per-game tree / output pool overflowed */
    BK_ERR_STATE = -5         /* call not valid in the handle's current phase */
} bk_status;

typedef struct bk_env bk_env;           /* a batch of Game values (game.rs:91-99) in HBM */
typedef struct bk_selfplay bk_selfplay; /* a batch of self-play clients (simulation.rs:267-296) */

/* ---- library ---------------------------------------------------------------------------------- */
const char* bk_last_error(void);
const char* bk_version(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int bk_device_count(void);

/* ---- static piece tables (blokus/src/pieces.rs:58-210) ------------------------------------------ */
/* Piece::points (pieces.rs:153) */
int bk_piece_points(int piece_id);
/* Piece::variants.len() (pieces.rs:185-209) */
int bk_piece_num_variants(int piece_id);
/* PieceVariant{offsets,width} and variant.len() (pieces.rs:66-98). offsets_out holds <= 5 ints.
 * Returns the number of offsets, or a negative status. */
int bk_piece_variant(int piece_id, int variant, int* width_out, int* len_out, int* offsets_out);

/* ---- game batch: Game (blokus/src/game.rs) -------------------------------------------------------- */
/* Game::reset() for n_games games on `device` (game.rs:102-114). */
int bk_env_create(int n_games, int device, bk_env** out);
void bk_env_destroy(bk_env* env);
int bk_env_num_games(const bk_env* env);
/* Game::reset() in place for every game. */
int bk_env_reset(bk_env* env);
/* Game::clone() (game.rs:91 #[derive(Clone)]): a new batch with identical state on the same device. */
int bk_env_clone(const bk_env* env, bk_env** out);

/* Game::apply(tile, piece_to_finish) (game.rs:150-194) for every game at once.
 *   tiles[g] < 0            : game g is skipped this call.
 *   piece_to_finish         : NULL, or per game the index into the mover's REMAINING piece list
 *                             (board.rs:151-153) to commit after this tile; < 0 means None.
 *   status_out              : NULL, or per game 0 ok / BK_ERR_ILLEGAL_MOVE / 1 skipped.
 * Deviation from the reference: an illegal tile is rejected WITHOUT mutating that
 * game (the reference mutates the board and then returns Err, game.rs:152-164; no caller relies on
 * that state).  Returns BK_ERR_ILLEGAL_MOVE if any game rejected its tile. */
int bk_env_apply(bk_env* env, const int32_t* tiles, const int32_t* piece_to_finish, int32_t* status_out);

/* Game::place_piece(p, v, o) (game.rs:116-144), in place (the reference returns a new Game; use
 * bk_env_clone first for value semantics).  p = index into the mover's remaining list, v = variant
 * index within the piece, o = stride-20 offset of the variant's bounding box.  p[g] < 0 skips game g.
 * Only valid at the start of a turn (which is how gui/src/app.rs:156 uses it). */
int bk_env_place_piece(bk_env* env, const int32_t* p, const int32_t* v, const int32_t* o, int32_t* status_out);

/* Game::get_legal_tiles() (game.rs:242-244) as a 0/1 mask per tile, out[g][400]. Mid-piece this is
 * the narrowed set, as in the reference. */
int bk_env_legal_mask(bk_env* env, uint8_t* out);
/* Same set as 20 row words per game (bit c of out[g][r] = tile r*20+c). */
int bk_env_legal_rows(bk_env* env, uint32_t* out);
/* The reference's own shape, `Vec<usize>` (game.rs:242-244): counts_out[g] tiles in tiles_out[g][0 .. counts_out[g]),
 * ASCENDING (the reference returns HashMap keys in arbitrary order); entries past the count are unspecified. */
int bk_env_legal_tiles(bk_env* env, int32_t* counts_out, int16_t* tiles_out);
/* Game::get_board() (game.rs:196-198): the reference's byte encoding, out[g][400]: occupied cell =
 * 0xF0 | owner(1..4); empty cell = OR of 1<<(4+p) over players p with an orthogonal neighbour
 * (board.rs:95-119). */
int bk_env_board(bk_env* env, uint8_t* out);
/* Board::get_anchors(player) (board.rs:143-145) as a 0/1 mask, out[g][400]; player < 0 = the
 * current player (Game::get_current_anchors, game.rs:238-240). */
int bk_env_anchors(bk_env* env, int player, uint8_t* out);
/* Game::current_player() (game.rs:225-228), out[g]. */
int bk_env_current_player(bk_env* env, int32_t* out);
/* Game::is_terminal() (game.rs:275-277), out[g] 0/1. */
int bk_env_is_terminal(bk_env* env, int32_t* out);
/* Game::is_player_active(p) (game.rs:279-281), out[g][4] 0/1. */
int bk_env_is_player_active(bk_env* env, int32_t* out);
/* Game::get_score() (game.rs:247-249; board.rs:155-181), out[g][4]. */
int bk_env_scores(bk_env* env, int32_t* out);
/* Game::get_payoff() (game.rs:252-272), out[g][4]. */
int bk_env_payoff(bk_env* env, float* out);
/* Game::get_board_state() (game.rs:283-311): 5 planes 20x20, mover-relative and rotated to the
 * mover's frame, out[g][5][20][20] as 0/1 bytes. */
int bk_env_board_state(bk_env* env, uint8_t* out);
/* Same planes written as float32 into DEVICE memory (the evaluator's input batch, replacing the
 * nested-list pickle of simulation.rs:50-52 / model/training.py:29-40). */
int bk_env_board_state_dev_f32(bk_env* env, float* dev_out);
/* Game::history (game.rs:94): counts_out[g] plies, players_out/tiles_out[g][BK_MAX_PLIES]. */
int bk_env_history(bk_env* env, int32_t* counts_out, int32_t* players_out, int32_t* tiles_out);
/* Remaining pieces per player as piece-id bit masks, out[g][4] (Board::get_pieces, board.rs:147;
 * the remaining LIST is the set bits in ascending id order). */
int bk_env_pieces(bk_env* env, uint32_t* out);
/* last_piece_lens (game.rs:98), out[g][4]. */
int bk_env_last_piece_lens(bk_env* env, int32_t* out);
/* Order-independent 64-bit digest of each game's full state (tests/trace hashes), out[g]. */
int bk_env_digest(bk_env* env, uint64_t* out);

/* ---- lockstep random playouts (BASELINE.json configs 1-2) --------------------------------------- */
#define BK_PLAYOUT_HASH 1u     /* also fold a per-ply state digest into hash_out (parity runs) */
#define BK_PLAYOUT_MIN_TILE 2u /* policy: always the smallest legal tile (seed-free trace)     */
#define BK_PLAYOUT_MAX_TILE 4u /* policy: always the largest legal tile  (seed-free trace)     */
#define BK_PLAYOUT_NEW_GAME 8u /* Game::reset (game.rs:102-114) first, inside the same launch: equals bk_env_reset + bk_env_playout */
/* Plays every game of the batch forward from its current state until it is terminal or max_plies
 * more tiles were applied (max_plies < 0: to the end), entirely on the device: legal-tile
 * generation, seeded choice, Game::apply, per ply.  Default policy: ascending legal tiles, index
 * floor(u*n) with u from Philox4x32-10 keyed (seed, first_game_id+g, ply).
 * The env keeps the final states; query them with the bk_env_* calls.  The kernels are ENQUEUED on the
 * handle's stream and the call returns; any later call that hands data to the host synchronises. */
int bk_env_playout(bk_env* env, uint64_t seed, uint32_t first_game_id, int max_plies, uint32_t flags);
/* As bk_env_playout, but game g's global id is game_ids[g] (HOST array of n_games entries, copied to
 * the device inside the call) — the form a sharded driver uses. */
int bk_env_playout_ids(bk_env* env, uint64_t seed, const uint32_t* game_ids, int max_plies, uint32_t flags);
/* One device->host gather of what a batch produced, straight into the caller's buffers on the handle's
 * stream (pin them for full PCIe rate): plies_out[g] (history length), scores_out[g][4], and the history
 * packed as uint16 (tile | player << 9), history_packed_out[g][BK_MAX_PLIES].  Pointers may be NULL. */
int bk_env_fetch(bk_env* env, int32_t* plies_out, int32_t* scores_out, uint16_t* history_packed_out);
/* The same gather enqueued on the handle's stream WITHOUT waiting (the host buffers should be pinned); the
 * results are valid after bk_env_sync.  With two handles (each has its own stream) a caller overlaps the
 * device->host copy of one batch with the playout of the next — the double-buffered end-to-end loop of bench.py. */
int bk_env_fetch_async(bk_env* env, int32_t* plies_out, int32_t* scores_out, uint16_t* history_packed_out);
int bk_env_sync(bk_env* env);
/* Results of the last bk_env_playout: per game steps applied by it, and the chained trace hash
 * (0 unless BK_PLAYOUT_HASH).  Either pointer may be NULL. */
int bk_env_playout_results(bk_env* env, int32_t* steps_out, uint64_t* hash_out);
/* Device time of the kernels launched by the last bk_env_playout / bk_env_apply, CUDA events on the
 * handle's stream. */
int bk_env_last_kernel_ms(bk_env* env, float* ms_out);
/* Region timing on the handle's stream: record event `which` (0 = begin, 1 = end) now; elapsed is
 * end - begin in ms after the stream has drained.  Lets a caller time several calls as one region with
 * CUDA events on the stream the kernels are launched on. */
int bk_env_event_record(bk_env* env, int which);
int bk_env_event_elapsed(bk_env* env, float* ms_out);
/* Work counters of the last bk_env_playout summed over the batch: [0] steps, [1] turn-start move
 * generations, [2] sum of 120*C_rem over those generations (SURVEY.md §8d algorithmic lane-ops). */
int bk_env_playout_counters(bk_env* env, uint64_t out[3]);

/* Measured integer-pipe peak of `device` in 32-bit lane-ops per second (LOP3/SHF mix, every SM busy):
 * the denominator of the integer roofline the rules engine is bound by. */
int bk_probe_int_peak(int device, double* lane_ops_per_s_out, float* ms_out);

/* ---- MCTS self-play clients (self_play/src/simulation.rs) ------------------------------------------ */
/* simulation.rs:14-22 plus the seed the reference lacks. */
typedef struct bk_config {
    uint32_t sims_per_move;
    uint32_t sample_moves;
    float c_base;
    float c_init;
    float dirichlet_alpha;
    float exploration_fraction;
    uint64_t seed;
} bk_config;

/* n_games clients with global ids first_game_id.. (RNG keyed by global id, so results do not depend
 * on how games are sharded over GPUs).  max_children_per_game bounds the per-ply child pool
 * (0 = worst case (sims_per_move+1)*126). */
int bk_selfplay_create(int n_games, int device, const bk_config* cfg, uint32_t first_game_id,
                       uint32_t max_children_per_game, bk_selfplay** out);
void bk_selfplay_destroy(bk_selfplay* sp);
/* Start every client over from Game::reset() with new global ids first_game_id.. (pools are reused). */
int bk_selfplay_reset(bk_selfplay* sp, uint32_t first_game_id);
/* training_game() with the fixed-prior stub evaluator (policy 1.0 on legal tiles, value 0.25 per
 * seat; BASELINE.json config 3) for up to max_plies more plies per game (< 0: to the end), fully on
 * the device: per ply mcts() = root evaluate, Dirichlet noise, sims_per_move x (select, apply,
 * evaluate/expand, backpropagate), policy record, select_action, Game::apply. */
int bk_selfplay_run_stub(bk_selfplay* sp, int max_plies);
/* External-evaluator protocol, one leaf per live game per round (replaces queue.put / pipe.recv of
 * simulation.rs:50-57 by one contiguous device batch).  Per ply:
 *   bk_selfplay_begin_ply     : start mcts() for every live game; its root position becomes the pending leaf.
 *   loop until no game is pending:
 *     bk_selfplay_leaf_planes   : dev_planes[R][5][20][20] float32 = get_board_state() of every pending
 *                                 position, DENSE: the R positions waiting for an answer are packed in game order
 *                                 into the first R rows of a buffer of capacity n_games (x leaves_per_round) rows —
 *                                 finished games and games whose ply is done cost the evaluator nothing.
 *                                 R = bk_selfplay_leaf_rows.  pending_out (host, may be NULL) = number of games
 *                                 that need the next expand_backup call.
 *     (caller runs its evaluator on the first R rows of the device batch)
 *     bk_selfplay_expand_backup : consume dev_policy[R][400] (mover frame) and dev_value[R][4]
 *                                 (relative seat), same row order: expand the pending position (children for legal tiles with
 *                                 policy > 0, priors exp(p)/sum), back the value up, then run further
 *                                 simulations until each game's next NON-terminal leaf is pending (terminal
 *                                 leaves are backed up on the device) or its sims_per_move are done.
 *                                 pending_out (host, may be NULL) = games that now wait for the evaluator.
 *   bk_selfplay_end_ply       : record the policy, select_action, Game::apply (BK_ERR_STATE if some game
 *                               still has a pending position). */
int bk_selfplay_begin_ply(bk_selfplay* sp);
int bk_selfplay_leaf_planes(bk_selfplay* sp, float* dev_planes, int32_t* pending_out);
int bk_selfplay_expand_backup(bk_selfplay* sp, const float* dev_policy, const float* dev_value, int32_t* pending_out);
/* Rows R of the evaluator batch the last bk_selfplay_leaf_planes wrote = positions waiting for an answer (at most
 * n_games in the exact mode, n_games x leaves_per_round in the multi-leaf mode). */
int bk_selfplay_leaf_rows(bk_selfplay* sp, int32_t* rows_out);
int bk_selfplay_end_ply(bk_selfplay* sp);
/* Opt-in throughput modes.  The default (flags 0, leaves_per_round 1) is the
 * reference's exact behaviour — mcts() runs sims_per_move simulations one at a time from a fresh tree
 * (simulation.rs:183,192-210) — and is the only mode whose visit counts are comparable with the reference.
 *   BK_MODE_SKIP_FORCED      : a root position with exactly ONE legal tile is not searched; its policy record
 *                              [(tile, sims_per_move visits)] = [(tile, 1.0)] and its action are what the search
 *                              would return anyway (simulation.rs:213-229), so the training tuple is unchanged.
 *   leaves_per_round K > 1   : (external-evaluator protocol) up to K simulations per game are in flight per
 *                              round, separated by virtual loss.  The evaluator batch is then DENSE: the leaves of
 *                              all games are packed in (game, slot) order into the first R rows of planes
 *                              [<= n*K][5][20][20] (R from bk_selfplay_leaf_rows after bk_selfplay_leaf_planes), and
 *                              expand_backup reads policy [R][400] / value [R][4] in the same order — the evaluator
 *                              never works on empty slots.
 *                              With K == 1 results equal the exact mode bit for bit (BK_MODE_FORCE_MULTI_LEAF
 *                              runs that code path with K == 1, for tests).
 *   BK_MODE_TREE_REUSE       : after a ply's action is played, the subtree below the chosen root child is
 *                              compacted in place and becomes the next ply's tree; that search adds fresh root
 *                              noise and only tops the root up to sims_per_move visits (the reference starts a
 *                              new tree every ply, simulation.rs:183).  Every recorded policy still sums to
 *                              sims_per_move visits.
 * Must be called between plies (BK_ERR_STATE otherwise).  bk_selfplay_run_stub honours BK_MODE_SKIP_FORCED and
 * BK_MODE_TREE_REUSE; leaves_per_round applies to the external-evaluator protocol. */
#define BK_MODE_SKIP_FORCED 1u
#define BK_MODE_FORCE_MULTI_LEAF 2u
#define BK_MODE_TREE_REUSE 4u
int bk_selfplay_set_mode(bk_selfplay* sp, uint32_t flags, int leaves_per_round);
/* Run every kernel of this handle (and of its bk_env) on the caller's CUDA stream (a cudaStream_t passed as
 * void*; NULL = the legacy default stream) so an evaluator enqueued on that stream needs no extra sync. */
int bk_selfplay_set_stream(bk_selfplay* sp, void* cuda_stream);
/* Number of games not yet terminal. */
int bk_selfplay_live_games(bk_selfplay* sp, int32_t* out);
/* The batch of games being played (borrowed; valid until bk_selfplay_destroy). */
bk_env* bk_selfplay_env(bk_selfplay* sp);
/* training_game()'s return value (simulation.rs:293-295) for every game:
 *   plies_out[g]; history via bk_env_history on bk_selfplay_env();
 *   policy records: policy_off_out[g][BK_MAX_PLIES+1] offsets into that game's slice of
 *   policy_tile_out / policy_visits_out (each [g][policy_cap]); prob = visits / sum(visits) in f32
 *   (simulation.rs:213-225).  payoff via bk_env_payoff.  Returns BK_ERR_CAPACITY if policy_cap is
 *   too small. */
int bk_selfplay_results(bk_selfplay* sp, int32_t* plies_out, int32_t* policy_off_out, int32_t policy_cap,
                        int16_t* policy_tile_out, uint32_t* policy_visits_out);
/* The same records as ONE compressed-row gather for large batches (one kernel, three copies instead of two copies per
 * game): bk_selfplay_results_sizes returns the totals and the per-game prefixes (n_games + 1 entries each; any pointer
 * may be NULL) and must be called first; bk_selfplay_results_packed then fills HOST buffers
 *   ply_ptr_out[total_plies + 1]   entry range of ply k of game g: [ply_ptr[ply_offset[g] + k], ply_ptr[ply_offset[g] + k + 1])
 *   tile_out[total_entries], visits_out[total_entries]
 * (pinned buffers give the full PCIe rate). */
int bk_selfplay_results_sizes(bk_selfplay* sp, int64_t* total_plies_out, int64_t* total_entries_out, int64_t* ply_offset_out,
                              int64_t* entry_offset_out);
int bk_selfplay_results_packed(bk_selfplay* sp, int64_t* ply_ptr_out, int16_t* tile_out, uint32_t* visits_out);
/* Root children of the LAST searched ply of each game (debug / Q parity): counts_out[g], then
 * [g][400] tile, visits, value_sum, prior. */
int bk_selfplay_last_root(bk_selfplay* sp, int32_t* counts_out, int16_t* tile_out, uint32_t* visits_out,
                          float* value_sum_out, float* prior_out);
/* The consumer side — model/training.py:70-119 `save()` — on the device: for every
 * searched ply of every game, in game-major order, the training triple
 *   states[ply][5][20][20]  planes 0..3 = squares laid before the ply by seats mover, mover+1, ..; plane 4 = the
 *                           recorded policy's tiles; rotated `mover` quarter turns (torch.rot90(k=mover))
 *   policies[ply][400]      visits / total visits (f32), rotated the same way
 *   values[ply][4]          the game's payoff (absolute seat order), repeated
 * bk_selfplay_training_sizes returns the total number of plies and the per-game prefix offsets (n_games + 1
 * entries; either pointer may be NULL) and must be called first; bk_selfplay_training_tensors fills DEVICE
 * buffers of total*2000, total*400 and total*4 floats. */
int bk_selfplay_training_sizes(bk_selfplay* sp, int64_t* total_plies_out, int64_t* ply_offset_out);
int bk_selfplay_training_tensors(bk_selfplay* sp, float* dev_states, float* dev_policies, float* dev_values);
/* Counters since creation: [0] simulations, [1] Game::apply calls, [2] turn-start move generations,
 * [3] sum of 120*C_rem, [4] child entries created, [5] nodes expanded. */
int bk_selfplay_counters(bk_selfplay* sp, uint64_t out[6]);
/* All 16 counter slots; [6..] are written only by instrumented probe builds (-DBK_PIPE_STATS: cycles the two warps of the
 * pipelined stub kernel spend waiting for each other, and total cycles). */
int bk_selfplay_counters_raw(bk_selfplay* sp, uint64_t out[16]);
/* Finer probe counters of instrumented builds (zeros otherwise); layout documented in csrc/bk_mcts_pipe.cuh. */
int bk_selfplay_probe_stats(bk_selfplay* sp, uint64_t out[32]);
int bk_selfplay_last_kernel_ms(bk_selfplay* sp, float* ms_out);

/* ---- leaf evaluator building block ------------------------------------------------ */
/* One 3x3 convolution of the reference's ResNet trunk (model/resnet.py:13-14: Conv2d(C, C, 3, padding=1) with
 * C = 256) as a hand-written tcgen05/TMEM/TMA kernel, BatchNorm folded by the caller into w/bias:
 *     y = act(conv3x3(x) + bias [+ residual])
 * All pointers are DEVICE memory.  x, residual, y: bf16, zero-padded NHWC [batch*441][256] with row index
 * image*441 + r*21 + c (r, c in 0..20; r == 20 and c == 20 are zero padding and stay zero in y).
 * w: bf16 [9 taps][256 out][256 in], tap = (dy+1)*3 + (dx+1).  bias: f32 [256].  residual may be NULL.
 * relu != 0 applies ReLU.  The kernel is enqueued on cuda_stream (a cudaStream_t passed as void*). */
int bk_conv3x3_bf16(const void* dev_x, const void* dev_w, const float* dev_bias, const void* dev_residual,
                    void* dev_y, int batch, int relu, void* cuda_stream);
/* Same kernel for a narrower input (model/resnet.py:51, Conv2d(5, 256, 3) with the 5 planes zero-extended to 64
 * channels): x bf16 [batch*441][in_channels], w bf16 [9][256][in_channels], in_channels in {64, 128, 192, 256};
 * y bf16 [batch*441][256]. */
int bk_conv3x3_bf16_in(const void* dev_x, const void* dev_w, const float* dev_bias, void* dev_y, int batch,
                       int in_channels, int relu, void* cuda_stream);

/* ---- leaf evaluator as one native object (BASELINE.json config 4) ------------------------ */
/* The reference's policy/value network, `ResNet(blocks, 256)` of model/resnet.py:44-94, in eval mode, entirely on this
 * library's kernels: input packing -> 2*blocks+1 tcgen05 convolutions -> one fused head kernel (both 1x1 head
 * convolutions + BatchNorm + ReLU, the policy's masked softmax x mask, the value's Linear(400,4) + tanh + softmax).
 * It replaces what the reference's inference server computes per batch (model/training.py:43-67).
 * Parameters are HOST pointers (copied once), BatchNorm folded by the caller (scale s = gamma / sqrt(var + eps)):
 *   w_in      bf16 [9][256][64]   model.input.weight, tap = ky*3 + kx, the 5 input planes zero-extended to 64
 *   b_in      f32  [256]          model.input.bias                (no BN / ReLU after it, resnet.py:79)
 *   w_blocks  bf16 [2*blocks][9][256][256]   conv1, conv2 of each block times their BN scale
 *   b_blocks  f32  [2*blocks][256]           (conv bias - running_mean) * s + beta
 *   head_w    f32  [2][256]       policy_head[0].weight, value_head[0].weight (1x1 convolutions)
 *   head_affine f32 [4]           policy {s, conv_bias*s + beta - mean*s}, value {same}: head = relu(dot * a0 + a1)
 *   lin_w     f32  [4][400], lin_b f32 [4]   value_head[4] (Linear)
 * max_rows = positions per forward pass the activation buffers hold (3 x max_rows x 441 x 512 B). */
typedef struct bk_evaluator bk_evaluator;
int bk_evaluator_create(int device, int blocks, int max_rows, const void* w_in, const float* b_in, const void* w_blocks,
                        const float* b_blocks, const float* head_w, const float* head_affine, const float* lin_w,
                        const float* lin_b, bk_evaluator** out);
void bk_evaluator_destroy(bk_evaluator* ev);
int bk_evaluator_max_rows(const bk_evaluator* ev);
/* `model(boards)` (resnet.py:69-94): dev_planes f32 [rows][5][20][20] -> dev_policy f32 [rows][400] (mover frame, zero
 * on illegal tiles), dev_value f32 [rows][4] (relative seats).  dev_logits [rows][400] / dev_vtanh [rows][4] (may be
 * NULL) receive the pre-softmax head outputs (policy_head(x), value_head(x)) for parity checks.  All DEVICE memory;
 * enqueued on cuda_stream, no synchronisation. */
int bk_evaluator_forward(bk_evaluator* ev, const float* dev_planes, int rows, float* dev_policy, float* dev_value,
                         float* dev_logits, float* dev_vtanh, void* cuda_stream);
/* training_game() (simulation.rs:267-296) for every client with this network as the evaluator, up to max_plies plies
 * (< 0: to the end).  The whole round stays on the device: the pending positions' planes are written straight into
 * the first convolution's input (no float planes), the network runs, expand + backup consume its outputs; per round
 * the host reads back 8 bytes (rows to evaluate, games waiting).  Honours bk_selfplay_set_mode (dense rows, multi-leaf
 * rounds, forced-ply shortcut, tree reuse).  rounds_out / evals_out (may be NULL): evaluator rounds, positions evaluated. */
int bk_selfplay_run_network(bk_selfplay* sp, bk_evaluator* ev, int max_plies, int64_t* rounds_out, int64_t* evals_out);
/* The evaluator's two non-convolution stages as stand-alone operators on DEVICE memory (building blocks, like
 * bk_conv3x3_bf16): float planes [rows][5][20][20] -> the first convolution's input x64 bf16 [rows*441][64] (which the
 * caller zero-initialised once: only channels 0..7 of the 400 real cells are written), and the fused heads on the
 * trunk's output act bf16 [rows*441][256]; dev_head_params f32 [2120] = head_w[512], head_affine[4], lin_w[1600], lin_b[4]. */
int bk_eval_pack_planes(const float* dev_planes, int rows, void* dev_x64, void* cuda_stream);
/* Game::get_board_state (game.rs:283-311) of every game of the batch written directly in that input layout
 * (dev_x64 bf16 [n_games*441][64], zero-initialised by the caller); synchronises the batch's stream. */
int bk_env_board_state_nhwc(bk_env* env, void* dev_x64);
int bk_eval_heads(const void* dev_act, const void* dev_x64, const float* dev_head_params, int rows, float* dev_policy,
                  float* dev_value, float* dev_logits, float* dev_vtanh, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* BLOKUS_B200_H */
