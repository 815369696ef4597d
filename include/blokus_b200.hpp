// blokus_b200.hpp — header-only C++ mirror of the reference's `blokus::game::Game`
// (blokus/src/game.rs:91-312) and of the self_play crate's client (self_play/src/simulation.rs:14-22,267-296,
// self_play/src/lib.rs:9-32) over the C ABI of blokus_b200.h.  Value semantics like the Rust type
// (`Game: Clone`): copying a Game clones the device state (bk_env_clone); Err(String) becomes
// blokus::Error carrying bk_last_error().  A Game is a batch of one; GameBatch exposes n games.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "blokus_b200.h"

namespace blokus {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc < 0) throw Error(rc, bk_last_error());
}

class GameBatch {
public:
    explicit GameBatch(int n_games, int device = 0) : n_(n_games) { check(bk_env_create(n_games, device, &env_)); }
    GameBatch(const GameBatch& o) : n_(o.n_) { check(bk_env_clone(o.env_, &env_)); }
    GameBatch(GameBatch&& o) noexcept : n_(o.n_), env_(o.env_) { o.env_ = nullptr; }
    GameBatch& operator=(GameBatch o) { std::swap(env_, o.env_); std::swap(n_, o.n_); return *this; }
    ~GameBatch() { bk_env_destroy(env_); }

    int size() const { return n_; }
    bk_env* handle() const { return env_; }
    void reset() { check(bk_env_reset(env_)); }                                              // game.rs:102
    void apply(const std::vector<int32_t>& tiles) { check(bk_env_apply(env_, tiles.data(), nullptr, nullptr)); }
    void apply(const std::vector<int32_t>& tiles, const std::vector<int32_t>& piece_to_finish) {
        check(bk_env_apply(env_, tiles.data(), piece_to_finish.data(), nullptr));             // game.rs:150
    }
    void place_piece(const std::vector<int32_t>& p, const std::vector<int32_t>& v, const std::vector<int32_t>& o) {
        check(bk_env_place_piece(env_, p.data(), v.data(), o.data(), nullptr));               // game.rs:116
    }
    std::vector<uint8_t> legal_mask() const { return bytes(bk_env_legal_mask, 400); }        // game.rs:242
    // Game::get_legal_tiles of game g in the reference's own shape (Vec<usize>; ascending here), bk_env_legal_tiles
    std::vector<int> legal_tiles(int g) const {
        const size_t n = static_cast<size_t>(n_);
        std::vector<int32_t> cnt(n);
        std::vector<int16_t> tiles(n * 400);
        check(bk_env_legal_tiles(env_, cnt.data(), tiles.data()));
        return std::vector<int>(tiles.begin() + long(g) * 400, tiles.begin() + long(g) * 400 + cnt[size_t(g)]);
    }
    std::vector<uint8_t> board() const { return bytes(bk_env_board, 400); }                  // game.rs:196
    std::vector<uint8_t> board_state() const { return bytes(bk_env_board_state, 2000); }     // game.rs:283
    std::vector<uint8_t> anchors(int player = -1) const {                                    // game.rs:238
        std::vector<uint8_t> out(static_cast<size_t>(n_) * 400);
        check(bk_env_anchors(env_, player, out.data()));
        return out;
    }
    std::vector<int32_t> current_player() const { return ints(bk_env_current_player, 1); }   // game.rs:225
    std::vector<int32_t> is_terminal() const { return ints(bk_env_is_terminal, 1); }         // game.rs:275
    std::vector<int32_t> is_player_active() const { return ints(bk_env_is_player_active, 4); }
    std::vector<int32_t> scores() const { return ints(bk_env_scores, 4); }                   // game.rs:247
    std::vector<float> payoff() const {                                                      // game.rs:252
        std::vector<float> out(static_cast<size_t>(n_) * 4);
        check(bk_env_payoff(env_, out.data()));
        return out;
    }
    // Game::history of game g as (player, tile) pairs (game.rs:94)
    std::vector<std::pair<int, int>> history(int g) const {
        const size_t n = static_cast<size_t>(n_);
        std::vector<int32_t> cnt(n), pl(n * BK_MAX_PLIES), tl(n * BK_MAX_PLIES);
        check(bk_env_history(env_, cnt.data(), pl.data(), tl.data()));
        std::vector<std::pair<int, int>> out;
        for (int i = 0; i < cnt[size_t(g)]; ++i)
            out.emplace_back(pl[size_t(g) * BK_MAX_PLIES + size_t(i)], tl[size_t(g) * BK_MAX_PLIES + size_t(i)]);
        return out;
    }
    // lockstep playout on the device (BASELINE.json configs 1-2)
    void playout(uint64_t seed, uint32_t first_game_id = 0, int max_plies = -1, uint32_t flags = 0) {
        check(bk_env_playout(env_, seed, first_game_id, max_plies, flags));
    }

private:
    template <class F>
    std::vector<uint8_t> bytes(F fn, size_t per) const {
        std::vector<uint8_t> out(static_cast<size_t>(n_) * per);
        check(fn(env_, out.data()));
        return out;
    }
    template <class F>
    std::vector<int32_t> ints(F fn, size_t per) const {
        std::vector<int32_t> out(static_cast<size_t>(n_) * per);
        check(fn(env_, out.data()));
        return out;
    }
    int n_ = 0;
    bk_env* env_ = nullptr;
};

// One game with the reference's method names.
class Game {
public:
    static Game reset(int device = 0) { return Game(device); }                               // game.rs:102
    explicit Game(int device = 0) : b_(1, device) {}
    void apply(int tile) { b_.apply({tile}); }                                               // apply(tile, None)
    void apply(int tile, int piece_to_finish) { b_.apply({tile}, {piece_to_finish}); }       // apply(tile, Some(p))
    Game place_piece(int p, int v, int o) const { Game ns = *this; ns.b_.place_piece({p}, {v}, {o}); return ns; }
    std::vector<uint8_t> get_board() const { return b_.board(); }
    int current_player() const { return b_.current_player()[0]; }
    std::vector<int> get_legal_tiles() const { return b_.legal_tiles(0); }                  // game.rs:242 (ascending)
    std::vector<int> get_current_anchors() const {
        std::vector<int> out;
        const auto m = b_.anchors(-1);
        for (int t = 0; t < 400; ++t) if (m[size_t(t)]) out.push_back(t);
        return out;
    }
    std::vector<int32_t> get_score() const { return b_.scores(); }
    std::vector<float> get_payoff() const { return b_.payoff(); }
    bool is_terminal() const { return b_.is_terminal()[0] != 0; }
    bool is_player_active(int player) const { return b_.is_player_active()[size_t(player)] != 0; }
    std::vector<uint8_t> get_board_state() const { return b_.board_state(); }
    std::vector<std::pair<int, int>> history() const { return b_.history(0); }
    // ids of the mover's remaining pieces in list order (game.rs:230; board.rs:147-153: indices into this list are
    // what place_piece / apply(Some(p)) take)
    std::vector<int> get_current_player_pieces() const { return remaining(current_player()); }
    // PieceVariant of the `piece`-th remaining piece of `player` (game.rs:234): offsets of its squares and its width
    struct PieceVariant { std::vector<int> offsets; int width = 0, len = 0, piece_id = 0; };
    PieceVariant get_piece(int player, int piece, int variant) const {
        const std::vector<int> ids = remaining(player);
        if (piece < 0 || size_t(piece) >= ids.size()) throw Error(BK_ERR_INVALID_ARG, "piece index out of range");
        PieceVariant pv;
        int offs[5];
        const int k = bk_piece_variant(ids[size_t(piece)], variant, &pv.width, &pv.len, offs);
        if (k < 0) throw Error(k, bk_last_error());
        pv.offsets.assign(offs, offs + k);
        pv.piece_id = ids[size_t(piece)];
        return pv;
    }
    GameBatch& batch() { return b_; }

private:
    std::vector<int> remaining(int player) const {
        std::vector<uint32_t> masks(4);
        check(bk_env_pieces(b_.handle(), masks.data()));
        std::vector<int> ids;
        for (int i = 0; i < 21; ++i) if ((masks[size_t(player)] >> i) & 1u) ids.push_back(i);
        return ids;
    }
    GameBatch b_;
};

// ---- self_play crate -------------------------------------------------------------------------------------------
namespace self_play {

// simulation.rs:14-22 (+ seed: the reference draws from thread_rng and cannot be seeded)
struct Config {
    uint32_t sims_per_move = 50, sample_moves = 30;
    float c_base = 19652.0f, c_init = 1.25f, dirichlet_alpha = 0.3f, exploration_fraction = 0.25f;
    uint64_t seed = 0;
    bk_config abi() const {
        return bk_config{sims_per_move, sample_moves, c_base, c_init, dirichlet_alpha, exploration_fraction, seed};
    }
};

// what training_game() returns (simulation.rs:293-295; lib.rs:15)
struct TrainingGame {
    std::vector<std::pair<int, int>> history;                 // (player, tile)
    std::vector<std::vector<std::pair<int, float>>> policies; // per ply: (tile, visits / total visits)
    std::vector<float> values;                                // payoff, absolute seat order
};

// The policy/value network `ResNet(blocks, 256)` (model/resnet.py:44-94, eval mode) resident on one device as a
// bk_evaluator: what the reference's inference server computes per batch (model/training.py:43-67), on this library's
// kernels.  Parameters are host arrays with the BatchNorms already folded (layout: blokus_b200.h, bk_evaluator_create).
class Evaluator {
public:
    Evaluator(int device, int blocks, int max_rows, const uint16_t* w_in, const float* b_in, const uint16_t* w_blocks,
              const float* b_blocks, const float* head_w, const float* head_affine, const float* lin_w, const float* lin_b) {
        check(bk_evaluator_create(device, blocks, max_rows, w_in, b_in, w_blocks, b_blocks, head_w, head_affine, lin_w, lin_b, &ev_));
    }
    Evaluator(const Evaluator&) = delete;
    Evaluator& operator=(const Evaluator&) = delete;
    ~Evaluator() { bk_evaluator_destroy(ev_); }
    bk_evaluator* handle() const { return ev_; }
    int max_rows() const { return bk_evaluator_max_rows(ev_); }
    // model(boards): device planes [rows][5][20][20] f32 -> device policy [rows][400], value [rows][4]
    void forward(const float* dev_planes, int rows, float* dev_policy, float* dev_value, void* cuda_stream = nullptr) {
        check(bk_evaluator_forward(ev_, dev_planes, rows, dev_policy, dev_value, nullptr, nullptr, cuda_stream));
    }

private:
    bk_evaluator* ev_ = nullptr;
};

// n self-play clients on one device; game g has global id first_game_id + g.
class SelfPlay {
public:
    SelfPlay(int n_games, const Config& cfg, uint32_t first_game_id = 0, int device = 0, uint32_t max_children_per_game = 0)
        : n_(n_games) {
        const bk_config c = cfg.abi();
        check(bk_selfplay_create(n_games, device, &c, first_game_id, max_children_per_game, &sp_));
    }
    SelfPlay(const SelfPlay&) = delete;
    SelfPlay& operator=(const SelfPlay&) = delete;
    ~SelfPlay() { bk_selfplay_destroy(sp_); }

    int size() const { return n_; }
    bk_selfplay* handle() const { return sp_; }
    void reset(uint32_t first_game_id) { check(bk_selfplay_reset(sp_, first_game_id)); }
    // opt-in throughput modes (BK_MODE_SKIP_FORCED, leaves per evaluator round); default = the reference's behaviour
    void set_mode(uint32_t flags, int leaves_per_round = 1) { check(bk_selfplay_set_mode(sp_, flags, leaves_per_round)); }
    // training_game() for every client with the fixed-prior stub evaluator, on the device (max_plies < 0: to the end)
    void run_stub(int max_plies = -1) { check(bk_selfplay_run_stub(sp_, max_plies)); }
    // training_game() with a caller-supplied evaluator working on DEVICE memory:
    //   eval(planes, policy, value, rows): planes[rows][5][20][20] f32 in; fill policy[rows][400] (mover frame) and
    //   value[rows][4] (relative seats) — the contract of the reference's inference server (model/training.py:43-67) on
    //   one contiguous batch.  rows = the positions waiting for an answer, dense in game order (at most n in the
    //   exact mode, n * leaves_per_round in the multi-leaf mode; the buffers must hold that many).
    template <class Eval>
    void run_evaluator(Eval&& eval, float* dev_planes, float* dev_policy, float* dev_value, int max_plies = -1) {
        int32_t live = 0;
        for (int ply = 0; max_plies < 0 || ply < max_plies; ++ply) {
            check(bk_selfplay_live_games(sp_, &live));
            if (live == 0) break;
            check(bk_selfplay_begin_ply(sp_));
            int32_t pending = 0;
            check(bk_selfplay_leaf_planes(sp_, dev_planes, &pending));
            while (pending > 0) {
                int32_t rows = 0;
                check(bk_selfplay_leaf_rows(sp_, &rows));
                if (rows > 0) eval(dev_planes, dev_policy, dev_value, int(rows));
                check(bk_selfplay_expand_backup(sp_, dev_policy, dev_value, &pending));
                if (pending > 0) check(bk_selfplay_leaf_planes(sp_, dev_planes, nullptr));
            }
            check(bk_selfplay_end_ply(sp_));
        }
    }
    // training_game() with the native network evaluator: the whole round stays inside the library (bk_selfplay_run_network)
    struct NetworkRun { int64_t rounds = 0, evals = 0; };
    NetworkRun run_network(Evaluator& ev, int max_plies = -1) {
        NetworkRun r;
        check(bk_selfplay_run_network(sp_, ev.handle(), max_plies, &r.rounds, &r.evals));
        return r;
    }
    std::vector<TrainingGame> results() const {
        const size_t n = static_cast<size_t>(n_);
        // one packed (compressed-row) gather of every game's policy records
        int64_t total_plies = 0, total_entries = 0;
        std::vector<int64_t> ply_off(n + 1);
        check(bk_selfplay_results_sizes(sp_, &total_plies, &total_entries, ply_off.data(), nullptr));
        std::vector<int64_t> ply_ptr(size_t(total_plies) + 1);
        std::vector<int16_t> tile(size_t(total_entries) + 1);
        std::vector<uint32_t> visits(size_t(total_entries) + 1);
        check(bk_selfplay_results_packed(sp_, ply_ptr.data(), tile.data(), visits.data()));
        bk_env* env = bk_selfplay_env(sp_);
        std::vector<int32_t> cnt(n), pl(n * BK_MAX_PLIES), tl(n * BK_MAX_PLIES);
        check(bk_env_history(env, cnt.data(), pl.data(), tl.data()));
        std::vector<float> pay(n * 4);
        check(bk_env_payoff(env, pay.data()));
        std::vector<TrainingGame> out(n);
        for (size_t g = 0; g < n; ++g) {
            TrainingGame& t = out[g];
            for (int i = 0; i < cnt[g]; ++i) t.history.emplace_back(pl[g * BK_MAX_PLIES + size_t(i)], tl[g * BK_MAX_PLIES + size_t(i)]);
            for (int64_t k = ply_off[g]; k < ply_off[g + 1]; ++k) {
                const int64_t a = ply_ptr[size_t(k)], b = ply_ptr[size_t(k) + 1];
                uint32_t total = 0;
                for (int64_t e = a; e < b; ++e) total += visits[size_t(e)];
                std::vector<std::pair<int, float>> pol;
                for (int64_t e = a; e < b; ++e)                                           // simulation.rs:222
                    pol.emplace_back(int(tile[size_t(e)]), float(visits[size_t(e)]) / float(total));
                t.policies.push_back(std::move(pol));
            }
            t.values.assign(pay.begin() + long(g) * 4, pay.begin() + long(g) * 4 + 4);
        }
        return out;
    }

private:
    int n_ = 0;
    bk_selfplay* sp_ = nullptr;
};

// batched play_training_game (lib.rs:9-32) with the stub evaluator: games first_game_id .. first_game_id + n - 1
inline std::vector<TrainingGame> play_training_games(uint32_t first_game_id, int n_games, const Config& cfg, int device = 0) {
    SelfPlay sp(n_games, cfg, first_game_id, device);
    sp.run_stub(-1);
    return sp.results();
}

}  // namespace self_play

}  // namespace blokus
