// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
// C entry points over the restatement so that tests/ and bench.py's cpu_baseline leg can drive it
// through ctypes (oracle/oracle.py).  Nothing under blokus-engine_b200/ links or loads this.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "blokus_oracle.hpp"
#include "mcts_oracle.hpp"
#include "rng_oracle.hpp"

using namespace orc;

namespace {

Shape shape_from_flat(const uint8_t* cells, int rows, int cols) {
    Shape s(rows, std::vector<bool>(cols));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) s[r][c] = cells[r * cols + c] != 0;
    return s;
}

void shape_to_flat(const Shape& s, uint8_t* out, int* rows, int* cols) {
    *rows = int(s.size());
    *cols = int(s[0].size());
    for (size_t r = 0; r < s.size(); ++r)
        for (size_t c = 0; c < s[r].size(); ++c) out[r * s[r].size() + c] = s[r][c];
}

// state digest used by trace hashes; the product computes the same words from its bitboards
uint64_t state_digest(const Game& g) {
    uint64_t sum = 0;
    uint32_t rows[5][20];
    std::memset(rows, 0, sizeof(rows));
    for (size_t i = 0; i < 400; ++i) {
        uint8_t owner = g.board.board[i] & 0x0F;
        if (owner) rows[owner - 1][i / 20] |= 1u << (i % 20);
    }
    for (size_t t : g.get_legal_tiles()) rows[4][t / 20] |= 1u << (t % 20);
    for (uint32_t b = 0; b < 5; ++b)
        for (uint32_t r = 0; r < 20; ++r)
            sum += splitmix64((uint64_t(b * 32 + r) << 32) | rows[b][r]);
    for (uint32_t p = 0; p < 4; ++p) {
        uint32_t mask = 0;
        for (const Piece& pc : g.board.pieces[p]) mask |= 1u << pc.id;
        sum += splitmix64((uint64_t(5 * 32 + p) << 32) | mask);
        sum += splitmix64((uint64_t(6 * 32 + p) << 32) | g.last_piece_lens[p]);
    }
    uint32_t meta = uint32_t(g.current_player());
    for (uint32_t p = 0; p < 4; ++p) meta |= (g.eliminated[p] ? 1u : 0u) << (4 + p);
    sum += splitmix64((uint64_t(7 * 32) << 32) | meta);
    return sum;
}

struct PlayoutOut {
    int n_plies = 0;
    uint64_t hash = 0;
    int scores[4] = {0, 0, 0, 0};
    float payoff[4] = {0, 0, 0, 0};
    int movegens = 0;
};

// policy: 0 = seeded uniform over ascending legal tiles, 1 = smallest tile, 2 = largest tile
PlayoutOut run_playout(uint64_t seed, uint32_t game_id, int policy, long max_plies, int16_t* tiles,
                       int8_t* players, int32_t* legal_counts, int want_hash) {
    Game g = Game::reset();
    PlayoutOut out;
    uint64_t h = 0;
    while (!g.is_terminal()) {
        if (max_plies >= 0 && out.n_plies >= max_plies) break;
        std::vector<size_t> legal = g.get_legal_tiles();
        size_t pick = 0;
        if (policy == 0) pick = playout_index(seed, game_id, uint32_t(out.n_plies), uint32_t(legal.size()));
        else if (policy == 2) pick = legal.size() - 1;
        size_t tile = legal[pick];
        int mover = int(g.current_player());
        if (legal_counts) legal_counts[out.n_plies] = int(legal.size());
        (void)g.apply(tile, -1);
        if (tiles) tiles[out.n_plies] = int16_t(tile);
        if (players) players[out.n_plies] = int8_t(mover);
        if (want_hash) {
            h = splitmix64(h ^ state_digest(g));
            h = splitmix64(h ^ (uint64_t(mover) | (uint64_t(tile) << 8)));
        }
        out.n_plies += 1;
    }
    out.hash = h;
    std::vector<int> sc = g.get_score();
    std::vector<float> pay = g.get_payoff();
    for (int i = 0; i < 4; ++i) { out.scores[i] = sc[i]; out.payoff[i] = pay[i]; }
    return out;
}

}  // namespace

extern "C" {

// ---- pieces.rs KAT surface -------------------------------------------------------------------
int orc_num_piece_types() { return int(NUM_PIECE_TYPES); }
int orc_piece_points(int piece_type) { return int(Piece(size_t(piece_type)).points); }
int orc_piece_num_variants(int piece_type) { return int(Piece(size_t(piece_type)).variants.size()); }
// variant data of a stock piece: returns number of offsets; width/len out; offsets out[<=5]
int orc_piece_variant(int piece_type, int var, int* width, int* len, int* offsets) {
    Piece p{size_t(piece_type)};
    const PieceVariant& v = p.variants.at(size_t(var));
    *width = int(v.width);
    *len = int(v.variant.size());
    for (size_t i = 0; i < v.offsets.size(); ++i) offsets[i] = int(v.offsets[i]);
    return int(v.offsets.size());
}
int orc_gen_variants_count(const uint8_t* cells, int rows, int cols) {
    return int(Piece::gen_variants(shape_from_flat(cells, rows, cols)).size());
}
// PieceVariant::new(shape): variant vector (out_variant, up to 81), offsets, width; returns len
int orc_variant_new(const uint8_t* cells, int rows, int cols, uint8_t* out_variant, int* out_offsets,
                    int* n_offsets, int* width) {
    PieceVariant v(shape_from_flat(cells, rows, cols));
    for (size_t i = 0; i < v.variant.size(); ++i) out_variant[i] = v.variant[i];
    for (size_t i = 0; i < v.offsets.size(); ++i) out_offsets[i] = int(v.offsets[i]);
    *n_offsets = int(v.offsets.size());
    *width = int(v.width);
    return int(v.variant.size());
}
void orc_rotate(const uint8_t* cells, int rows, int cols, uint8_t* out, int* orows, int* ocols) {
    shape_to_flat(Piece::rotate(shape_from_flat(cells, rows, cols)), out, orows, ocols);
}
void orc_flip(const uint8_t* cells, int rows, int cols, uint8_t* out, int* orows, int* ocols) {
    shape_to_flat(Piece::flip(shape_from_flat(cells, rows, cols)), out, orows, ocols);
}
// board.rs:220-225: is_valid_move of an ad-hoc shape on a fresh board
int orc_fresh_board_is_valid(int player, const uint8_t* cells, int rows, int cols, int offset) {
    Board b;
    return b.is_valid_move(size_t(player), PieceVariant(shape_from_flat(cells, rows, cols)), size_t(offset)) ? 1 : 0;
}
int orc_fresh_board_len() { Board b; return int(b.board.size()); }

// ---- Game handle -----------------------------------------------------------------------------
void* orc_game_new() { return new Game(Game::reset()); }
void orc_game_free(void* g) { delete static_cast<Game*>(g); }
void* orc_game_clone(void* g) { return new Game(*static_cast<Game*>(g)); }
int orc_game_apply(void* g, int tile, int piece_to_finish) {
    return static_cast<Game*>(g)->apply(size_t(tile), long(piece_to_finish)).empty() ? 0 : -1;
}
// returns 0 and replaces *g on success, -1 on Err (g untouched, as place_piece works on a clone)
int orc_game_place_piece(void* g, int p, int v, int o) {
    Game* game = static_cast<Game*>(g);
    if (p < 0 || size_t(p) >= game->board.pieces[game->current_player()].size()) return -2;
    if (v < 0 || size_t(v) >= game->board.pieces[game->current_player()][size_t(p)].variants.size()) return -2;
    auto res = game->place_piece(size_t(p), size_t(v), size_t(o));
    if (!res.second.empty()) return -1;
    *game = res.first;
    return 0;
}
int orc_game_legal_tiles(void* g, int* out) {
    auto v = static_cast<Game*>(g)->get_legal_tiles();
    for (size_t i = 0; i < v.size(); ++i) out[i] = int(v[i]);
    return int(v.size());
}
int orc_game_num_placements(void* g) {
    std::set<Move> all;
    for (const auto& kv : static_cast<Game*>(g)->legal_tiles) all.insert(kv.second.begin(), kv.second.end());
    return int(all.size());
}
void orc_game_board(void* g, uint8_t* out) { std::memcpy(out, static_cast<Game*>(g)->get_board().data(), 400); }
int orc_game_current_player(void* g) { return int(static_cast<Game*>(g)->current_player()); }
int orc_game_is_terminal(void* g) { return static_cast<Game*>(g)->is_terminal() ? 1 : 0; }
int orc_game_is_player_active(void* g, int p) { return static_cast<Game*>(g)->is_player_active(size_t(p)) ? 1 : 0; }
void orc_game_scores(void* g, int* out) { auto s = static_cast<Game*>(g)->get_score(); for (int i = 0; i < 4; ++i) out[i] = s[i]; }
void orc_game_payoff(void* g, float* out) { auto s = static_cast<Game*>(g)->get_payoff(); for (int i = 0; i < 4; ++i) out[i] = s[i]; }
void orc_game_last_piece_lens(void* g, int* out) { for (int i = 0; i < 4; ++i) out[i] = int(static_cast<Game*>(g)->last_piece_lens[i]); }
void orc_game_board_state(void* g, uint8_t* out) {
    Planes p = static_cast<Game*>(g)->get_board_state();
    for (size_t a = 0; a < 5; ++a) for (size_t r = 0; r < 20; ++r) for (size_t c = 0; c < 20; ++c) out[(a * 20 + r) * 20 + c] = p[a][r][c];
}
int orc_game_anchors(void* g, int player, int* out) {
    Game* game = static_cast<Game*>(g);
    auto a = game->board.get_anchors(player < 0 ? game->current_player() : size_t(player));
    std::set<size_t> s(a.begin(), a.end());
    int n = 0;
    for (size_t t : s) out[n++] = int(t);
    return n;
}
// ids of the pieces a player still holds, in remaining-list order
int orc_game_pieces(void* g, int player, int* out) {
    Game* game = static_cast<Game*>(g);
    int n = 0;
    for (const Piece& p : game->board.pieces[size_t(player)]) out[n++] = int(p.id);
    return n;
}
int orc_game_history(void* g, int* players, int* tiles) {
    Game* game = static_cast<Game*>(g);
    for (size_t i = 0; i < game->history.size(); ++i) { players[i] = game->history[i].first; tiles[i] = game->history[i].second; }
    return int(game->history.size());
}
uint64_t orc_game_digest(void* g) { return state_digest(*static_cast<Game*>(g)); }

// ---- playouts (BASELINE.json configs 1 and 2) ---------------------------------------------------
int orc_playout(uint64_t seed, uint32_t game_id, int policy, int max_plies, int16_t* tiles, int8_t* players,
                int32_t* legal_counts, int* scores, float* payoff, uint64_t* hash) {
    PlayoutOut o = run_playout(seed, game_id, policy, max_plies, tiles, players, legal_counts, hash != nullptr);
    for (int i = 0; i < 4; ++i) { if (scores) scores[i] = o.scores[i]; if (payoff) payoff[i] = o.payoff[i]; }
    if (hash) *hash = o.hash;
    return o.n_plies;
}

// n_games playouts spread over n_threads host threads; returns wall seconds, total steps in *steps.
// hashes/plies/scores are optional per-game outputs.
double orc_playout_batch(uint64_t seed, uint32_t first_game, int n_games, int n_threads, int want_hash,
                         uint64_t* hashes, int32_t* plies, int32_t* scores, int64_t* steps) {
    std::atomic<int> next{0};
    std::atomic<int64_t> total{0};
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_games) break;
            PlayoutOut o = run_playout(seed, first_game + uint32_t(i), 0, -1, nullptr, nullptr, nullptr, want_hash);
            if (hashes) hashes[i] = o.hash;
            if (plies) plies[i] = o.n_plies;
            if (scores) for (int k = 0; k < 4; ++k) scores[i * 4 + k] = o.scores[k];
            total += o.n_plies;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (steps) *steps = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- MCTS self-play (BASELINE.json config 3) ------------------------------------------------------
struct OrcConfig {
    uint32_t sims_per_move, sample_moves;
    float c_base, c_init, dirichlet_alpha, exploration_fraction;
    uint64_t seed;
};
typedef void (*orc_eval_fn)(void* user, int id, const uint8_t* planes, float* policy, float* value);

static Config to_cfg(const OrcConfig* c) {
    return Config{c->sims_per_move, c->sample_moves, c->c_base, c->c_init, c->dirichlet_alpha,
                  c->exploration_fraction, c->seed};
}

// One self-play game with the stub evaluator (eval == NULL) or a caller-supplied one.
// Outputs (all optional except n_plies return): history players/tiles [max 400]; per-ply root
// children flattened: root_off[n_plies+1], then tile/visits/value_sum/prior arrays of root_off[n]
// entries (capacity root_cap); payoff[4]; sims.  Returns n_plies, or -1 if root_cap is too small.
int orc_selfplay_game(const OrcConfig* c, int game_id, int max_plies, orc_eval_fn eval, void* user,
                      int* players, int* tiles, int* root_off, int root_cap, int* r_tile,
                      uint32_t* r_visits, float* r_value_sum, float* r_prior, float* payoff, int64_t* sims) {
    Evaluator ev = stub_evaluator;
    if (eval) {
        ev = [eval, user](int id, const Planes& p, std::vector<float>& pol, std::vector<float>& val) {
            uint8_t flat[2000];
            for (size_t a = 0; a < 5; ++a) for (size_t r = 0; r < 20; ++r) for (size_t cc = 0; cc < 20; ++cc) flat[(a * 20 + r) * 20 + cc] = p[a][r][cc];
            pol.assign(400, 0.0f);
            val.assign(4, 0.0f);
            eval(user, id, flat, pol.data(), val.data());
        };
    }
    GameRecord rec = training_game(to_cfg(c), ev, game_id, max_plies);
    const int n = int(rec.policies.size());
    int off = 0;
    for (int i = 0; i < n; ++i) {
        if (root_off) root_off[i] = off;
        for (const RootChild& rc : rec.roots[size_t(i)]) {
            if (off >= root_cap) return -1;
            if (r_tile) r_tile[off] = rc.tile;
            if (r_visits) r_visits[off] = rc.visits;
            if (r_value_sum) r_value_sum[off] = rc.value_sum;
            if (r_prior) r_prior[off] = rc.prior;
            ++off;
        }
    }
    if (root_off) root_off[n] = off;
    for (size_t i = 0; i < rec.history.size(); ++i) {
        if (players) players[i] = rec.history[i].first;
        if (tiles) tiles[i] = rec.history[i].second;
    }
    if (payoff) for (int i = 0; i < 4; ++i) payoff[i] = rec.values[size_t(i)];
    if (sims) *sims = int64_t(rec.sims);
    return n;
}

// play_test_game (simulation.rs:298-332) with two caller-supplied evaluators; returns payoff[0]
float orc_test_game(int game_id, uint64_t seed, orc_eval_fn model, orc_eval_fn baseline, void* user, int* players,
                    int* tiles, int* n_plies) {
    auto wrap = [user](orc_eval_fn fn) {
        return Evaluator([fn, user](int id, const Planes& p, std::vector<float>& pol, std::vector<float>& val) {
            uint8_t flat[2000];
            for (size_t a = 0; a < 5; ++a) for (size_t r = 0; r < 20; ++r) for (size_t cc = 0; cc < 20; ++cc) flat[(a * 20 + r) * 20 + cc] = p[a][r][cc];
            pol.assign(400, 0.0f);
            val.assign(4, 0.0f);
            fn(user, id, flat, pol.data(), val.data());
        });
    };
    std::vector<std::pair<int, int>> hist;
    float r = test_game(game_id, wrap(model), wrap(baseline), seed, &hist);
    for (size_t i = 0; i < hist.size(); ++i) { if (players) players[i] = hist[i].first; if (tiles) tiles[i] = hist[i].second; }
    if (n_plies) *n_plies = int(hist.size());
    return r;
}

// n_games stub-evaluator games over n_threads threads, each cut at max_plies; returns wall seconds.
double orc_selfplay_batch(const OrcConfig* c, int first_game, int n_games, int n_threads, int max_plies,
                          int64_t* sims) {
    std::atomic<int> next{0};
    std::atomic<int64_t> total{0};
    Config cfg = to_cfg(c);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_games) break;
            GameRecord rec = training_game(cfg, stub_evaluator, first_game + i, max_plies);
            total += int64_t(rec.sims);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (sims) *sims = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- RNG spec KAT surface ---------------------------------------------------------------------
void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) {
    Philox4 b = philox4x32_10(seed, c0, c1, c2, c3);
    for (int i = 0; i < 4; ++i) out[i] = b.v[i];
}
uint32_t orc_playout_index(uint64_t seed, uint32_t game, uint32_t ply, uint32_t n) { return playout_index(seed, game, ply, n); }
float orc_action_uniform(uint64_t seed, uint32_t game, uint32_t ply) { return action_uniform(seed, game, ply); }
void orc_dirichlet(uint64_t seed, uint32_t game, uint32_t ply, uint32_t n, float alpha, float* out) {
    auto v = dirichlet_noise(seed, game, ply, n, alpha);
    for (uint32_t i = 0; i < n; ++i) out[i] = v[i];
}
double orc_det_log(double x) { return det_log(x); }
double orc_det_exp(double x) { return det_exp(x); }
uint64_t orc_splitmix64(uint64_t x) { return splitmix64(x); }
// UCB factor table the product uploads: (ln((N + c_base + 1)/c_base) + c_init) * sqrt(N), f32,
// simulation.rs:91-93 evaluated with the host libm exactly as ucb_score() above does.
void orc_ucb_factor_table(float c_base, float c_init, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        float pv = float(i);
        out[i] = (std::log((pv + c_base + 1.0f) / c_base) + c_init) * std::sqrt(pv);
    }
}

}  // extern "C"
