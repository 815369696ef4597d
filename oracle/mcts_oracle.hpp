// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// CPU restatement of the reference's `self_play` crate (self_play/src/node.rs:8-41,
// self_play/src/simulation.rs:14-231,267-296): one game of MCTS self-play, one simulation in
// flight, a fresh tree every ply, the game cloned per simulation.  Control flow and f32 expression
// order are kept exactly; the three non-reproducible ingredients are canonicalised as SURVEY.md
// Appendix D lays out:
//   * HashMap iteration  -> ascending tile order (std::map), so `>=` / `max_by` resolve ties to
//     the highest tile index;
//   * rand::thread_rng() -> the seeded SPEC of rng_oracle.hpp;
//   * the Python queue/pipe round trip (simulation.rs:50-57) -> an Evaluator callback with the same
//     frames (planes rotated to the mover, policy in that frame, value in relative-seat order).
// "Parity unpinned": the reference has no tests for this crate.  The noise-free core (evaluate,
// ucb_score, select_child, backpropagate, mcts) is cross-checked bit for bit against an independently written second
// restatement (oracle/py_restatement.py, tests/test_py_restatement.py).
#pragma once
#include <cmath>
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "blokus_oracle.hpp"
#include "rng_oracle.hpp"

namespace orc {

// simulation.rs:14-22 (+ seed, which the reference does not have)
struct Config {
    size_t sims_per_move;
    size_t sample_moves;
    float c_base;
    float c_init;
    float dirichlet_alpha;
    float exploration_fraction;
    uint64_t seed;
};

// node.rs:8-41
struct Node {
    std::map<size_t, Node> children;
    size_t to_play = 0;
    float value_sum = 0.0f;
    uint32_t visits = 0;
    float prior;
    explicit Node(float p = 0.0f) : prior(p) {}
    bool is_expanded() const { return !children.empty(); }
    float value() const { return visits == 0 ? 0.0f : value_sum / float(visits); }
};

// planes[5][20][20] (mover frame) -> policy[400] (mover frame), value[4] (relative seat)
using Evaluator = std::function<void(int id, const Planes&, std::vector<float>&, std::vector<float>&)>;

// The fixed-prior stub of BASELINE.json config 3: what model/resnet.py:84-92 would return for a
// network whose masked softmax is replaced by a constant: 1.0 on legal tiles (plane 4), else 0;
// value 0.25 for every seat.
inline void stub_evaluator(int, const Planes& planes, std::vector<float>& policy,
                           std::vector<float>& value) {
    policy.assign(400, 0.0f);
    for (size_t r = 0; r < 20; ++r)
        for (size_t c = 0; c < 20; ++c)
            if (planes[4][r][c]) policy[r * 20 + c] = 1.0f;
    value.assign(4, 0.25f);
}

// f32 exp.  The reference calls f32::exp (simulation.rs:71), i.e. the platform libm's expf, whose last
// bit is not specified (glibc documents 0.502 ULP).  Canonical form used by oracle AND kernels: exp in
// f64, rounded once to f32 — correctly rounded except for ~2^-29 double-rounding cases, and the same on
// a CPU and a GPU for the arguments that occur.
inline float exp_f32(float x) { return float(std::exp(double(x))); }

// simulation.rs:25-34
inline std::vector<float> rotate_policy(const std::vector<float>& state) {
    std::vector<float> rotated(400, 0.0f);
    for (size_t i = 0; i < 20; ++i)
        for (size_t j = 0; j < 20; ++j) rotated[j * 20 + (20 - 1 - i)] = state[i * 20 + j];
    return rotated;
}

// simulation.rs:37-83
inline std::vector<float> evaluate(Node& node, const Game& game, const Evaluator& ev, int id) {
    if (game.is_terminal()) return game.get_payoff();
    Planes rep = game.get_board_state();
    std::vector<float> policy, value;
    ev(id, rep, policy, value);
    const size_t cur = game.current_player();
    for (size_t k = 0; k < cur; ++k) policy = rotate_policy(policy);
    {   // value.rotate_right(cur)
        std::vector<float> v(4);
        for (size_t i = 0; i < 4; ++i) v[(i + cur) % 4] = value[i];
        value = v;
    }
    std::vector<std::pair<size_t, float>> exp_policy;
    for (size_t tile : game.get_legal_tiles())
        if (policy[tile] > 0.0f) exp_policy.emplace_back(tile, exp_f32(policy[tile]));
    float total = 0.0f;
    for (const auto& tp : exp_policy) total += tp.second;
    node.to_play = cur;
    for (const auto& tp : exp_policy) node.children.emplace(tp.first, Node(tp.second / total));
    return value;
}

// simulation.rs:88-98
inline float ucb_score(const Node& parent, const Node& child, const Config& cfg) {
    const float c_base = cfg.c_base, c_init = cfg.c_init;
    const float parent_visits = float(parent.visits);
    const float exploration_constant =
        (std::log((parent_visits + c_base + 1.0f) / c_base) + c_init) * std::sqrt(parent_visits) /
        (1.0f + float(child.visits));
    const float prior_score = exploration_constant * child.prior;
    const float value_score = child.value();
    return prior_score + value_score;
}

// simulation.rs:101-114
inline void add_exploration_noise(Node& root, const Config& cfg, uint32_t game, uint32_t ply) {
    const size_t n = root.children.size();
    if (n <= 1) return;
    std::vector<float> noise = dirichlet_noise(cfg.seed, game, ply, uint32_t(n), cfg.dirichlet_alpha);
    size_t i = 0;
    for (auto& kv : root.children) {
        Node& node = kv.second;
        node.prior = node.prior * (1.0f - cfg.exploration_fraction) + noise[i] * cfg.exploration_fraction;
        ++i;
    }
}

// simulation.rs:118-130
inline size_t softmax_sample(const std::vector<std::pair<size_t, uint32_t>>& dist, float sample) {
    uint32_t total = 0;
    for (const auto& tv : dist) total += tv.second;
    float sum = 0.0f;
    for (const auto& tv : dist) {
        sum += float(tv.second) / float(total);
        if (sum > sample) return tv.first;
    }
    return dist.back().first;
}

// simulation.rs:135-147
inline size_t select_child(const Node& node, const Config& cfg) {
    float best_score = 0.0f;
    size_t best_action = 0;
    for (const auto& kv : node.children) {
        float score = ucb_score(node, kv.second, cfg);
        if (score >= best_score) { best_score = score; best_action = kv.first; }
    }
    return best_action;
}

// simulation.rs:150-161
inline size_t select_action(const Node& root, size_t num_moves, const Config& cfg, uint32_t game,
                            uint32_t ply) {
    std::vector<std::pair<size_t, uint32_t>> dist;
    for (const auto& kv : root.children) dist.emplace_back(kv.first, kv.second.visits);
    if (num_moves < cfg.sample_moves) return softmax_sample(dist, action_uniform(cfg.seed, game, ply));
    size_t best = dist[0].first;  // max_by keeps the LAST maximum
    uint32_t bv = dist[0].second;
    for (const auto& tv : dist)
        if (tv.second >= bv) { bv = tv.second; best = tv.first; }
    return best;
}

// simulation.rs:164-171
inline void backpropagate(const std::vector<size_t>& path, Node& root, const std::vector<float>& values) {
    Node* node = &root;
    for (size_t tile : path) {
        node = &node->children.at(tile);
        node->visits += 1;
        node->value_sum += values[node->to_play];
    }
}

struct RootChild { int tile; uint32_t visits; float value_sum; float prior; };

struct GameRecord {
    std::vector<std::pair<int, int>> history;
    std::vector<std::vector<std::pair<int, float>>> policies;
    std::vector<float> values;
    std::vector<std::vector<RootChild>> roots;  // per-ply root children after the last simulation
    uint64_t sims = 0;
};

// simulation.rs:174-231
inline size_t mcts(const Game& game, GameRecord& rec, const Config& cfg, const Evaluator& ev, int id) {
    const uint32_t ply = uint32_t(game.history.size());
    Node root(0.0f);
    evaluate(root, game, ev, id);
    add_exploration_noise(root, cfg, uint32_t(id), ply);
    for (size_t s = 0; s < cfg.sims_per_move; ++s) {
        root.visits += 1;
        Node* node = &root;
        Game scratch = game;  // simulation.rs:196 clones the whole game
        std::vector<size_t> path;
        while (node->is_expanded()) {
            size_t action = select_child(*node, cfg);
            node = &node->children.at(action);
            (void)scratch.apply(action, -1);
            path.push_back(action);
        }
        std::vector<float> values = evaluate(*node, scratch, ev, id);
        backpropagate(path, root, values);
        rec.sims += 1;
    }
    uint32_t total = 0;
    for (const auto& kv : root.children) total += kv.second.visits;
    std::vector<std::pair<int, float>> probs;
    std::vector<RootChild> rc;
    for (const auto& kv : root.children) {
        probs.emplace_back(int(kv.first), float(kv.second.visits) / float(total));
        rc.push_back(RootChild{int(kv.first), kv.second.visits, kv.second.value_sum, kv.second.prior});
    }
    rec.policies.push_back(probs);
    rec.roots.push_back(rc);
    return select_action(root, rec.policies.size(), cfg, uint32_t(id), ply);
}

// simulation.rs:267-296; max_plies < 0 plays to the end.
inline GameRecord training_game(const Config& cfg, const Evaluator& ev, int id, long max_plies = -1) {
    Game game = Game::reset();
    GameRecord rec;
    while (!game.is_terminal()) {
        if (max_plies >= 0 && long(game.history.size()) >= max_plies) break;
        size_t action = mcts(game, rec, cfg, ev, id);
        (void)game.apply(action, -1);
    }
    rec.values = game.get_payoff();
    rec.history = game.history;
    return rec;
}

// simulation.rs:233-265.  The random baseline move draws gen_range(0..n) from thread_rng (:250) and takes
// keys().nth(index) of a HashMap; canonical form: ascending children, index from the seeded SPEC
// (purpose 3, same multiply-shift as the playout policy).
inline size_t best_action(const Game& game, const Evaluator& ev, int id, uint64_t seed) {
    Node root(0.0f);
    evaluate(root, game, ev, id);
    if (game.current_player() != 0) {
        const size_t n = root.children.size();
        Philox4 b = philox4x32_10(seed, uint32_t(id), uint32_t(game.history.size()), 3u, 0u);
        size_t index = size_t((uint64_t(b.v[0]) * n) >> 32);
        auto it = root.children.begin();
        std::advance(it, long(index));
        return it->first;
    }
    float highest_prior = 0.0f;
    size_t best = 0;
    for (const auto& kv : root.children)
        if (kv.second.prior > highest_prior) { highest_prior = kv.second.prior; best = kv.first; }
    return best;
}

// simulation.rs:298-332: seat 0 asks `model`, the other seats ask `baseline`; returns payoff[0]
inline float test_game(int id, const Evaluator& model, const Evaluator& baseline, uint64_t seed,
                       std::vector<std::pair<int, int>>* history_out = nullptr) {
    Game game = Game::reset();
    while (!game.is_terminal()) {
        const Evaluator& q = game.current_player() == 0 ? model : baseline;
        size_t action = best_action(game, q, id, seed);
        (void)game.apply(action, -1);
    }
    if (history_out) *history_out = game.history;
    return game.get_payoff()[0];
}

}  // namespace orc
