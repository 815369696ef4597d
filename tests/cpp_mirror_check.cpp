// Compiled and run by tests/test_cpp_mirror.py: the C++ mirror of `Game` over the C ABI.
// Without a CUDA device construction must throw (no CPU fallback); with one, the seed-free "smallest legal
// tile" game of SURVEY.md Appendix C must come out: 314 plies, scores [15,-35,-4,-3].
#include <cstdio>

#include "blokus_b200.hpp"

int main() {
    if (bk_device_count() == 0) {
        try {
            blokus::Game g = blokus::Game::reset();
            std::printf("FAIL: constructed a game without a device\n");
            return 1;
        } catch (const blokus::Error& e) {
            std::printf("no-device ok: [%d] %s\n", e.code, e.what());
            return e.code == BK_ERR_CUDA ? 0 : 1;
        }
    }
    blokus::Game g = blokus::Game::reset();
    int plies = 0;
    while (!g.is_terminal()) { g.apply(g.get_legal_tiles().front()); ++plies; }
    const auto sc = g.get_score();
    std::printf("plies %d scores %d %d %d %d\n", plies, sc[0], sc[1], sc[2], sc[3]);
    bool ok = plies == 314 && sc[0] == 15 && sc[1] == -35 && sc[2] == -4 && sc[3] == -3 && g.history().size() == 314;
    blokus::Game h = blokus::Game::reset();
    try { h.apply(399); ok = false; } catch (const blokus::Error& e) { ok = ok && e.code == BK_ERR_ILLEGAL_MOVE; }
    ok = ok && h.get_current_player_pieces().size() == 21;
    const auto domino_v = h.get_piece(0, 1, 1);           // vertical domino: offsets {0, 20}, width 1 (pieces.rs:240-251)
    ok = ok && domino_v.offsets.size() == 2 && domino_v.offsets[1] == 20 && domino_v.width == 1 && domino_v.len == 21;
    blokus::Game moved = h.place_piece(0, 0, 0);          // monomino on the start corner; h itself is untouched
    ok = ok && moved.get_piece(0, 0, 0).piece_id == 1;    // player 0's list now starts with the domino
    ok = ok && h.history().empty() && moved.history().size() == 1 && moved.current_player() == 1;
    // self_play crate mirror: two stub self-play games, 24 simulations a move, six plies
    blokus::self_play::Config cfg;
    cfg.sims_per_move = 24; cfg.sample_moves = 3; cfg.seed = 5;
    blokus::self_play::SelfPlay sp(2, cfg, /*first_game_id=*/7);
    sp.set_mode(BK_MODE_SKIP_FORCED, 1);
    sp.run_stub(6);
    const auto games = sp.results();
    ok = ok && games.size() == 2;
    for (const auto& tg : games) {
        ok = ok && tg.history.size() == 6 && tg.policies.size() == 6 && tg.values.size() == 4;
        for (size_t k = 0; k < tg.policies.size(); ++k) {
            float sum = 0.0f;
            bool has_action = false;
            for (const auto& tp : tg.policies[k]) { sum += tp.second; has_action = has_action || tp.first == tg.history[k].second; }
            ok = ok && sum > 0.999f && sum < 1.001f && has_action;
        }
    }
    std::printf(ok ? "device ok\n" : "FAIL\n");
    return ok ? 0 : 1;
}
