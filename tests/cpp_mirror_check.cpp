// Compiled and run by tests/test_cpp_mirror.py: the C++ mirror of `Game` over the C ABI.
// Without a CUDA device construction must throw (no CPU fallback); with one, the seed-free "smallest legal
// tile" game of SURVEY.md Appendix C must come out: 314 plies, scores [15,-35,-4,-3].
#include <cstdio>
#include <cstring>
#include <vector>

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
    // the native network evaluator behind the self_play client: ResNet(1, 256) with synthetic folded parameters, a few
    // plies of real self-play with the whole round inside the library.  The tensor-core kernel does not exist in the
    // tests' CPU emulator build, which must say so (BK_ERR_STATE) instead of computing something else.
    {
        std::vector<uint16_t> w_in(9 * 256 * 64, 0), w_blk(2 * 9 * 256 * 256, 0);
        std::vector<float> b_in(256, 0.01f), b_blk(2 * 256, 0.0f), head_w(512), head_aff = {1.0f, 0.1f, 1.0f, 0.1f}, lin_w(1600), lin_b(4, 0.0f);
        uint32_t x = 12345u;
        auto bf16 = [&](float scale) { x = x * 1664525u + 1013904223u; const float f = (float(x >> 8) / 8388608.0f - 1.0f) * scale;
                                       uint32_t u; std::memcpy(&u, &f, 4); return uint16_t(u >> 16); };
        for (size_t t = 0; t < 9; ++t) for (size_t o = 0; o < 256; ++o) for (size_t c = 0; c < 5; ++c) w_in[(t * 256 + o) * 64 + c] = bf16(0.15f);
        for (auto& w : w_blk) w = bf16(0.02f);
        for (size_t i = 0; i < 512; ++i) head_w[i] = 0.01f * float(int(i % 13) - 6);
        for (size_t i = 0; i < 1600; ++i) lin_w[i] = 0.002f * float(int(i % 7) - 3);
        blokus::self_play::Evaluator ev(0, 1, 8, w_in.data(), b_in.data(), w_blk.data(), b_blk.data(), head_w.data(), head_aff.data(),
                                        lin_w.data(), lin_b.data());
        blokus::self_play::Config c2;
        c2.sims_per_move = 16; c2.sample_moves = 2; c2.seed = 9;
        blokus::self_play::SelfPlay net(4, c2, 3);
        net.set_mode(0, 2);
        const bool emulated = std::strstr(bk_version(), "emulator") != nullptr;
        try {
            const auto run = net.run_network(ev, 3);
            ok = ok && !emulated && run.rounds > 0 && run.evals > 0 && ev.max_rows() == 8;
            for (const auto& tg : net.results()) {
                ok = ok && tg.history.size() == 3 && tg.policies.size() == 3;
                for (const auto& pol : tg.policies) { float sum = 0.0f; for (const auto& tp : pol) sum += tp.second; ok = ok && sum > 0.999f && sum < 1.001f; }
            }
        } catch (const blokus::Error& e) {
            ok = ok && emulated && e.code == BK_ERR_STATE;
        }
    }
    std::printf(ok ? "device ok\n" : "FAIL\n");
    return ok ? 0 : 1;
}
