"""Derived known answers of SURVEY.md Appendix C (established by an independent throw-away Python
restatement during the survey) — the only pins available for game.rs, which has no reference tests."""
import numpy as np


def test_reset_position(orc):
    g = orc.Game()
    assert g.num_placements() == 58
    assert g.legal_tiles() == [0, 1, 2, 3, 4, 20, 21, 22, 23, 40, 41, 42, 60, 61, 80]
    assert g.current_player() == 0 and not g.is_terminal()
    assert g.anchors(0) == [0] and g.anchors(1) == [19] and g.anchors(2) == [399] and g.anchors(3) == [380]


def test_min_tile_trace(orc):
    r = orc.playout(0, 0, policy=1)
    assert r["n_plies"] == 314
    assert list(r["scores"]) == [15, -35, -4, -3]
    assert list(r["payoff"]) == [1.0, 0.0, 0.0, 0.0]
    assert r["tiles"][:20].tolist() == [0, 1, 2, 3, 4, 15, 16, 17, 18, 19, 319, 339, 359, 379, 399, 300, 320, 340, 360, 380]
    assert r["players"][:12].tolist() == [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2]


def test_max_tile_trace(orc):
    r = orc.playout(0, 0, policy=2)
    assert r["n_plies"] == 314
    assert list(r["scores"]) == [15, 15, 15, -42]
    assert np.allclose(r["payoff"], [1 / 3, 1 / 3, 1 / 3, 0])
    assert r["tiles"][:20].tolist() == [80, 60, 40, 20, 0, 99, 79, 59, 39, 19, 399, 398, 397, 396, 395, 384, 383, 382, 381, 380]


def test_last_piece_lens_of_traces(orc):
    for policy, lens in ((1, [4, 2, 3, 4]), (2, [4, 2, 4, 5])):
        g = orc.Game()
        while not g.is_terminal():
            lt = g.legal_tiles()
            g.apply(lt[0] if policy == 1 else lt[-1])
        assert g.last_piece_lens() == lens


def test_random_game_statistics(orc):
    """Plies/game stay in the survey's observed band (241..298 over 200 games; allow some slack)."""
    b = orc.playout_batch(7, 0, 40, n_threads=4, want_hash=False)
    assert 230 <= b["plies"].min() and b["plies"].max() <= 310
    assert abs(b["plies"].mean() - 272.9) < 8
