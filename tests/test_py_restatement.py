"""Triangulation of the oracle: a second, independently written restatement of the reference (oracle/py_restatement.py,
Python dict/set transliteration of blokus/src/*.rs and the noise-free core of self_play/src/simulation.rs) must agree
with the C++ oracle on every ply of driven games and on every root of small searches.  The reference itself cannot
run here (Rust, no toolchain); two separately written readings agreeing is the strongest pin available."""
import numpy as np
import pytest

from oracle import py_restatement as R


def test_piece_tables_agree(orc):
    for pid in range(21):
        p = R.Piece(pid)
        assert p.points == orc.piece_points(pid)
        assert len(p.variants) == orc.piece_num_variants(pid)
        for v, pv in enumerate(p.variants):
            ov = orc.piece_variant(pid, v)
            assert (pv.width, len(pv.variant), pv.offsets) == (ov["width"], ov["len"], ov["offsets"]), (pid, v)


def _drive(orc, pick, max_plies=10**9, check_planes_every=7):
    a, b = orc.Game(), R.Game()
    ply = 0
    while not b.is_terminal() and ply < max_plies:
        la, lb = a.legal_tiles(), b.get_legal_tiles()
        assert la == lb, f"legal tiles differ at ply {ply}"
        assert a.current_player() == b.current_player and a.is_terminal() == b.is_terminal()
        assert a.board().tolist() == b.board.board, f"board bytes differ at ply {ply}"
        for p in range(4):
            assert sorted(a.anchors(p)) == sorted(b.board.anchors[p])
            assert a.pieces(p) == [pc.id for pc in b.board.pieces[p]]
        if ply % check_planes_every == 0:
            assert a.board_state().astype(bool).tolist() == b.get_board_state()
        t = lb[pick(ply, len(lb))]
        assert a.apply(t)
        b.apply(t)
        ply += 1
    assert a.is_terminal() == b.is_terminal()
    assert a.scores() == b.get_score() and a.payoff() == b.get_payoff() and a.last_piece_lens() == b.last_piece_lens
    assert a.history() == b.history
    return ply, b


def test_min_and_max_tile_games_agree(orc):
    n1, g1 = _drive(orc, lambda ply, n: 0)
    assert n1 == 314 and g1.get_score() == [15, -35, -4, -3]
    n2, g2 = _drive(orc, lambda ply, n: n - 1)
    assert n2 == 314 and g2.get_score() == [15, 15, 15, -42]


def test_pseudo_random_games_agree(orc):
    for seed in (1, 2, 3, 4):
        state = [seed * 2654435761 % 2**32]

        def pick(ply, n):
            state[0] = (state[0] * 1664525 + 1013904223) % 2**32
            return (state[0] >> 8) % n
        plies, _ = _drive(orc, pick)
        assert 200 <= plies <= 330


def test_place_piece_and_piece_to_finish_agree(orc):
    a, b = orc.Game(), R.Game()
    # the domino on the start corner, committed with Some(piece) although a longer placement contains it
    assert a.apply(0) and a.apply(1, piece_to_finish=1)
    b.apply(0); b.apply(1, 1)
    assert a.current_player() == b.current_player == 1 and a.legal_tiles() == b.get_legal_tiles()
    assert a.pieces(0) == [pc.id for pc in b.board.pieces[0]]
    # player 1 (anchor 19): 'Three' (index 3 of the full list) lying on 17,18,19 is valid; at 18 it overflows the row
    with pytest.raises(ValueError):
        b.place_piece(3, 0, 18)
    assert a.clone().place_piece(3, 0, 18) != 0
    nb = b.place_piece(3, 0, 17)
    assert a.place_piece(3, 0, 17) == 0
    assert a.current_player() == nb.current_player == 2 and a.history() == nb.history
    assert a.board().tolist() == nb.board.board and a.legal_tiles() == nb.get_legal_tiles()


@pytest.mark.parametrize("which", ["stub", "network"])
def test_mcts_roots_agree(orc, which):
    """mcts() without noise (exploration_fraction 0) and with greedy action choice (sample_moves 0): root children,
    visit counts and f32 value sums of every searched ply, and the actions, equal the C++ oracle's."""
    import parity
    batched, single = parity.fixed_network(3)
    sims, plies = 100, 12
    cfg = orc.make_config(sims_per_move=sims, sample_moves=0, c_base=19652.0, c_init=1.25, dirichlet_alpha=0.3,
                          exploration_fraction=0.0, seed=1)
    ref = orc.selfplay_game(cfg, 0, max_plies=plies, evaluator=single if which == "network" else None)

    def ev(state):
        planes = np.asarray(state, dtype=np.float32)
        if which == "network":
            pol, val = single(0, planes)
            return [R.f32(float(x)) for x in pol], [R.f32(float(x)) for x in val]
        return [1.0 if x else 0.0 for x in planes[4].reshape(400)], [0.25] * 4

    g = R.Game()
    for k in range(plies):
        root = R.mcts(g, sims, R.f32(19652.0), R.f32(1.25), ev, R.exp_f32_default)
        tiles = sorted(root.children)
        assert tiles == ref["roots"][k]["tile"].tolist(), f"children differ at ply {k}"
        assert [root.children[t].visits for t in tiles] == ref["roots"][k]["visits"].tolist(), f"visits differ at ply {k}"
        assert [np.float32(root.children[t].value_sum) for t in tiles] == ref["roots"][k]["value_sum"].tolist()
        assert [np.float32(root.children[t].prior) for t in tiles] == ref["roots"][k]["prior"].tolist()
        action = R.best_action_by_visits(root)
        assert action == int(ref["tiles"][k])
        g.apply(action)


def test_second_restatement_passes_the_reference_unit_tests():
    """blokus/src/pieces.rs:225-301 and board.rs:213-225, against oracle/py_restatement.py as well."""
    T, F = True, False
    assert R.Piece(0).points == 1 and len(R.Piece(0).variants) == len(R.gen_variants([[T]]))
    assert R.Piece(1).points == 2 and len(R.Piece(1).variants) == len(R.gen_variants([[T, T]]))
    assert R.Piece(2).points == 3 and len(R.Piece(2).variants) == 4
    assert R.Piece(19).points == 5 and len(R.Piece(19).variants) == 8
    v = R.PieceVariant([[T]])
    assert v.variant == [True] and v.offsets == [0] and v.width == 1
    v = R.PieceVariant([[T], [T]])
    assert len(v.variant) == 21 and v.offsets == [0, 20] and v.width == 1
    assert R.rotate([[T, T]]) == [[T], [T]] and R.rotate([[T, T], [T, F]]) == [[T, T], [F, T]]
    assert R.flip([[T, T]]) == [[T, T]] and R.flip([[T, T], [T, F]]) == [[T, T], [F, T]]
    assert [len(R.gen_variants(s)) for s in ([[T, T]], [[T, T], [T, F]], [[T, T, T], [T, F, F]])] == [2, 4, 8]
    b = R.Board()
    assert len(b.board) == 400
    assert b.is_valid_move(0, R.PieceVariant([[T, T]]), 0) is True and b.is_valid_move(0, R.PieceVariant([[T, T]]), 19) is False
