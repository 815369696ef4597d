"""Kernel-source logic checks on the CPU warp emulator (tests/warp_emu): the same parity assertions the
GPU tests make, at sizes the emulator finishes in seconds.  These prove the SOURCE is right before any
GPU minute is spent; the GPU tests (test_gpu_env.py) prove the sm_100a BUILD is."""
import parity


def test_emu_stepwise_full_state(emu_lib, orc):
    plies = parity.check_stepwise(emu_lib, orc, n_games=2, seed=11, max_plies=60, full_every=1)
    assert plies == 60


def test_emu_stepwise_to_terminal(emu_lib, orc):
    plies = parity.check_stepwise(emu_lib, orc, n_games=3, seed=5, full_every=25)
    assert plies > 240


def test_emu_playout_traces(emu_lib, orc):
    parity.check_playout(emu_lib, orc, n_games=4, seed=99, first_game_id=1000)


def test_emu_playout_cut_and_resume(emu_lib, orc):
    """Cuts of 1-7 plies land in the middle of turns in both narrowing forms (more / fewer than 32 surviving placements)."""
    parity.check_playout_cuts(emu_lib, orc, n_games=3, seed=21, cuts=[1, 1, 1, 2, 3, 1, 5, 7, 1, 1, 37, 1, 2, 90, 1, 3])
    parity.check_playout_cuts(emu_lib, orc, n_games=2, seed=22, cuts=[1, 2, 6, 1, 30, 1], apply_after=[1, 0, 2, 1, 3, 1])


def test_emu_playout_new_game_flag(emu_lib, orc):
    parity.check_playout_new_game(emu_lib, n_games=3, seed=5, first_game_id=40)


def test_emu_seed_free_traces(emu_lib, orc):
    r = parity.check_playout(emu_lib, orc, n_games=1, seed=0, flags=parity.PLAYOUT_MIN_TILE)
    assert int(r["steps"][0]) == 314
    r = parity.check_playout(emu_lib, orc, n_games=1, seed=0, flags=parity.PLAYOUT_MAX_TILE)
    assert int(r["steps"][0]) == 314


def test_emu_illegal_move(emu_lib, orc):
    parity.check_illegal_move(emu_lib, orc)


def test_emu_place_piece(emu_lib, orc):
    assert parity.check_place_piece(emu_lib, orc, seed=3, n_turns=12) == 12


def test_emu_piece_to_finish(emu_lib, orc):
    parity.check_piece_to_finish(emu_lib, orc, seed=4, n_steps=40)


def test_emu_game_mirror_get_piece(emu_lib, orc):
    """Game::get_piece / get_current_player_pieces (game.rs:230-236) of the Python mirror against the oracle's tables."""
    from blokus_self_play import Game
    g = Game.reset(lib=emu_lib)
    assert g.get_current_player_pieces() == list(range(21))
    pv = g.get_piece(0, 1, 1)
    ref = orc.piece_variant(1, 1)
    assert (pv["offsets"], pv["width"], pv["len"], pv["piece_id"]) == (ref["offsets"], ref["width"], ref["len"], 1)
    g2 = g.place_piece(0, 0, 0)
    assert g2.get_piece(0, 0, 0)["piece_id"] == 1 and g.get_piece(0, 0, 0)["piece_id"] == 0
