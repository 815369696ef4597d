"""Parity tests proper: the sm_100a build, through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


def test_gpu_stepwise_full_state(cuda_lib, orc):
    plies = parity.check_stepwise(cuda_lib, orc, n_games=16, seed=21, full_every=1)
    assert plies > 240


def test_gpu_playout_traces_config1(cuda_lib, orc):
    """BASELINE.json config 1: single game, seeded random legal moves, bit-exact trace."""
    for seed in (0, 1, 2):
        parity.check_playout(cuda_lib, orc, n_games=1, seed=seed)


def test_gpu_seed_free_traces(cuda_lib, orc):
    r = parity.check_playout(cuda_lib, orc, n_games=1, seed=0, flags=parity.PLAYOUT_MIN_TILE)
    assert int(r["steps"][0]) == 314
    r = parity.check_playout(cuda_lib, orc, n_games=1, seed=0, flags=parity.PLAYOUT_MAX_TILE)
    assert int(r["steps"][0]) == 314


def test_gpu_playout_4096_config2(cuda_lib, orc):
    """BASELINE.json config 2 at full size: 4096 lockstep games; 256 of them checked against the oracle
    ply by ply (trace hash), all of them through size-independent invariants."""
    from blokus_self_play import GameBatch
    res = parity.check_playout(cuda_lib, orc, n_games=4096, seed=777, n_check=256)
    assert res["steps"].min() >= 200 and res["steps"].max() <= 356
    b = GameBatch(4096, lib=cuda_lib)
    r2 = b.playout(seed=777, flags=parity.PLAYOUT_HASH)
    assert np.array_equal(r2["hash"], res["hash"])          # deterministic
    board = b.board()
    owners = board & 0x0F
    sc = b.scores()
    pcs = b.pieces()
    ll = b.last_piece_lens()
    for p in range(4):
        tiles = (owners == p + 1).sum(axis=1)
        bonus = np.where(pcs[:, p] == 0, 15 + 5 * (ll[:, p] == 1), 0)
        assert np.array_equal(sc[:, p], tiles - 89 + bonus)
    hist = b.history()
    assert [len(h) for h in hist] == r2["steps"].tolist()
    assert not b.legal_mask().any()                            # terminal: no legal tile anywhere
    pay = b.payoff()
    assert np.allclose(pay.sum(axis=1), 1.0)


def test_gpu_sharding_invariance(cuda_lib, orc):
    """Games are keyed by GLOBAL id: a shard starting at id 2048 equals the tail of the full batch."""
    from blokus_self_play import GameBatch
    a = GameBatch(512, lib=cuda_lib)
    ra = a.playout(seed=5, first_game_id=0, flags=parity.PLAYOUT_HASH)
    b = GameBatch(256, lib=cuda_lib)
    rb = b.playout(seed=5, first_game_id=256, flags=parity.PLAYOUT_HASH)
    assert np.array_equal(ra["hash"][256:], rb["hash"])


def test_gpu_illegal_move(cuda_lib, orc):
    parity.check_illegal_move(cuda_lib, orc)


def test_gpu_place_piece(cuda_lib, orc):
    assert parity.check_place_piece(cuda_lib, orc, seed=8, n_turns=70) >= 50


def test_gpu_piece_to_finish(cuda_lib, orc):
    for seed in range(4):
        parity.check_piece_to_finish(cuda_lib, orc, seed=seed, n_steps=120)


def test_gpu_clone_is_independent(cuda_lib, orc):
    from blokus_self_play import GameBatch
    a = GameBatch(8, lib=cuda_lib)
    a.playout(seed=1, max_plies=50)
    c = a.clone()
    assert np.array_equal(a.digest(), c.digest())
    c.playout(seed=2, max_plies=10)
    assert not np.array_equal(a.digest(), c.digest())
    a2 = a.clone()
    assert np.array_equal(a.digest(), a2.digest()) and a.history() == a2.history()


def test_gpu_max_plies_prefix(cuda_lib, orc):
    """Stopping after k plies and resuming gives the same games as one uninterrupted run."""
    from blokus_self_play import GameBatch
    a = GameBatch(64, lib=cuda_lib)
    a.playout(seed=9, max_plies=100)
    a.playout(seed=9)
    b = GameBatch(64, lib=cuda_lib)
    b.playout(seed=9)
    assert np.array_equal(a.digest(), b.digest()) and a.history() == b.history()


def test_gpu_playout_cut_and_resume(cuda_lib, orc):
    """The turn-structured playout keeps the turn in progress in registers: stopped after 1-7 plies (in the middle of turns,
    in both narrowing forms) it must leave exactly the state Game::apply would have left — full state against the oracle
    after every cut, some plies through Game::apply from the cut state, and the oracle's uninterrupted traces at the end."""
    parity.check_playout_cuts(cuda_lib, orc, n_games=24, seed=21, cuts=[1, 1, 1, 2, 3, 1, 5, 7, 1, 1, 37, 1, 2, 90, 1, 3, 40, 2, 60, 1])
    parity.check_playout_cuts(cuda_lib, orc, n_games=8, seed=22, cuts=[1, 2, 6, 1, 30, 1, 100, 1], apply_after=[1, 0, 2, 1, 3, 1, 2, 1])


def test_gpu_playout_new_game_flag(cuda_lib, orc):
    parity.check_playout_new_game(cuda_lib, n_games=512, seed=5, first_game_id=40)


def test_gpu_playout_65536_games(cuda_lib, orc):
    """16x BASELINE.json config 2's width in one launch: every game ends, scores stay in the rules' range, plies in the
    survey's band, and trace hashes of games spread over the batch equal the oracle's; a resume from a mid-turn cut
    (the playout loop's window-form legal set has to be materialised at the cut) gives the same games."""
    from blokus_self_play import GameBatch, PLAYOUT_HASH
    n, seed = 65536, 4242
    b = GameBatch(n, lib=cuda_lib)
    r = b.playout(seed=seed, flags=PLAYOUT_HASH)
    assert bool(b.is_terminal().all())
    sc = b.scores()
    assert sc.min() >= -89 and sc.max() <= 20
    assert 180 <= int(r["steps"].min()) and int(r["steps"].max()) <= 340 and abs(float(r["steps"].mean()) - 272.9) < 3
    extremes = (int(np.argmin(r["steps"])), int(np.argmax(r["steps"])))      # the shortest and the longest game of the batch
    for g in (0, 1, 4095, 4096, 32768, 65535) + extremes:
        ref = orc.playout(seed, g, 0)
        assert ref["n_plies"] == int(r["steps"][g]) and ref["hash"] == int(r["hash"][g]) and list(ref["scores"]) == sc[g].tolist()
    digest = b.digest()
    c = GameBatch(n, lib=cuda_lib)
    c.playout(seed=seed, max_plies=37)          # most games are mid-turn after 37 tiles
    c.playout(seed=seed, max_plies=1)
    c.playout(seed=seed)
    assert np.array_equal(c.digest(), digest)
    b.close(); c.close()
