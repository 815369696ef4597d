"""Edge cases of the C ABI: empty / skipped inputs, terminal games, capacity overflow, bad arguments —
on the CPU emulator build (logic) and, gpu-marked, on the sm_100a build."""
import ctypes as C

import numpy as np
import pytest

import parity
from blokus_self_play import BkError, Config, GameBatch, SelfPlay


def _edge_suite(lib, orc):
    # all games skipped: nothing changes, every status is 1
    b = GameBatch(3, lib=lib)
    d0 = b.digest().copy()
    assert b.apply([-1, -1, -1]).tolist() == [1, 1, 1]
    assert np.array_equal(b.digest(), d0)
    # ragged progress: one game advances, the others stay at the start
    assert b.apply([0, -1, -1]).tolist() == [0, 1, 1]
    assert [len(h) for h in b.history()] == [1, 0, 0]
    # out-of-range tiles are illegal, not a crash
    assert b.apply([400, 9999, -1], strict=False).tolist() == [-3, -3, 1]
    # a terminal game rejects every tile (is_terminal -> legal set empty, game.rs:275)
    t = GameBatch(1, lib=lib)
    t.playout(seed=3)
    assert t.is_terminal()[0] and not t.legal_mask().any()
    before = t.digest()[0]
    assert t.apply([0], strict=False)[0] == -3 and t.digest()[0] == before
    assert t.playout(seed=3)["total_steps"] == 0               # playing on from a finished game is a no-op
    # place_piece is only valid at the start of a turn
    m = GameBatch(1, lib=lib)
    m.apply([1])                                               # mid-piece now (tile 1 does not complete a piece)
    assert m.legal_tiles()[0] != [] and m.current_player()[0] == 0
    assert m.place_piece([0], [0], [0], strict=False)[0] == -3
    # null pointers / bad arguments are reported, not dereferenced
    assert lib.bk_env_apply(m._h, None, None, None) == -1 and b"null" in lib.bk_last_error()
    assert lib.bk_env_legal_mask(m._h, None) == -1
    assert lib.bk_env_anchors(m._h, 7, None) == -1
    h = C.c_void_p()
    assert lib.bk_env_create(0, 0, C.byref(h)) == -1
    for bad in (dict(sims_per_move=0), dict(c_base=0.0), dict(dirichlet_alpha=0.0)):
        kw = dict(sims_per_move=8, sample_moves=0, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=0)
        kw.update(bad)
        with pytest.raises(BkError) as e:
            SelfPlay(1, Config(**kw), lib=lib)
        assert e.value.code == -1


def _capacity_suite(lib):
    """A child pool that is too small is reported as BK_ERR_CAPACITY — never a silent overflow."""
    cfg = Config(sims_per_move=32, sample_moves=0, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=1)
    sp = SelfPlay(1, cfg, lib=lib, max_children_per_game=40)    # the root alone has 15 children, each child ~15 more
    with pytest.raises(BkError) as e:
        sp.run_stub(2)
    assert e.value.code == -4 and "pool full" in str(e.value)
    ok = SelfPlay(1, cfg, lib=lib, max_children_per_game=4096)
    ok.run_stub(2)
    assert len(ok.policy_records()[0]) == 2


def _mode_suite(lib, xp):
    """bk_selfplay_set_mode called repeatedly with a growing leaves_per_round, after tree reuse and after a packed
    gather (round-1 advisor finding: the re-allocation freed buffers it did not own), and an evaluator whose policy
    is <= 0 everywhere: reported as BK_ERR_STATE, nothing played or recorded (the reference would unwrap None)."""
    from blokus_self_play import MODE_TREE_REUSE
    cfg = Config(sims_per_move=24, sample_moves=4, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=3)
    if xp == "torch":
        import torch
        mk = lambda pl: (pl[:, 4].reshape(-1, 400).clone(), torch.full((pl.shape[0], 4), 0.25, device=pl.device))
        bad = lambda pl: (torch.zeros((pl.shape[0], 400), device=pl.device), torch.full((pl.shape[0], 4), 0.25, device=pl.device))
    else:
        mk = lambda pl: (pl[:, 4].reshape(-1, 400).copy(), np.full((pl.shape[0], 4), 0.25, dtype=np.float32))
        bad = lambda pl: (np.zeros((pl.shape[0], 400), dtype=np.float32), np.full((pl.shape[0], 4), 0.25, dtype=np.float32))
    sp = SelfPlay(5, cfg, lib=lib)
    sp.set_mode(MODE_TREE_REUSE, 1)
    sp.run_evaluator(mk, max_plies=2, buffers=parity._buffers(xp))
    assert all(len(r) == 2 for r in sp.policy_records())          # packed gather allocates its buffers
    sp.set_mode(MODE_TREE_REUSE, 2)                               # first d_pend allocation
    sp.run_evaluator(mk, max_plies=2, buffers=parity._buffers(xp))
    sp.set_mode(MODE_TREE_REUSE, 6)                               # grows d_pend; nothing else may be freed
    sp.run_evaluator(mk, max_plies=2, buffers=parity._buffers(xp))
    sp.set_mode(0, 3)                                             # smaller K: reuses the allocation
    sp.run_evaluator(mk, max_plies=2, buffers=parity._buffers(xp))
    recs = sp.policy_records()
    assert all(len(r) == 8 and all(int(v.sum()) == 24 for _, v in r) for r in recs)
    sp.close()
    for leaves in (1, 4):
        z = SelfPlay(3, cfg, lib=lib)
        z.set_mode(0, leaves)
        with pytest.raises(BkError) as e:
            z.run_evaluator(bad, max_plies=1, buffers=parity._buffers(xp))
        assert e.value.code == -5 and "no selectable child" in str(e.value)
        assert [len(h) for h in z.env.history()] == [0, 0, 0]    # nothing was played from the failed search
        z.close()


def test_emu_set_mode_regrow_and_dead_policy(emu_lib):
    _mode_suite(emu_lib, "numpy")


@pytest.mark.gpu
def test_gpu_set_mode_regrow_and_dead_policy(cuda_lib):
    _mode_suite(cuda_lib, "torch")


def test_emu_edge_cases(emu_lib, orc):
    _edge_suite(emu_lib, orc)


def test_emu_capacity_overflow_is_reported(emu_lib):
    _capacity_suite(emu_lib)


@pytest.mark.gpu
def test_gpu_edge_cases(cuda_lib, orc):
    _edge_suite(cuda_lib, orc)
    _capacity_suite(cuda_lib)


@pytest.mark.gpu
def test_gpu_largest_batch_config5_shard(cuda_lib, orc):
    """8192 games (config 5's per-GPU shard) through the env path: every game ends, invariants hold,
    and a sample is checked ply by ply against the oracle."""
    res = parity.check_playout(cuda_lib, orc, n_games=8192, seed=31, first_game_id=3 * 8192, n_check=64)
    assert res["steps"].min() >= 200


def _misaligned_records_suite(lib, buffers):
    """Games advanced OUTSIDE the search (env.apply before it) have more history than policy records: the training
    tensors and the reference tuple would be misaligned, so both refuse (round-1 advisor finding)."""
    cfg = Config(sims_per_move=8, sample_moves=2, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=2)
    sp = SelfPlay(2, cfg, lib=lib)
    sp.env.apply([0, 0])                           # one ply played by hand
    sp.run_stub(3)
    with pytest.raises(BkError) as e:
        sp.training_tensors(buffers=buffers)
    assert e.value.code == -5 and "do not align" in str(e.value)
    with pytest.raises(ValueError):
        sp.game_data()
    sp.close()
    ok = SelfPlay(2, cfg, lib=lib)
    ok.run_stub(3)
    st, po, va, offs = ok.training_tensors(buffers=buffers)
    assert len(st) == 6 and offs.tolist() == [0, 3, 6] and len(ok.game_data()[1][0]) == 3
    ok.close()


def test_emu_misaligned_records_are_refused(emu_lib):
    _misaligned_records_suite(emu_lib, parity.HostBuffers())


@pytest.mark.gpu
def test_gpu_misaligned_records_are_refused(cuda_lib):
    _misaligned_records_suite(cuda_lib, None)
