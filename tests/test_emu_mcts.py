"""MCTS kernel-source logic checks on the CPU warp emulator (small trees; the GPU tests go to 800 sims)."""
import numpy as np

import parity


def test_emu_selfplay_stub_alpha003(emu_lib, orc):
    parity.check_selfplay_stub(emu_lib, orc, 2, dict(sims_per_move=24, sample_moves=4, c_base=19652, c_init=1.25,
                                                    dirichlet_alpha=0.03, exploration_fraction=0.25, seed=77),
                               first_game_id=5, max_plies=8)


def test_emu_selfplay_stub_shipped_config(emu_lib, orc):
    """model/training.py:267-272 values (alpha 0.3), greedy action from the first ply (sample_moves 0)."""
    parity.check_selfplay_stub(emu_lib, orc, 1, dict(sims_per_move=16, sample_moves=0, c_base=19652, c_init=1.25,
                                                    dirichlet_alpha=0.3, exploration_fraction=0.25, seed=3),
                               first_game_id=0, max_plies=12)


def test_emu_external_evaluator_protocol(emu_lib, orc):
    info = parity.check_selfplay_evaluator(emu_lib, orc, 2, dict(sims_per_move=20, sample_moves=3, c_base=19652, c_init=1.25,
                                                                 dirichlet_alpha=0.3, exploration_fraction=0.25, seed=9),
                                           first_game_id=3, max_plies=5, xp="numpy")
    assert info["plies"] == 5 and info["rounds"] >= 5 * 20


def test_emu_external_stub_equals_fused_kernel(emu_lib, orc):
    """The stub evaluated on the HOST through the protocol (policy 1.0 on legal tiles, value 0.25) must give
    the fused device kernel's result bit for bit."""
    import numpy as np
    from blokus_self_play import SelfPlay, Config
    kw = dict(sims_per_move=16, sample_moves=2, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=4)
    a = SelfPlay(2, Config(**kw), first_game_id=1, lib=emu_lib)
    a.run_stub(4)
    b = SelfPlay(2, Config(**kw), first_game_id=1, lib=emu_lib)
    b.run_evaluator(lambda pl: (pl[:, 4].reshape(-1, 400).copy(), np.full((pl.shape[0], 4), 0.25, dtype=np.float32)), 4, buffers=parity.HostBuffers())
    assert a.env.history() == b.env.history()
    for ra, rb in zip(a.policy_records(), b.policy_records()):
        for (t1, v1), (t2, v2) in zip(ra, rb):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)
    for x, y in zip(a.last_root(), b.last_root()):
        assert np.array_equal(x["prior"], y["prior"]) and np.array_equal(x["value_sum"], y["value_sum"])


def test_emu_training_tensors(emu_lib, orc):
    n = parity.check_training_tensors(emu_lib, orc, 1, dict(sims_per_move=6, sample_moves=2, c_base=19652, c_init=1.25,
                                                            dirichlet_alpha=0.3, exploration_fraction=0.25, seed=1),
                                      max_plies=9, xp="numpy")
    assert n == 9


def test_emu_arena_play_test_game(emu_lib, orc):
    parity.check_arena(emu_lib, orc, n_games=1, seed=2)


def test_emu_throughput_modes(emu_lib, orc):
    """Row f3: multi-leaf (virtual loss) rounds and the forced-ply shortcut, against the exact mode."""
    r = parity.check_throughput_modes(emu_lib, 2, dict(sims_per_move=12, sample_moves=2, c_base=19652, c_init=1.25,
                                                       dirichlet_alpha=0.3, exploration_fraction=0.25, seed=6),
                                      max_plies=7, xp="numpy", leaves=3)
    assert r["rounds_multi"] < r["rounds_exact"]


def test_emu_skip_forced_stub(emu_lib, orc):
    full, skipped = parity.check_skip_forced_stub(emu_lib, 1, dict(sims_per_move=8, sample_moves=2, c_base=19652, c_init=1.25,
                                                                   dirichlet_alpha=0.3, exploration_fraction=0.25, seed=2),
                                                  max_plies=12)
    assert skipped < full


def test_emu_tree_reuse(emu_lib, orc):
    full, reuse = parity.check_tree_reuse(emu_lib, 2, dict(sims_per_move=20, sample_moves=2, c_base=19652, c_init=1.25,
                                                          dirichlet_alpha=0.3, exploration_fraction=0.25, seed=12),
                                          max_plies=8, xp="numpy")
    assert reuse < full


def test_emu_packed_results_equal_per_game_records(emu_lib, orc):
    """bk_selfplay_results_packed (one CSR gather) carries exactly the per-game policy records."""
    import numpy as np
    from blokus_self_play import SelfPlay, Config
    sp = SelfPlay(3, Config(sims_per_move=10, sample_moves=2, c_base=19652, c_init=1.25, dirichlet_alpha=0.3,
                            exploration_fraction=0.25, seed=8), first_game_id=2, lib=emu_lib)
    sp.run_stub(5)
    recs = sp.policy_records_unpacked()
    ply_off, ply_ptr, tiles, visits = sp.policy_records_packed()
    assert ply_off.tolist() == [0, 5, 10, 15] and len(ply_ptr) == 16 and ply_ptr[-1] == len(tiles) == len(visits)
    for g in range(3):
        for k, (t, v) in enumerate(recs[g]):
            a, b = ply_ptr[ply_off[g] + k], ply_ptr[ply_off[g] + k + 1]
            assert np.array_equal(tiles[a:b], t) and np.array_equal(visits[a:b], v)
    sp.close()


def test_emu_dense_rows_across_scan_chunks(emu_lib, orc):
    """The multi-leaf mode's slot scan (one CTA, 256 games per chunk) on more games than one chunk: dense rows must
    line up with the games (a wrong carry would hand game g another game's policy).  All games start from the same
    position and the root noise is off, so with a position-dependent policy every game must end the ply with the SAME
    root priors, children and visit counts."""
    import numpy as np
    from blokus_self_play import SelfPlay, Config
    n = 262
    cfg = Config(sims_per_move=2, sample_moves=0, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.0, seed=5)
    calls = []

    def ev(planes):
        pl = np.asarray(planes)
        calls.append(pl.shape[0])
        ramp = (np.arange(400, dtype=np.float32) % 7 + 1.0) / 7.0
        own = pl[:, 0].reshape(-1, 400).sum(axis=1, keepdims=True)          # depends on the position, not only on the legal set
        return (pl[:, 4].reshape(-1, 400) * ramp * (1.0 + 0.1 * own)).astype(np.float32), \
            np.tile(np.array([0.4, 0.3, 0.2, 0.1], np.float32), (pl.shape[0], 1))
    a = SelfPlay(n, cfg, lib=emu_lib)
    a.set_mode(0, 2)
    a.run_evaluator(ev, max_plies=1, buffers=parity.HostBuffers())
    assert calls[0] == n and max(calls) <= 2 * n            # the root round is one row per game; later rounds are dense
    roots = a.last_root()
    for g in (1, 255, 256, 257, n - 1):
        for key in ("tile", "visits", "prior", "value_sum"):
            assert np.array_equal(roots[0][key], roots[g][key]), (g, key)
    assert int(roots[0]["visits"].sum()) == 2
    a.close()


def _pipeline_equals_one_warp_kernel(lib, n_games, sims, max_plies):
    """The two-warp pipelined stub kernel (bk_mcts_pipe.cuh) against the one-warp kernel: histories, every policy record,
    the last root's value sums and priors, and the counters — complete games, so terminal leaves (remembered in the
    entry, speculation rolled back on first discovery) are covered."""
    import os
    from blokus_self_play import SelfPlay, Config
    cfg = Config(sims_per_move=sims, sample_moves=8, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=12)
    out = []
    for pipe in ("0", "1"):
        os.environ["BK_STUB_PIPE"] = pipe
        try:
            sp = SelfPlay(n_games, cfg, first_game_id=70, lib=lib)
        finally:
            del os.environ["BK_STUB_PIPE"]
        sp.run_stub(max_plies)
        out.append((sp.env.history(), sp.policy_records(), sp.last_root(), sp.counters(), sp.env.payoff().tolist()))
        sp.close()
    (h0, r0, l0, c0, p0), (h1, r1, l1, c1, p1) = out
    assert h0 == h1 and p0 == p1
    for ga, gb in zip(r0, r1):
        assert len(ga) == len(gb)
        for (t1, v1), (t2, v2) in zip(ga, gb):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)
    for a, b in zip(l0, l1):
        assert np.array_equal(a["value_sum"], b["value_sum"]) and np.array_equal(a["prior"], b["prior"])
    assert c0["sims"] == c1["sims"] and c0["nodes"] == c1["nodes"] and c0["entries"] == c1["entries"]
    assert c1["applies"] <= c0["applies"]              # revisited terminal leaves are not re-applied by the pipeline
    return h1


def test_emu_pipeline_equals_one_warp_kernel(emu_lib):
    h = _pipeline_equals_one_warp_kernel(emu_lib, 1, 8, -1)
    assert all(len(x) > 200 for x in h)
