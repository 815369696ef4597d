"""MCTS parity tests proper: the sm_100a build through the C ABI against the CPU oracle.
North star: visit counts exact under fixed priors and seeded noise, Q within 1e-5 relative."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

CONFIG3 = dict(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
               exploration_fraction=0.25, seed=20261018)


def test_gpu_selfplay_full_games_64_sims(cuda_lib, orc):
    """Whole games to the end (every ply's root visit vector + action trace + payoff)."""
    r = parity.check_selfplay_stub(cuda_lib, orc, 6, dict(CONFIG3, sims_per_move=64, seed=11), first_game_id=100)
    assert r["plies"] > 6 * 230


def test_gpu_selfplay_config3_prefix(cuda_lib, orc):
    """BASELINE.json config 3 parameters (800 sims, alpha 0.03, frac 0.25), first plies of a few games."""
    parity.check_selfplay_stub(cuda_lib, orc, 64, CONFIG3, first_game_id=0, max_plies=5, n_check=4)


def test_gpu_selfplay_shipped_config(cuda_lib, orc):
    """model/training.py:267-272: 50 sims, alpha 0.3, sample_moves 30."""
    parity.check_selfplay_stub(cuda_lib, orc, 4, dict(sims_per_move=50, sample_moves=30, c_base=19652, c_init=1.25,
                                                     dirichlet_alpha=0.3, exploration_fraction=0.25, seed=5))


def test_gpu_selfplay_resume_equals_one_run(cuda_lib, orc):
    """max_plies prefixes compose: 3 + 4 plies == 7 plies."""
    from blokus_self_play import SelfPlay, Config
    cfg = Config(**dict(CONFIG3, sims_per_move=100))
    a = SelfPlay(16, cfg, lib=cuda_lib)
    a.run_stub(3)
    a.run_stub(4)
    b = SelfPlay(16, cfg, lib=cuda_lib)
    b.run_stub(7)
    assert a.env.history() == b.env.history()
    ra, rb = a.policy_records(), b.policy_records()
    for g in range(16):
        assert len(ra[g]) == len(rb[g]) == 7
        for (t1, v1), (t2, v2) in zip(ra[g], rb[g]):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)


def test_gpu_selfplay_1024_games_invariants(cuda_lib, orc):
    """Config 3 at full width (1024 games, 800 sims) for two plies: tree invariants for every game and
    sharding invariance (a 256-game shard at id 512 equals that slice of the full batch)."""
    from blokus_self_play import SelfPlay, Config
    cfg = Config(**CONFIG3)
    sp = SelfPlay(1024, cfg, lib=cuda_lib)
    sp.run_stub(2)
    recs = sp.policy_records()
    for g in range(1024):
        assert len(recs[g]) == 2
        for tiles, visits in recs[g]:
            assert int(visits.sum()) == 800 and np.all(np.diff(tiles) > 0)
    c = sp.counters()
    assert c["sims"] == 1024 * 2 * 800 and c["nodes"] <= c["sims"] + 2 * 1024
    shard = SelfPlay(256, cfg, first_game_id=512, lib=cuda_lib)
    shard.run_stub(2)
    rs = shard.policy_records()
    for g in range(256):
        for (t1, v1), (t2, v2) in zip(rs[g], recs[512 + g]):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)
    assert shard.env.history() == sp.env.history()[512:768]


def test_gpu_external_evaluator_protocol(cuda_lib, orc):
    """Non-uniform fixed network through begin_ply / leaf_planes / expand_backup / end_ply (device batches)."""
    info = parity.check_selfplay_evaluator(cuda_lib, orc, 8, dict(CONFIG3, sims_per_move=96, dirichlet_alpha=0.3, seed=2),
                                           first_game_id=40, max_plies=12, xp="torch")
    assert info["plies"] == 12


def test_gpu_play_training_game_reference_signature(cuda_lib, orc):
    """play_training_game(id, config, inference_queue, pipe) — the reference's entry point and IPC protocol
    (self_play/src/lib.rs:9-32, simulation.rs:50-57), a whole game, tuple shape of simulation.rs:293-295."""
    n = parity.check_play_training_game(cuda_lib, orc, dict(sims_per_move=12, sample_moves=30, c_base=19652, c_init=1.25,
                                                            dirichlet_alpha=0.3, exploration_fraction=0.25, seed=6))
    assert n > 200


def test_gpu_external_stub_equals_fused_kernel(cuda_lib, orc):
    import torch
    from blokus_self_play import SelfPlay, Config
    kw = dict(CONFIG3, sims_per_move=200)
    a = SelfPlay(32, Config(**kw), first_game_id=9, lib=cuda_lib)
    a.run_stub(6)
    b = SelfPlay(32, Config(**kw), first_game_id=9, lib=cuda_lib)
    b.run_evaluator(lambda pl: (pl[:, 4].reshape(-1, 400), torch.full((pl.shape[0], 4), 0.25, device=pl.device)), 6)
    assert a.env.history() == b.env.history()
    for ra, rb in zip(a.policy_records(), b.policy_records()):
        for (t1, v1), (t2, v2) in zip(ra, rb):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)


def test_gpu_training_tensors(cuda_lib, orc):
    """model/training.py `save()` on the device, whole games, bit-exact."""
    n = parity.check_training_tensors(cuda_lib, orc, 6, dict(CONFIG3, sims_per_move=24, seed=8), max_plies=-1, xp="torch")
    assert n > 6 * 230


def test_gpu_arena_play_test_game(cuda_lib, orc):
    """play_test_game / arena on the B200 game engine."""
    parity.check_arena(cuda_lib, orc, n_games=6, seed=5)


def test_gpu_throughput_modes(cuda_lib, orc):
    """Row f3 on the device: multi-leaf path with one leaf == exact mode bit for bit; forced-ply shortcut keeps the
    training tuple; 8 leaves per round keep the visit-sum invariant and cut the evaluator rounds."""
    r = parity.check_throughput_modes(cuda_lib, 8, dict(CONFIG3, sims_per_move=96, sample_moves=6, dirichlet_alpha=0.3, seed=8),
                                      max_plies=14, xp="torch", leaves=8)
    assert r["rounds_multi"] * 3 < r["rounds_exact"]


def test_gpu_skip_forced_stub_full_games(cuda_lib, orc):
    full, skipped = parity.check_skip_forced_stub(cuda_lib, 8, dict(CONFIG3, sims_per_move=48, seed=4), max_plies=-1)
    assert skipped < full


def test_gpu_config5_shard_width(cuda_lib, orc):
    """BASELINE.json config 5's per-GPU share at full width: 8192 games x 800 sims (the 28-games-per-SM instantiation of
    the stub kernel), two plies — visit-sum invariant for every game, and a 48-game batch at global id 3000 (the
    all-registers instantiation) reproduces that slice: results depend on the global game id only."""
    from blokus_self_play import SelfPlay, Config
    cfg = Config(**CONFIG3)
    big = SelfPlay(8192, cfg, first_game_id=0, lib=cuda_lib)
    big.run_stub(2)
    recs = big.policy_records()
    assert all(len(r) == 2 and all(int(v.sum()) == 800 for _, v in r) for r in recs)
    hist = big.env.history()
    small = SelfPlay(48, cfg, first_game_id=3000, lib=cuda_lib)
    small.run_stub(2)
    assert small.env.history() == hist[3000:3048]
    for ra, rb in zip(small.policy_records(), recs[3000:3048]):
        for (t1, v1), (t2, v2) in zip(ra, rb):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)
    c = big.counters()
    assert c["sims"] == 8192 * 2 * 800
    big.close(); small.close()


def test_gpu_tree_reuse(cuda_lib, orc):
    """Row f3, tree reuse on the device (256 sims, 24 plies, 6 games): invariants, determinism across launch
    splits, fused kernel == evaluator protocol, fewer simulations than the exact mode."""
    full, reuse = parity.check_tree_reuse(cuda_lib, 6, dict(CONFIG3, sims_per_move=256, sample_moves=6, dirichlet_alpha=0.3, seed=21),
                                          max_plies=24, xp="torch")
    assert reuse < full


def test_gpu_selfplay_config3_deep_prefix(cuda_lib, orc):
    """BASELINE.json config 3 parameters (800 sims, alpha 0.03) carried 56 plies into a game — past the opening, where
    roots have many children (the > 32-children select path), trees are deep and the narrowing cache is in every form —
    every ply's root visit vector, the action trace and the last root's value sums against the oracle."""
    parity.check_selfplay_stub(cuda_lib, orc, 8, CONFIG3, first_game_id=40, max_plies=56, n_check=1)


def test_gpu_throughput_modes_at_config3_width(cuda_lib, orc):
    """All three opt-in modes together at config 3's search size (800 sims) on 128 games for 12 plies through the
    evaluator protocol (4 leaves per round, tree reuse, forced-ply shortcut), the evaluator being a device-side stub:
    no pool overflows, every recorded policy sums to 800 visits, games advance; and complete games of the fused
    kernel with tree reuse + shortcut finish with valid scores."""
    import torch
    from blokus_self_play import SelfPlay, Config, MODE_SKIP_FORCED, MODE_TREE_REUSE
    cfg = Config(**CONFIG3)
    sp = SelfPlay(128, cfg, lib=cuda_lib)
    sp.set_mode(MODE_SKIP_FORCED | MODE_TREE_REUSE, 4)

    def ev(planes):                                   # policy = legal mask, value = [0.4, 0.3, 0.2, 0.1]
        n = planes.shape[0]
        return planes[:, 4].reshape(n, 400).clone(), torch.tensor([0.4, 0.3, 0.2, 0.1], device=planes.device).repeat(n, 1)
    info = sp.run_evaluator(ev, max_plies=12)
    assert info["plies"] == 12
    hist = sp.env.history()
    for g, recs in enumerate(sp.policy_records()):
        assert len(recs) == 12
        for k, (tiles, visits) in enumerate(recs):
            assert int(visits.sum()) == 800 and np.all(np.diff(tiles) > 0) and hist[g][k][1] in tiles.tolist()
    sp.close()
    full = SelfPlay(64, cfg, first_game_id=900, lib=cuda_lib)
    full.set_mode(MODE_SKIP_FORCED | MODE_TREE_REUSE, 1)
    full.run_stub(-1)
    assert bool(full.env.is_terminal().all())
    sc = full.env.scores()
    assert sc.min() >= -89 and sc.max() <= 20
    for recs in full.policy_records():
        assert all(int(v.sum()) == 800 for _, v in recs)
    full.close()


def test_gpu_results_packed_equal_per_game_abi(cuda_lib, orc):
    """bk_selfplay_results (per-game copies) and bk_selfplay_results_packed (one CSR gather) carry the same records."""
    from blokus_self_play import SelfPlay, Config
    sp = SelfPlay(40, Config(**dict(CONFIG3, sims_per_move=64)), first_game_id=9, lib=cuda_lib)
    sp.run_stub(9)
    a, b = sp.policy_records_unpacked(), sp.policy_records()
    for ga, gb in zip(a, b):
        assert len(ga) == len(gb) == 9
        for (t1, v1), (t2, v2) in zip(ga, gb):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)
    sp.close()


def test_gpu_dense_rows_many_games(cuda_lib, orc):
    """Multi-leaf mode with more games than one chunk of the slot scan (600 games, 4 leaves per round): with the root noise
    off and greedy actions every game is the same game, so a misplaced evaluator row would break the equality."""
    import torch
    from blokus_self_play import SelfPlay, Config
    cfg = Config(**dict(CONFIG3, sims_per_move=32, sample_moves=0, exploration_fraction=0.0))
    ramp = ((torch.arange(400, device="cuda") % 7 + 1.0) / 7.0).float()

    def ev(planes):
        n = planes.shape[0]
        own = planes[:, 0].reshape(n, 400).sum(dim=1, keepdim=True)
        return planes[:, 4].reshape(n, 400) * ramp * (1.0 + 0.1 * own), torch.tensor([0.4, 0.3, 0.2, 0.1], device=planes.device).repeat(n, 1)
    sp = SelfPlay(600, cfg, lib=cuda_lib)
    sp.set_mode(0, 4)
    sp.run_evaluator(ev, max_plies=5)
    hist = sp.env.history()
    assert all(h == hist[0] for h in hist)
    recs = sp.policy_records()
    for g in (1, 255, 256, 257, 511, 512, 599):
        for (t1, v1), (t2, v2) in zip(recs[0], recs[g]):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)
    sp.close()


def test_gpu_pipeline_equals_one_warp_kernel(cuda_lib):
    """Two-warp pipelined stub kernel == one-warp kernel, complete games at 200 sims (terminal leaves inside the trees)."""
    from test_emu_mcts import _pipeline_equals_one_warp_kernel
    h = _pipeline_equals_one_warp_kernel(cuda_lib, 24, 200, -1)
    assert all(len(x) > 200 for x in h)


def test_gpu_pipeline_tables_beyond_shared_memory(cuda_lib):
    """1100 simulations per move: the UCB factor tables no longer fit the pipelined kernel's shared-memory staging
    (1024 entries) and are read from global memory; 5000: the exhaustive check of the short division is skipped
    (sims > 4096) and the kernels use the IEEE division.  Both still equal the one-warp kernel bit for bit."""
    from test_emu_mcts import _pipeline_equals_one_warp_kernel
    _pipeline_equals_one_warp_kernel(cuda_lib, 8, 1100, 3)
    _pipeline_equals_one_warp_kernel(cuda_lib, 4, 5000, 1)


def test_gpu_one_warp_instantiations_agree(cuda_lib):
    """Every residency instantiation of the one-warp search kernel (1 / 12 / 16 / 20 / 28 games per SM, chosen by batch
    size in production, forced here through BK_STUB_MIN_BLOCKS) plays the same games: exact mode against the two-warp
    pipeline, and the opt-in modes (forced-ply shortcut + tree reuse) against each other — histories, policy records
    and payoffs, 40 plies at 160 sims."""
    import os
    from blokus_self_play import SelfPlay, Config, MODE_SKIP_FORCED, MODE_TREE_REUSE
    cfg = Config(**dict(CONFIG3, sims_per_move=160, dirichlet_alpha=0.3, seed=77))

    def run(min_blocks, flags):
        old = os.environ.get("BK_STUB_MIN_BLOCKS")
        if min_blocks:
            os.environ["BK_STUB_MIN_BLOCKS"] = str(min_blocks)       # read by bk_selfplay_create
        try:
            sp = SelfPlay(24, cfg, first_game_id=500, lib=cuda_lib)
        finally:
            if min_blocks:
                if old is None:
                    del os.environ["BK_STUB_MIN_BLOCKS"]
                else:
                    os.environ["BK_STUB_MIN_BLOCKS"] = old
        if flags:
            sp.set_mode(flags, 1)
        sp.run_stub(40)
        out = (sp.env.history(), sp.policy_records(), sp.env.payoff().tolist())
        sp.close()
        return out

    def same(a, b):
        assert a[0] == b[0] and a[2] == b[2]
        for ra, rb in zip(a[1], b[1]):
            assert len(ra) == len(rb)
            for (t1, v1), (t2, v2) in zip(ra, rb):
                assert np.array_equal(t1, t2) and np.array_equal(v1, v2)

    for flags in (0, MODE_SKIP_FORCED | MODE_TREE_REUSE):
        base = run(0, flags)                 # the library's own choice: the pipeline in exact mode, <1, true> with modes
        assert all(len(h) >= 40 for h in base[0])
        for mb in (1, 12, 16, 20, 28):
            same(base, run(mb, flags))
