"""Parity checks shared by the CPU-emulator tests and the GPU tests: the same assertions run against
whichever build of the kernel sources `lib` is."""
import numpy as np

from blokus_self_play import GameBatch, PLAYOUT_HASH, PLAYOUT_MIN_TILE, PLAYOUT_MAX_TILE, BkError


class HostBuffers:
    """Evaluator batches in HOST memory: only meaningful with the tests' CPU warp-emulator build of the kernel sources
    (tests/warp_emu), where "device" pointers are host pointers.  The product's default is TorchBuffers (CUDA)."""

    @staticmethod
    def zeros(shape):
        return np.zeros(shape, dtype=np.float32)

    empty = zeros

    @staticmethod
    def f32(a):
        return np.ascontiguousarray(a, dtype=np.float32)

    @staticmethod
    def ptr(a):
        return a.ctypes.data

    @staticmethod
    def to_host(a):
        return np.asarray(a)

    @staticmethod
    def from_host(a):
        return np.ascontiguousarray(a, dtype=np.float32)


def _buffers(xp):
    return HostBuffers() if xp == "numpy" else None


def oracle_mask(tiles):
    m = np.zeros(400, dtype=np.uint8)
    m[list(tiles)] = 1
    return m


def compare_state(batch, games, what="all"):
    """Every accessor of the batch against the oracle games (bit-exact)."""
    n = batch.n
    lm = batch.legal_mask()
    lt = batch.legal_tiles()                     # the list form (bk_env_legal_tiles), ascending
    cur = batch.current_player()
    term = batch.is_terminal()
    for g in range(n):
        assert np.array_equal(lm[g], oracle_mask(games[g].legal_tiles())), f"legal tiles differ, game {g}"
        assert lt[g] == sorted(games[g].legal_tiles()), f"legal tile list differs, game {g}"
        assert term[g] == games[g].is_terminal()
        assert cur[g] == games[g].current_player(), f"current player differs, game {g}"
    if what != "all":
        return
    board = batch.board()
    planes = batch.board_state()
    act = batch.is_player_active()
    sc = batch.scores()
    pay = batch.payoff()
    pcs = batch.pieces()
    ll = batch.last_piece_lens()
    dg = batch.digest()
    anch = [batch.anchors(p) for p in range(4)]
    anch_cur = batch.anchors(-1)
    hist = batch.history()
    for g in range(n):
        og = games[g]
        assert np.array_equal(board[g], og.board()), f"board bytes differ, game {g}"
        assert np.array_equal(planes[g], og.board_state()), f"planes differ, game {g}"
        assert act[g].tolist() == [og.is_player_active(p) for p in range(4)]
        assert sc[g].tolist() == og.scores()
        assert pay[g].tolist() == og.payoff()
        assert ll[g].tolist() == og.last_piece_lens()
        for p in range(4):
            mask = 0
            for pid in og.pieces(p):
                mask |= 1 << pid
            assert int(pcs[g, p]) == mask
            assert np.array_equal(anch[p][g], oracle_mask(og.anchors(p))), f"anchors differ, game {g} player {p}"
        assert np.array_equal(anch_cur[g], oracle_mask(og.anchors(-1)))
        assert int(dg[g]) == og.digest()
        assert hist[g] == og.history()


def check_stepwise(lib, orc, n_games=4, seed=0, max_plies=10**9, full_every=1):
    """Lockstep random games through Game::apply on both sides, comparing after every ply."""
    rng = np.random.default_rng(seed)
    batch = GameBatch(n_games, lib=lib)
    games = [orc.Game() for _ in range(n_games)]
    compare_state(batch, games)
    ply = 0
    while ply < max_plies and not all(g.is_terminal() for g in games):
        tiles = []
        for g in games:
            if g.is_terminal():
                tiles.append(-1)
            else:
                lt = g.legal_tiles()
                tiles.append(int(lt[rng.integers(len(lt))]))
        st = batch.apply(tiles)
        for i, g in enumerate(games):
            if tiles[i] >= 0:
                assert st[i] == 0
                assert g.apply(tiles[i])
            else:
                assert st[i] == 1
        ply += 1
        compare_state(batch, games, "all" if ply % full_every == 0 else "light")
    compare_state(batch, games)
    batch.close()
    return ply


def check_playout_cuts(lib, orc, n_games, seed, cuts, apply_after=(), first_game_id=0):
    """The persistent playout stopped after `cuts` plies (mostly in the middle of a turn, where the turn it keeps
    in registers has to be written back as Game::apply would have left it) and picked up again: after every cut the
    FULL state equals the oracle game that replayed the same tiles; optionally some plies go through Game::apply
    from the cut state (the narrowing cache the playout stored is the one apply continues from); the finished games
    equal the oracle's uninterrupted traces."""
    batch = GameBatch(n_games, lib=lib)
    games = [orc.Game() for _ in range(n_games)]
    done = [0] * n_games
    rng = np.random.default_rng(seed)

    def sync():
        hist = batch.history()
        for g in range(n_games):
            for pl, t in hist[g][done[g]:]:
                assert games[g].current_player() == pl and games[g].apply(int(t))
            done[g] = len(hist[g])
        compare_state(batch, games)

    applied = False
    for i, k in enumerate(cuts):
        batch.playout(seed=seed, first_game_id=first_game_id, max_plies=k)
        sync()
        for _ in range(apply_after[i] if i < len(apply_after) else 0):
            tiles = []
            for g in games:
                lt = g.legal_tiles()
                tiles.append(-1 if g.is_terminal() else int(lt[rng.integers(len(lt))]))
            batch.apply(tiles)
            applied = True
            sync()
    batch.playout(seed=seed, first_game_id=first_game_id)
    sync()
    assert batch.is_terminal().all()
    if not applied:
        hist = batch.history()
        sc = batch.scores()
        for g in range(n_games):
            ref = orc.playout(seed, first_game_id + g, 0)
            assert [t for _, t in hist[g]] == ref["tiles"].tolist() and list(ref["scores"]) == sc[g].tolist()
    batch.close()


def check_playout_new_game(lib, n_games, seed, first_game_id=0):
    """PLAYOUT_NEW_GAME (Game::reset inside the playout launch) equals reset() followed by playout(), whatever the
    handle held before: states, histories (cleared past the end), steps and trace hashes."""
    from blokus_self_play import PLAYOUT_NEW_GAME
    a = GameBatch(n_games, lib=lib)
    b = GameBatch(n_games, lib=lib)
    b.playout(seed=seed + 1, first_game_id=7)                       # leftovers of another batch in b's buffers
    for flags in (0, PLAYOUT_HASH):
        a.reset()
        ra = a.playout(seed=seed, first_game_id=first_game_id, flags=flags)
        rb = b.playout(seed=seed, first_game_id=first_game_id, flags=flags | PLAYOUT_NEW_GAME)
        assert np.array_equal(ra["steps"], rb["steps"]) and np.array_equal(ra["hash"], rb["hash"])
        assert np.array_equal(a.digest(), b.digest()) and a.history() == b.history()
        assert np.array_equal(a.scores(), b.scores()) and np.array_equal(a.legal_mask(), b.legal_mask())
    rb = b.playout(seed=seed, first_game_id=first_game_id, max_plies=9, flags=PLAYOUT_NEW_GAME)   # a cut right after the reset
    a.reset()
    a.playout(seed=seed, first_game_id=first_game_id, max_plies=9)
    assert np.array_equal(a.digest(), b.digest()) and a.history() == b.history() and int(rb["steps"].max()) == 9
    a.close(); b.close()


def check_playout(lib, orc, n_games, seed, first_game_id=0, n_check=None, flags=0):
    """Device-resident playout to the end vs the oracle's trace (hash of every ply's full state)."""
    batch = GameBatch(n_games, lib=lib)
    res = batch.playout(seed=seed, first_game_id=first_game_id, flags=flags | PLAYOUT_HASH)
    assert batch.is_terminal().all()
    scores = batch.scores()
    hist = batch.history()
    policy = 1 if flags & PLAYOUT_MIN_TILE else (2 if flags & PLAYOUT_MAX_TILE else 0)
    idx = range(n_games) if n_check is None else np.linspace(0, n_games - 1, n_check).astype(int)
    for g in idx:
        ref = orc.playout(seed, first_game_id + int(g), policy)
        assert ref["n_plies"] == int(res["steps"][g])
        assert ref["hash"] == int(res["hash"][g]), f"trace hash differs, game {g}"
        assert list(ref["scores"]) == scores[g].tolist()
        assert [t for _, t in hist[g]] == ref["tiles"].tolist()
        assert [p for p, _ in hist[g]] == ref["players"].tolist()
    assert int(res["total_steps"]) == int(np.sum(res["steps"]))
    batch.close()
    return res


def check_illegal_move(lib, orc):
    """An illegal tile is rejected with an error and the game is left untouched."""
    batch = GameBatch(3, lib=lib)
    before = batch.digest().copy()
    st = batch.apply([399, 0, 5], strict=False)  # 399 and 5 are not legal for player 0 at the start
    assert st.tolist() == [-3, 0, -3]
    after = batch.digest()
    assert after[0] == before[0] and after[2] == before[2] and after[1] != before[1]
    try:
        batch.apply([399, -1, -1])
        raised = False
    except BkError as e:
        raised = e.code == -3 and "Invalid move" in str(e)
    assert raised
    batch.close()


def check_place_piece(lib, orc, seed=0, n_turns=40):
    """Game::place_piece against the oracle: random (p, v, o) proposals, legal and illegal."""
    rng = np.random.default_rng(seed)
    batch = GameBatch(1, lib=lib)
    og = orc.Game()
    done_legal = 0
    for _ in range(n_turns):
        if og.is_terminal():
            break
        pieces = og.pieces(og.current_player())
        # a few random proposals (mostly illegal), then a legal one built from the oracle's anchors
        for _k in range(3):
            p = int(rng.integers(0, len(pieces) + 1))
            v = int(rng.integers(0, 9))
            o = int(rng.integers(0, 400))
            probe = og.clone()
            rc = probe.place_piece(p, v, o)
            st = batch.clone().place_piece([p], [v], [o], strict=False)
            assert (st[0] == 0) == (rc == 0), (p, v, o, rc, st)
        found = None
        anchors = og.anchors(-1)
        order = rng.permutation(len(pieces))
        for p in order:
            nv = orc.piece_num_variants(pieces[p])
            for v in rng.permutation(nv):
                var = orc.piece_variant(pieces[p], int(v))
                for a in anchors:
                    for off in var["offsets"]:
                        if off <= a:
                            probe = og.clone()
                            if probe.place_piece(int(p), int(v), a - off) == 0:
                                found = (int(p), int(v), a - off)
                                break
                    if found:
                        break
                if found:
                    break
            if found:
                break
        assert found is not None
        assert og.place_piece(*found) == 0
        st = batch.place_piece([found[0]], [found[1]], [found[2]])
        assert st[0] == 0
        compare_state(batch, [og])
        done_legal += 1
    batch.close()
    return done_legal


def check_piece_to_finish(lib, orc, seed=0, n_steps=60):
    """Game::apply(tile, Some(p)) (game.rs:176-187): the reference commits remaining-list entry p
    blindly after this tile; random mixes of None / Some(p) must track the oracle exactly."""
    rng = np.random.default_rng(seed)
    batch = GameBatch(1, lib=lib)
    og = orc.Game()
    for _ in range(n_steps):
        if og.is_terminal():
            break
        pieces = og.pieces(og.current_player())
        lt = og.legal_tiles()
        tile = int(lt[rng.integers(len(lt))])
        fin = int(rng.integers(len(pieces))) if rng.random() < 0.3 else None
        assert og.apply(tile, fin)
        st = batch.apply([tile], None if fin is None else [fin])
        assert st[0] == 0
        compare_state(batch, [og])
    # out-of-range piece index: rejected, game untouched (the reference would panic in Vec::remove)
    if not og.is_terminal():
        before = batch.digest()[0]
        lt = og.legal_tiles()
        st = batch.apply([lt[0]], [21], strict=False)
        assert st[0] == -3 and batch.digest()[0] == before
    batch.close()


def check_selfplay_stub(lib, orc, n_games, cfg_kwargs, first_game_id=0, max_plies=-1, n_check=None):
    """training_game() with the fixed-prior stub: per-ply root visit vectors, action trace, priors and
    value sums of the last root — all bit-exact against the oracle (visit counts exact; Q = W/N follows)."""
    from blokus_self_play import SelfPlay, Config
    cfg = Config(**cfg_kwargs)
    ocfg = orc.make_config(cfg.sims_per_move, cfg.sample_moves, float(cfg.c_base), float(cfg.c_init),
                           float(cfg.dirichlet_alpha), float(cfg.exploration_fraction), cfg.seed)
    sp = SelfPlay(n_games, cfg, first_game_id=first_game_id, lib=lib)
    sp.run_stub(max_plies)
    recs = sp.policy_records()
    hist = sp.env.history()
    roots = sp.last_root()
    pay = sp.env.payoff()
    term = sp.env.is_terminal()
    ctr = sp.counters()
    idx = range(n_games) if n_check is None else np.linspace(0, n_games - 1, n_check).astype(int)
    for g in idx:
        ref = orc.selfplay_game(ocfg, first_game_id + int(g), max_plies=max_plies)
        assert ref["n_plies"] == len(recs[g]), (g, ref["n_plies"], len(recs[g]))
        assert [t for _, t in hist[g]] == ref["tiles"].tolist(), f"action trace differs, game {g}"
        assert [p for p, _ in hist[g]] == ref["players"].tolist()
        for k in range(ref["n_plies"]):
            tiles, visits = recs[g][k]
            assert np.array_equal(tiles, ref["roots"][k]["tile"]), f"root children differ, game {g} ply {k}"
            assert np.array_equal(visits, ref["roots"][k]["visits"]), f"visit counts differ, game {g} ply {k}"
            assert int(visits.sum()) == cfg.sims_per_move
        last = ref["roots"][-1]
        assert np.array_equal(roots[g]["tile"], last["tile"])
        assert np.array_equal(roots[g]["visits"], last["visits"])
        assert np.array_equal(roots[g]["prior"], last["prior"]), f"priors (Dirichlet noise) differ, game {g}"
        # Q within 1e-5 relative (north star); in fact the f32 value sums are identical
        q_dev = roots[g]["value_sum"] / np.maximum(roots[g]["visits"], 1)
        q_ref = last["value_sum"] / np.maximum(last["visits"], 1)
        assert np.allclose(q_dev, q_ref, rtol=1e-5, atol=0)
        assert np.array_equal(roots[g]["value_sum"], last["value_sum"])
        if max_plies < 0:
            assert term[g] and pay[g].tolist() == ref["payoff"].tolist()
    total_plies = sum(len(r) for r in recs)
    assert ctr["sims"] == total_plies * cfg.sims_per_move
    sp.close()
    return {"plies": total_plies, "counters": ctr}


def fixed_network(seed=0):
    """A constant 'network': policy = table (mover frame) x legal mask (plane 4), value = fixed relative-seat
    vector — what model/resnet.py:84-92 returns for a net that ignores its input.  Non-uniform, so selection is
    not all ties, and the frame rotation of the policy and the rotate_right of the value both matter.  When a
    position has two or more legal tiles the first one (mover frame) gets p == 0 and must be dropped
    (simulation.rs:70); a position whose legal tiles all had p <= 0 would panic the reference (unwrap on None)."""
    rng = np.random.default_rng(seed)
    table = rng.uniform(0.05, 3.0, size=400).astype(np.float32)
    vrel = np.array([0.4, 0.3, 0.2, 0.1], dtype=np.float32)

    def one(mask400):
        pol = (table * mask400).astype(np.float32)
        nz = np.flatnonzero(pol)
        if len(nz) >= 2:
            pol[nz[0]] = 0.0
        return pol

    def batched(planes):
        pl = np.asarray(planes)
        n = pl.shape[0]
        return np.stack([one(pl[i, 4].reshape(400)) for i in range(n)]), np.tile(vrel, (n, 1))

    def single(_gid, planes):
        return one(np.asarray(planes, dtype=np.float32)[4].reshape(400)), vrel

    return batched, single


def check_selfplay_evaluator(lib, orc, n_games, cfg_kwargs, first_game_id=0, max_plies=4, xp="numpy", net_seed=0):
    """The external-evaluator protocol (begin_ply / leaf_planes / expand_backup / end_ply) with a fixed
    non-uniform network, against the oracle driven by the same network through its evaluator callback."""
    from blokus_self_play import SelfPlay, Config, host_evaluator
    batched, single = fixed_network(net_seed)
    cfg = Config(**cfg_kwargs)
    ocfg = orc.make_config(cfg.sims_per_move, cfg.sample_moves, float(cfg.c_base), float(cfg.c_init),
                           float(cfg.dirichlet_alpha), float(cfg.exploration_fraction), cfg.seed)
    sp = SelfPlay(n_games, cfg, first_game_id=first_game_id, lib=lib)
    info = sp.run_evaluator(batched if xp == "numpy" else host_evaluator(batched), max_plies=max_plies, buffers=_buffers(xp))
    recs = sp.policy_records()
    hist = sp.env.history()
    roots = sp.last_root()
    for g in range(n_games):
        ref = orc.selfplay_game(ocfg, first_game_id + g, max_plies=max_plies, evaluator=single)
        assert ref["n_plies"] == len(recs[g])
        assert [t for _, t in hist[g]] == ref["tiles"].tolist(), f"action trace differs, game {g}"
        for k in range(ref["n_plies"]):
            assert np.array_equal(recs[g][k][0], ref["roots"][k]["tile"]), f"children differ, game {g} ply {k}"
            assert np.array_equal(recs[g][k][1], ref["roots"][k]["visits"]), f"visits differ, game {g} ply {k}"
        last = ref["roots"][-1]
        assert np.array_equal(roots[g]["prior"], last["prior"])
        assert np.array_equal(roots[g]["value_sum"], last["value_sum"])
    sp.close()
    return info


class FakeQueue:
    """Stands in for multiprocessing.Manager().Queue + the server side of the Pipe (model/training.py:194-201):
    put() runs the 'model' at once and the answer is handed out by recv()."""

    def __init__(self, model):
        self.model = model
        self.answers = []
        self.requests = 0

    def put(self, item):
        gid, planes = item
        assert isinstance(planes, list) and len(planes) == 5 and len(planes[0]) == 20 and isinstance(planes[0][0][0], bool)
        pol, val = self.model(gid, np.asarray(planes, dtype=np.float32))
        self.answers.append((np.asarray(pol, dtype=np.float32).tolist(), np.asarray(val, dtype=np.float32).tolist()))
        self.requests += 1

    def recv(self):
        return self.answers.pop(0)


def check_play_training_game(lib, orc, cfg_kwargs, game_id=7, xp="torch"):
    """The reference's own entry point signature, play_training_game(id, config, inference_queue, pipe)."""
    from blokus_self_play import play_training_game, Config
    _, single = fixed_network(1)
    q = FakeQueue(single)
    cfg = Config(**cfg_kwargs)
    history, policies, values = play_training_game(game_id, cfg, q, q, lib=lib, buffers=_buffers(xp))
    ocfg = orc.make_config(cfg.sims_per_move, cfg.sample_moves, float(cfg.c_base), float(cfg.c_init),
                           float(cfg.dirichlet_alpha), float(cfg.exploration_fraction), cfg.seed)
    ref = orc.selfplay_game(ocfg, game_id, evaluator=single)
    assert len(history) == len(policies) == ref["n_plies"]                 # simulation.rs:293-295
    assert [t for _, t in history] == ref["tiles"].tolist() and [p for p, _ in history] == ref["players"].tolist()
    assert values == ref["payoff"].tolist()
    for k, pol in enumerate(policies):
        vis = ref["roots"][k]["visits"]
        probs = vis.astype(np.float32) / np.float32(vis.sum())
        assert [t for t, _ in pol] == ref["roots"][k]["tile"].tolist()
        assert np.array_equal(np.array([p for _, p in pol], dtype=np.float32), probs)
    # one request per evaluated position: the root of every ply plus every non-terminal leaf
    assert q.requests >= ref["n_plies"]
    return len(history)


def check_training_tensors(lib, orc, n_games, cfg_kwargs, max_plies, xp):
    """Device-side `save()` (model/training.py:70-119) against its numpy restatement applied to the tuples
    the same run returns — and the state planes against Game::get_board_state of the oracle game at that ply."""
    from blokus_self_play import SelfPlay, Config
    from oracle.save_oracle import save_arrays
    sp = SelfPlay(n_games, Config(**cfg_kwargs), first_game_id=2, lib=lib)
    sp.run_stub(max_plies)
    st, po, va, offs = sp.training_tensors(buffers=_buffers(xp))
    if xp == "torch":
        st, po, va = st.cpu().numpy(), po.cpu().numpy(), va.cpu().numpy()
    data = sp.game_data()
    assert offs[-1] == sum(len(d[0]) for d in data) == len(st)
    for g, game in enumerate(data):
        rs, rp, rv = save_arrays(game)
        a, b = int(offs[g]), int(offs[g + 1])
        assert np.array_equal(st[a:b], rs), f"states differ, game {g}"
        assert np.array_equal(po[a:b], rp), f"policies differ, game {g}"
        assert np.array_equal(va[a:b], rv), f"values differ, game {g}"
        # with the stub every legal tile is a root child, so the planes equal get_board_state() ply by ply
        og = orc.Game()
        for i, (_, tile) in enumerate(game[0]):
            assert np.array_equal(st[a + i], og.board_state().astype(np.float32)), f"planes != get_board_state, game {g} ply {i}"
            og.apply(tile)
    sp.close()
    return len(st)


def check_arena(lib, orc, n_games=2, seed=5):
    """play_test_game / play_test_games (simulation.rs:233-265,298-332) against the oracle's restatement."""
    from blokus_self_play import play_test_games, play_test_game
    bm, sm = fixed_network(11)
    bb, sb = fixed_network(12)
    ids = list(range(30, 30 + n_games))
    scores, hists = play_test_games(ids, bm, bb, seed=seed, lib=lib)
    for g, gid in enumerate(ids):
        ref = orc.test_game(gid, sm, sb, seed=seed)
        assert [t for _, t in hists[g]] == ref["tiles"].tolist(), f"arena trace differs, game {gid}"
        assert [p for p, _ in hists[g]] == ref["players"].tolist()
        assert scores[g] == ref["score"]
    # the reference's own signature and IPC protocol, one game
    qm, qb = FakeQueue(sm), FakeQueue(sb)
    shared = []
    qm.answers = shared
    qb.answers = shared

    class Pipe:
        def recv(self):
            return shared.pop(0)
    s = play_test_game(ids[0], qm, qb, Pipe(), seed=seed, lib=lib)
    assert s == scores[0]
    assert qm.requests > 0 and qb.requests > 0 and qm.requests + qb.requests == len(hists[0])
    return scores


def _records_equal(a, b):
    ra, rb = a.policy_records(), b.policy_records()
    assert len(ra) == len(rb)
    for ga, gb in zip(ra, rb):
        assert len(ga) == len(gb)
        for (t1, v1), (t2, v2) in zip(ga, gb):
            assert np.array_equal(t1, t2) and np.array_equal(v1, v2)


def check_throughput_modes(lib, n_games, cfg_kwargs, max_plies, xp, leaves=4, net_seed=0):
    """SURVEY §8f row f3 — the opt-in throughput modes have no reference counterpart, so they are checked
    against the exact mode and through invariants:
      * multi-leaf code path with ONE leaf per round == exact mode, bit for bit;
      * SKIP_FORCED leaves the training tuple unchanged (external protocol and fused stub kernel);
      * K leaves per round: every ply's root visits sum to sims_per_move, fewer evaluator rounds, game valid."""
    from blokus_self_play import SelfPlay, Config, host_evaluator, MODE_SKIP_FORCED, MODE_FORCE_MULTI_LEAF
    batched, _ = fixed_network(net_seed)
    ev = batched if xp == "numpy" else host_evaluator(batched)
    cfg = Config(**cfg_kwargs)

    exact = SelfPlay(n_games, cfg, lib=lib)
    info_exact = exact.run_evaluator(ev, max_plies=max_plies, buffers=_buffers(xp))

    one = SelfPlay(n_games, cfg, lib=lib)
    one.set_mode(MODE_FORCE_MULTI_LEAF, 1)
    one.run_evaluator(ev, max_plies=max_plies, buffers=_buffers(xp))
    assert one.env.history() == exact.env.history()
    _records_equal(one, exact)
    for x, y in zip(one.last_root(), exact.last_root()):
        assert np.array_equal(x["value_sum"], y["value_sum"]) and np.array_equal(x["prior"], y["prior"])
    assert one.counters() == exact.counters()
    one.close()

    skip = SelfPlay(n_games, cfg, lib=lib)
    skip.set_mode(MODE_SKIP_FORCED, 1)
    info_skip = skip.run_evaluator(ev, max_plies=max_plies, buffers=_buffers(xp))
    assert skip.env.history() == exact.env.history()
    _records_equal(skip, exact)
    assert skip.counters()["sims"] <= exact.counters()["sims"]
    skip.close()

    multi = SelfPlay(n_games, cfg, lib=lib)
    multi.set_mode(0, leaves)
    info_multi = multi.run_evaluator(ev, max_plies=max_plies, buffers=_buffers(xp))
    recs = multi.policy_records()
    hist = multi.env.history()
    for g in range(n_games):
        assert len(recs[g]) == info_multi["plies"] or multi.env.is_terminal()[g]
        for k, (tiles, visits) in enumerate(recs[g]):
            assert int(visits.sum()) == cfg.sims_per_move, (g, k, visits)
            assert np.all(np.diff(tiles) > 0)
            assert hist[g][k][1] in tiles.tolist()
    assert multi.counters()["sims"] == sum(len(r) for r in recs) * cfg.sims_per_move
    assert info_multi["rounds"] < info_exact["rounds"]
    # determinism of the multi-leaf mode
    again = SelfPlay(n_games, cfg, lib=lib)
    again.set_mode(0, leaves)
    again.run_evaluator(ev, max_plies=max_plies, buffers=_buffers(xp))
    assert again.env.history() == hist
    _records_equal(again, multi)
    again.close()
    multi.close()
    exact.close()
    return {"rounds_exact": info_exact["rounds"], "rounds_multi": info_multi["rounds"], "rounds_skip": info_skip["rounds"]}


def check_skip_forced_stub(lib, n_games, cfg_kwargs, max_plies):
    """Fused stub kernel with SKIP_FORCED: same histories and policy records (as visit vectors: a forced root's
    single child holds all sims_per_move visits either way), fewer simulations run."""
    from blokus_self_play import SelfPlay, Config, MODE_SKIP_FORCED
    cfg = Config(**cfg_kwargs)
    a = SelfPlay(n_games, cfg, lib=lib)
    a.run_stub(max_plies)
    b = SelfPlay(n_games, cfg, lib=lib)
    b.set_mode(MODE_SKIP_FORCED, 1)
    b.run_stub(max_plies)
    assert a.env.history() == b.env.history()
    _records_equal(a, b)
    ca, cb = a.counters(), b.counters()
    a.close(); b.close()
    return ca["sims"], cb["sims"]


def check_tree_reuse(lib, n_games, cfg_kwargs, max_plies, xp, net_seed=0):
    """Row f3, tree reuse (opt-in; no reference counterpart — the reference starts a new tree every ply):
      * every recorded policy sums to sims_per_move visits, children ascending, the played tile among them;
      * fewer simulations are run than in the exact mode (the kept subtree's visits are not repeated);
      * deterministic; the fused stub kernel and the evaluator protocol driven by the stub's own policy/value agree;
      * stop-and-resume (max_plies prefixes) equals one run; combined with the forced-ply shortcut and with
        multi-leaf rounds the invariants still hold."""
    from blokus_self_play import SelfPlay, Config, host_evaluator, MODE_TREE_REUSE, MODE_SKIP_FORCED
    cfg = Config(**cfg_kwargs)

    def invariants(sp, plies):
        recs, hist = sp.policy_records(), sp.env.history()
        for g in range(n_games):
            assert len(recs[g]) == min(plies, len(hist[g])) or sp.env.is_terminal()[g]
            for k, (tiles, visits) in enumerate(recs[g]):
                assert int(visits.sum()) == cfg.sims_per_move, (g, k, visits)
                assert np.all(np.diff(tiles) > 0) and hist[g][k][1] in tiles.tolist()

    exact = SelfPlay(n_games, cfg, lib=lib)
    exact.run_stub(max_plies)
    sims_exact = exact.counters()["sims"]
    exact.close()

    a = SelfPlay(n_games, cfg, lib=lib)
    a.set_mode(MODE_TREE_REUSE, 1)
    a.run_stub(max_plies)
    invariants(a, max_plies)
    assert a.counters()["sims"] < sims_exact

    b = SelfPlay(n_games, cfg, lib=lib)                       # prefixes compose (the kept tree survives between launches)
    b.set_mode(MODE_TREE_REUSE, 1)
    b.run_stub(max_plies // 2)
    b.run_stub(max_plies - max_plies // 2)
    assert b.env.history() == a.env.history()
    _records_equal(a, b)
    b.close()

    def stub_eval(pl):                                        # the stub as an external evaluator
        pl = np.asarray(pl.detach().cpu().numpy() if hasattr(pl, "detach") else pl)
        return pl[:, 4].reshape(-1, 400).astype(np.float32).copy(), np.full((pl.shape[0], 4), 0.25, dtype=np.float32)
    ev = stub_eval if xp == "numpy" else host_evaluator(stub_eval)
    c = SelfPlay(n_games, cfg, lib=lib)
    c.set_mode(MODE_TREE_REUSE, 1)
    c.run_evaluator(ev, max_plies=max_plies, buffers=_buffers(xp))
    assert c.env.history() == a.env.history()
    _records_equal(a, c)
    assert c.counters()["sims"] == a.counters()["sims"]
    c.close()

    d = SelfPlay(n_games, cfg, lib=lib)                       # with the forced-ply shortcut and 3 leaves per round
    d.set_mode(MODE_TREE_REUSE | MODE_SKIP_FORCED, 3)
    batched, _ = fixed_network(net_seed)
    d.run_evaluator(batched if xp == "numpy" else host_evaluator(batched), max_plies=max_plies, buffers=_buffers(xp))
    invariants(d, max_plies)
    d.close()
    sims_reuse = a.counters()["sims"]
    a.close()
    return sims_exact, sims_reuse
