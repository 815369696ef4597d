"""BASELINE.json's own configurations at their REAL sizes against committed oracle vectors:

  config 2  tests/golden/config2_hashes.json  — all 4096 games' trace hashes (make_config2_golden.py)
  config 3  tests/golden/config3_games.json   — 8 COMPLETE games at 800 sims/move, alpha 0.03, frac 0.25: every ply's
                                                 root visit vector, value sums and noisy priors as digests, the action
                                                 trace and the payoff (make_config3_golden.py)

CPU side: the oracle still reproduces (a bounded sample of) what it wrote.  GPU side (-m gpu): the sm_100a library
through the C ABI reproduces ALL of it, bit for bit — the fused kernel at the full 1024-game width, ply-by-ply
stepping (value sums / priors of every root), and the external-evaluator protocol."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
C2 = json.load(open(os.path.join(HERE, "golden", "config2_hashes.json")))
C3 = json.load(open(os.path.join(HERE, "golden", "config3_games.json")))


def digest(*arrays) -> str:
    h = hashlib.blake2b(digest_size=8)
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def vdig(tile, visits):
    return digest(np.asarray(tile, dtype="<i2"), np.asarray(visits, dtype="<u4"))


def fdig(x):
    return digest(np.asarray(x, dtype="<f4"))


# ---- CPU: the oracle against its own committed vectors ----------------------------------------------------------
def test_oracle_reproduces_config2_sample(orc):
    r = orc.playout_batch(C2["seed"], 0, 256, n_threads=os.cpu_count() or 1, want_hash=True)
    assert [format(int(h), "016x") for h in r["hashes"]] == C2["hash_hex"][:256]
    assert r["plies"].tolist() == C2["plies"][:256] and r["scores"].tolist() == C2["scores"][:256]
    assert sum(C2["plies"]) == C2["total_plies"] and len(C2["hash_hex"]) == C2["n_games"] == 4096


def test_oracle_reproduces_config3_first_plies(orc):
    cfg = orc.make_config(**C3["config"])
    for game in C3["games"][:3]:
        r = orc.selfplay_game(cfg, game["game_id"], max_plies=2)
        assert r["tiles"].tolist() == game["tiles"][:2]
        for k, root in enumerate(r["roots"]):
            ref = game["roots"][k]
            assert vdig(root["tile"], root["visits"]) == ref["v"] and fdig(root["value_sum"]) == ref["w"]
            assert fdig(root["prior"]) == ref["p"] and len(root["tile"]) == ref["n"]
    for game in C3["games"]:
        assert game["n_plies"] == len(game["tiles"]) == len(game["roots"]) and game["sims"] == 800 * game["n_plies"]


# ---- GPU ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_config2_all_4096_trace_hashes(cuda_lib, orc):
    """Config 2 at full size: EVERY game's trace hash, ply count and final scores equal the committed oracle vectors,
    and equal the oracle run live on the box for all 4096 ids."""
    from blokus_self_play import GameBatch, PLAYOUT_HASH
    b = GameBatch(4096, lib=cuda_lib)
    r = b.playout(seed=C2["seed"], first_game_id=0, flags=PLAYOUT_HASH)
    assert [format(int(h), "016x") for h in r["hash"]] == C2["hash_hex"]
    assert r["steps"].tolist() == C2["plies"] and b.scores().tolist() == C2["scores"]
    assert int(r["total_steps"]) == C2["total_plies"]
    live = orc.playout_batch(C2["seed"], 0, 4096, n_threads=os.cpu_count() or 1, want_hash=True)
    assert np.array_equal(live["hashes"], r["hash"].astype(np.uint64)) and live["plies"].tolist() == r["steps"].tolist()
    # the kernel bench.py times is the PLAIN instantiation (no per-ply digest) started with BK_PLAYOUT_NEW_GAME (reset inside
    # the launch): same batch through it — every game's complete move history, final state digest and scores must be those
    # of the run the oracle just vouched for
    from blokus_self_play import PLAYOUT_NEW_GAME
    c = GameBatch(4096, lib=cuda_lib)
    c.playout(seed=C2["seed"] + 1, first_game_id=99)                  # leftovers of another batch in the buffers
    rc = c.playout(seed=C2["seed"], first_game_id=0, flags=PLAYOUT_NEW_GAME)
    assert rc["steps"].tolist() == C2["plies"] and c.scores().tolist() == C2["scores"]
    assert np.array_equal(c.digest(), b.digest()) and c.history() == b.history()
    b.close(); c.close()


def _check_game_records(game, hist, recs, payoff):
    assert [t for _, t in hist] == game["tiles"], f"action trace differs, game {game['game_id']}"
    assert [p for p, _ in hist] == game["players"]
    assert len(recs) == game["n_plies"]
    for k, (tiles, visits) in enumerate(recs):
        assert len(tiles) == game["roots"][k]["n"] and int(visits.sum()) == 800
        assert vdig(tiles, visits) == game["roots"][k]["v"], f"root visits differ, game {game['game_id']} ply {k}"
    assert [float(x).hex() for x in payoff] == game["payoff_hex"]


@pytest.mark.gpu
def test_gpu_config3_complete_games_full_batch(cuda_lib):
    """Config 3 exactly as BASELINE.json states it — 1024 games, 800 sims/move, alpha 0.03, frac 0.25, one launch of
    the fused kernel, complete games: the 8 golden games (ids spread over the batch) agree on EVERY ply's root
    visit vector, on the action trace and on the payoff; every other game keeps the invariants."""
    from blokus_self_play import SelfPlay, Config
    sp = SelfPlay(1024, Config(**C3["config"]), first_game_id=0, lib=cuda_lib)
    sp.run_stub(-1)
    assert bool(sp.env.is_terminal().all())
    hist, recs, pay = sp.env.history(), sp.policy_records(), sp.env.payoff()
    for game in C3["games"]:
        g = game["game_id"]
        _check_game_records(game, hist[g], recs[g], pay[g])
    for g in range(1024):
        assert len(recs[g]) == len(hist[g]) and all(int(v.sum()) == 800 for _, v in recs[g])
    c = sp.counters()
    assert c["sims"] == 800 * sum(len(h) for h in hist)
    sp.close()


@pytest.mark.gpu
def test_gpu_config3_complete_games_every_root_value_sums(cuda_lib):
    """The same golden games stepped one ply per launch, reading the root's child block after each: visit vectors,
    VALUE SUMS (so Q = W/N is bit-exact; the north star asks 1e-5 relative) and noisy priors of every ply."""
    from blokus_self_play import SelfPlay, Config
    games = [g for g in C3["games"] if g["game_id"] < 4]
    sp = SelfPlay(4, Config(**C3["config"]), first_game_id=0, lib=cuda_lib)
    ply = 0
    while sp.live_games() > 0:
        live = ~sp.env.is_terminal()
        sp.run_stub(1)
        roots = sp.last_root()
        for game in games:
            g = game["game_id"]
            if not live[g]:
                continue
            ref = game["roots"][ply]
            assert vdig(roots[g]["tile"], roots[g]["visits"]) == ref["v"], (g, ply)
            assert fdig(roots[g]["value_sum"]) == ref["w"], f"value sums differ, game {g} ply {ply}"
            assert fdig(roots[g]["prior"]) == ref["p"], f"priors differ, game {g} ply {ply}"
        ply += 1
    assert ply == max(g["n_plies"] for g in games)
    hist, recs, pay = sp.env.history(), sp.policy_records(), sp.env.payoff()
    for game in games:
        _check_game_records(game, hist[game["game_id"]], recs[game["game_id"]], pay[game["game_id"]])
    sp.close()


@pytest.mark.gpu
def test_gpu_config3_complete_games_evaluator_protocol(cuda_lib):
    """The same golden games through begin_ply / leaf_planes / expand_backup / end_ply with the stub as an EXTERNAL
    device evaluator (policy 1.0 on legal tiles, value 0.25): ~2.4e5 evaluator rounds, complete games, every ply."""
    import torch
    from blokus_self_play import SelfPlay, Config
    games = [g for g in C3["games"] if g["game_id"] < 4]
    sp = SelfPlay(4, Config(**C3["config"]), first_game_id=0, lib=cuda_lib)
    quarter = torch.full((4, 4), 0.25, device="cuda")
    info = sp.run_evaluator(lambda pl: (pl[:, 4].reshape(-1, 400), quarter[: pl.shape[0]]), -1)
    assert info["plies"] == max(g["n_plies"] for g in games)
    hist, recs, pay = sp.env.history(), sp.policy_records(), sp.env.payoff()
    for game in games:
        _check_game_records(game, hist[game["game_id"]], recs[game["game_id"]], pay[game["game_id"]])
    last = sp.last_root()
    for game in games:                    # (a game that ended earlier has had its tree reset by the later begin_ply calls)
        if game["n_plies"] == info["plies"]:
            assert fdig(last[game["game_id"]]["value_sum"]) == game["roots"][-1]["w"]
    sp.close()
