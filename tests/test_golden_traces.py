"""Committed golden vectors (tests/golden/traces.json, made by tests/golden/make_trace_golden.py):
the oracle must still reproduce them (CPU), and the sm_100a library must reproduce them through the C ABI (GPU)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traces.json")))


def test_oracle_reproduces_golden_playouts(orc):
    for p in GOLD["playouts"]:
        r = orc.playout(GOLD["seed"], p["game_id"], p["policy"])
        assert r["n_plies"] == p["n_plies"] and r["tiles"].tolist() == p["tiles"] and r["players"].tolist() == p["players"]
        assert r["legal_counts"].tolist() == p["legal_counts"]
        assert [int(x) for x in r["scores"]] == p["scores"] and int(r["hash"]) == p["hash"]
        assert [float(x) for x in r["payoff"]] == p["payoff"]


def test_oracle_reproduces_golden_selfplay(orc):
    for s in GOLD["selfplay"]:
        if s["case"] == "config3_prefix" and s["game_id"] != 0:
            continue                      # 800-sim cases are slow on the CPU: one is enough here
        r = orc.selfplay_game(orc.make_config(**s["config"]), s["game_id"], max_plies=s["max_plies"])
        assert r["tiles"].tolist() == s["tiles"]
        for a, b in zip(r["roots"], s["roots"]):
            assert a["tile"].tolist() == b["tile"] and a["visits"].tolist() == b["visits"]
        assert [float(v).hex() for v in r["roots"][-1]["value_sum"]] == s["last_root_value_sum_hex"]
        assert [float(v).hex() for v in r["roots"][-1]["prior"]] == s["last_root_prior_hex"]


@pytest.mark.gpu
def test_gpu_reproduces_golden_playouts(cuda_lib):
    from blokus_self_play import GameBatch, PLAYOUT_HASH, PLAYOUT_MIN_TILE, PLAYOUT_MAX_TILE
    for p in GOLD["playouts"]:
        b = GameBatch(1, lib=cuda_lib)
        flags = PLAYOUT_HASH | {0: 0, 1: PLAYOUT_MIN_TILE, 2: PLAYOUT_MAX_TILE}[p["policy"]]
        r = b.playout(seed=GOLD["seed"], first_game_id=p["game_id"], flags=flags)
        assert int(r["steps"][0]) == p["n_plies"] and int(r["hash"][0]) == p["hash"]
        assert b.history()[0] == list(zip(p["players"], p["tiles"]))
        assert b.scores()[0].tolist() == p["scores"] and b.payoff()[0].tolist() == p["payoff"]
        b.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_selfplay(cuda_lib):
    from blokus_self_play import SelfPlay, Config
    for s in GOLD["selfplay"]:
        sp = SelfPlay(1, Config(**s["config"]), first_game_id=s["game_id"], lib=cuda_lib)
        sp.run_stub(s["max_plies"])
        assert sp.env.history()[0] == list(zip(s["players"], s["tiles"]))
        for (tiles, visits), ref in zip(sp.policy_records()[0], s["roots"]):
            assert tiles.tolist() == ref["tile"] and visits.tolist() == ref["visits"]
        last = sp.last_root()[0]
        assert [float(v).hex() for v in last["value_sum"]] == s["last_root_value_sum_hex"]
        assert [float(v).hex() for v in last["prior"]] == s["last_root_prior_hex"]
        sp.close()
