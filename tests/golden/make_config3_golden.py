"""Generates tests/golden/config3_games.json: COMPLETE self-play games at BASELINE.json config 3's real parameters
(800 sims/move, fixed uniform priors, Dirichlet alpha 0.03, frac 0.25, sample_moves 30) from the CPU oracle
(oracle/mcts_oracle.hpp, the restatement of self_play/src/simulation.rs:174-231,267-296).

Per game: the action trace, the payoff, and for EVERY ply three 64-bit digests of the root's child block
    v = blake2b(tiles int16 LE || visits uint32 LE)      visit vector (north star: exact)
    w = blake2b(value_sum float32 LE bits)               value sums (=> Q bit-exact, north star asks 1e-5 relative)
    p = blake2b(prior float32 LE bits)                   priors after the root noise
The reference itself is Rust and cannot run here (no rustc/cargo): these are vectors of the RESTATEMENT, frozen so
that neither the oracle nor the CUDA path can drift, and so the GPU tests can check complete 800-sim games (the oracle
needs minutes per game) against committed data.

    python tests/golden/make_config3_golden.py        (from the repo root; ~10 min on 8 cores)
"""
import hashlib
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

SEED = 20261018
CONFIG3 = dict(sims_per_move=800, sample_moves=30, c_base=19652.0, c_init=1.25, dirichlet_alpha=0.03,
               exploration_fraction=0.25, seed=SEED)
GAME_IDS = [0, 1, 2, 3, 137, 512, 777, 1023]     # all inside config 3's 1024-game batch; 0..3 consecutive


def digest(*arrays) -> str:
    h = hashlib.blake2b(digest_size=8)
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def root_digests(tile, visits, value_sum, prior):
    return {"v": digest(np.asarray(tile, dtype="<i2"), np.asarray(visits, dtype="<u4")),
            "w": digest(np.asarray(value_sum, dtype="<f4")), "p": digest(np.asarray(prior, dtype="<f4")),
            "n": int(len(tile))}


def one(gid):
    r = orc.selfplay_game(orc.make_config(**CONFIG3), gid, max_plies=-1)
    return {"game_id": gid, "n_plies": int(r["n_plies"]), "tiles": r["tiles"].tolist(), "players": r["players"].tolist(),
            "payoff_hex": [float(x).hex() for x in r["payoff"]], "sims": int(r["sims"]),
            "roots": [root_digests(x["tile"], x["visits"], x["value_sum"], x["prior"]) for x in r["roots"]]}


if __name__ == "__main__":
    orc.lib()
    with ThreadPoolExecutor(max_workers=min(len(GAME_IDS), os.cpu_count() or 1)) as ex:
        games = list(ex.map(one, GAME_IDS))
    out = {"config": CONFIG3, "games": games}
    path = os.path.join(ROOT, "tests", "golden", "config3_games.json")
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", sum(g["n_plies"] for g in games), "plies,",
          sum(g["sims"] for g in games), "sims")
