"""Generates tests/golden/traces.json from the CPU oracle (oracle/, the C++ restatement of blokus/src/*.rs and
self_play/src/simulation.rs).  The reference itself is Rust and cannot run here (no rustc/cargo), so these are
vectors of the RESTATEMENT, pinned to the reference only as far as tests/test_oracle_*.py pin the oracle; they
freeze today's behaviour so that neither the oracle nor the CUDA path can drift silently, and they let the GPU
tests run against committed data.

    python tests/golden/make_trace_golden.py        (from the repo root; rewrites traces.json)
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

SEED = 20261018
out = {"seed": SEED, "playouts": [], "selfplay": []}
# config 1: single games, seeded random legal moves (policy 0), plus the two seed-free traces (policies 1, 2)
for game_id, policy in [(0, 0), (1, 0), (4095, 0), (65535, 0), (0, 1), (0, 2)]:
    r = orc.playout(SEED, game_id, policy)
    out["playouts"].append({"game_id": game_id, "policy": policy, "n_plies": int(r["n_plies"]),
                            "tiles": r["tiles"].tolist(), "players": r["players"].tolist(),
                            "legal_counts": r["legal_counts"].tolist(), "scores": [int(x) for x in r["scores"]],
                            "payoff": [float(x) for x in r["payoff"]], "hash": int(r["hash"])})
# config 3 parameters on a prefix, and a small-simulation whole game
CASES = [
    ("config3_prefix", dict(sims_per_move=800, sample_moves=30, c_base=19652.0, c_init=1.25, dirichlet_alpha=0.03,
                            exploration_fraction=0.25, seed=SEED), [0, 1023], 3),
    ("shipped_config", dict(sims_per_move=50, sample_moves=30, c_base=19652.0, c_init=1.25, dirichlet_alpha=0.3,
                            exploration_fraction=0.25, seed=5), [2], 40),
]
for name, kw, ids, plies in CASES:
    cfg = orc.make_config(**kw)
    for gid in ids:
        r = orc.selfplay_game(cfg, gid, max_plies=plies)
        out["selfplay"].append({"case": name, "config": kw, "game_id": gid, "max_plies": plies,
                                "tiles": r["tiles"].tolist(), "players": r["players"].tolist(),
                                "roots": [{"tile": x["tile"].tolist(), "visits": x["visits"].tolist()} for x in r["roots"]],
                                "last_root_value_sum_hex": [float(v).hex() for v in r["roots"][-1]["value_sum"]],
                                "last_root_prior_hex": [float(v).hex() for v in r["roots"][-1]["prior"]]})
path = os.path.join(ROOT, "tests", "golden", "traces.json")
json.dump(out, open(path, "w"), separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes")
