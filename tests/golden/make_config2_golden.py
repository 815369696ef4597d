"""Generates tests/golden/config2_hashes.json: BASELINE.json config 2 (4096 lockstep random-playout games, seed 777,
global ids 0..4095) played to the end by the CPU oracle (oracle/blokus_oracle.hpp, the restatement of
blokus/src/board.rs + game.rs).  Per game: the trace hash (a digest of the FULL game state after every ply — own
rows, legal set, remaining pieces, last piece lengths, seat, eliminated mask), the number of plies and the final scores.
The reference itself is Rust and cannot run here; these are vectors of the restatement (see make_trace_golden.py).

    python tests/golden/make_config2_golden.py        (from the repo root; ~20 s on 8 cores)
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

SEED, N = 777, 4096
r = orc.playout_batch(SEED, 0, N, n_threads=os.cpu_count() or 1, want_hash=True)
out = {"seed": SEED, "n_games": N, "total_plies": int(r["steps"]), "hash_hex": [format(int(h), "016x") for h in r["hashes"]],
       "plies": r["plies"].tolist(), "scores": r["scores"].tolist()}
path = os.path.join(ROOT, "tests", "golden", "config2_hashes.json")
json.dump(out, open(path, "w"), separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes;", out["total_plies"], "plies in", round(r["seconds"], 1), "s")
