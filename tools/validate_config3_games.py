"""One-off validation beyond the committed fixtures: COMPLETE config-3 games (800 sims/move, alpha 0.03, frac 0.25) of
more global ids, oracle (CPU, one game per thread) against the fused kernel run at the full 1024-game width — every
ply's root children and visit vector, the action trace and the payoff.  Prints one JSON line (kept under profiles/)."""
import argparse, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import numpy as np
from blokus_self_play import SelfPlay, Config
from oracle import oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=16)
ap.add_argument("--count", type=int, default=32)
ap.add_argument("--seed", type=int, default=20261018)
ap.add_argument("--widths", default="1024", help="batch widths to run the fused kernel at (1024: two-warp pipeline; 8192: one warp per game, 28 per SM)")
a = ap.parse_args()
kw = dict(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=a.seed)
runs = {}
for width in [int(w) for w in a.widths.split(",")]:
    sp = SelfPlay(width, Config(**kw), first_game_id=0)
    ms = sp.run_stub(-1)
    ply_off, ply_ptr, tiles_p, visits_p = sp.policy_records_packed()      # packed: no per-ply Python objects for the games not checked
    want = range(a.first, a.first + a.count)
    recs = {g: [(tiles_p[ply_ptr[k]:ply_ptr[k + 1]].copy(), visits_p[ply_ptr[k]:ply_ptr[k + 1]].copy()) for k in range(int(ply_off[g]), int(ply_off[g + 1]))] for g in want}
    full_hist = sp.env.history()
    runs[width] = (ms, {g: full_hist[g] for g in want}, recs, sp.env.payoff())
    del full_hist
    sp.close()
ocfg = orc.make_config(**{**kw, "c_base": 19652.0})
orc.lib()
ids = list(range(a.first, a.first + a.count))
t0 = time.time()
with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
    refs = list(ex.map(lambda g: orc.selfplay_game(ocfg, g, max_plies=-1), ids))
cpu_s = time.time() - t0
out = {"config": "config-3 games (800 sims/move, stub, alpha 0.03, frac 0.25), complete, fused kernel at each batch width (global ids from 0: the "
                 "same games at every width; 1024 = two-warp pipeline, 8192 = one warp per game at 28 per SM)",
       "checked_ids": [ids[0], ids[-1]], "games_checked": len(ids), "oracle_seconds": cpu_s, "oracle_threads": os.cpu_count(), "widths": {}}
any_bad = False
for width, (ms, hist, recs, pay) in runs.items():
    bad = []
    plies = 0
    for g, ref in zip(ids, refs):
        ok = [t for _, t in hist[g]] == ref["tiles"].tolist() and [p for p, _ in hist[g]] == ref["players"].tolist() and len(recs[g]) == ref["n_plies"]
        ok = ok and pay[g].tolist() == ref["payoff"].tolist()
        if ok:
            for k in range(ref["n_plies"]):
                if not (np.array_equal(recs[g][k][0], ref["roots"][k]["tile"]) and np.array_equal(recs[g][k][1], ref["roots"][k]["visits"])):
                    ok = False
                    break
        plies += ref["n_plies"]
        if not ok:
            bad.append(g)
    any_bad = any_bad or bool(bad)
    out["widths"][str(width)] = {"plies_checked": plies, "sims_checked": 800 * plies, "mismatching_games": bad, "identical": not bad, "gpu_kernel_ms": ms}
out["identical"] = not any_bad
print(json.dumps(out))
sys.exit(1 if any_bad else 0)
